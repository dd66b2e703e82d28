"""B200-native drop-in for the sparse-marginalization hot path of ``lvmhelpers``.

Same module and symbol names as the reference package
(deep-spin/sparse-marginalization-lvm, ``lvmhelpers/``) for the path this repo
replaces -- ``marg``, ``structmarg``, ``sparsemap``, ``pbinary_topk``,
``pbernoulli``, ``pbudget``, ``psequence``, ``utils`` -- backed by hand-written
sm_100a kernels behind the C ABI of ``include/smlvm.h`` (``libsmlvm.so``).

Importing the compute modules without the built library raises: there is no CPU
or eager fallback.
"""
__version__ = "0.1.0"
