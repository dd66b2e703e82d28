"""B200-native drop-in for the sparse-marginalization hot path of ``lvmhelpers``.

Same module and symbol names as the reference package
(deep-spin/sparse-marginalization-lvm, ``lvmhelpers/``) for the path this repo
replaces -- ``marg``, ``structmarg``, ``sparsemap``, ``pbinary_topk``,
``pbernoulli``, ``pbudget``, ``psequence``, ``utils`` -- backed by hand-written
sm_100a kernels behind the C ABI of ``include/smlvm.h`` (``libsmlvm.so``).

Importing the compute modules without the built library raises: there is no CPU
or eager fallback.

What is NOT replaced -- the competing estimators ``sfe``, ``nvil``, ``gumbel``,
``sum_and_sample`` that every ``experiments/*/train.py`` also imports -- is not
shipped here.  This package extends its ``__path__`` with any other ``lvmhelpers``
directory found LATER on ``sys.path`` (``pkgutil.extend_path``), so with this
package first and the reference checkout second on the path those submodules
resolve to the reference's own, untouched files, while every module listed above
resolves here.  ``lvmhelpers.utils`` forwards the names it does not define
(``populate_common_params``, pure argparse) the same way.
"""
import pkgutil

__path__ = pkgutil.extend_path(__path__, __name__)
__version__ = "0.2.0"
