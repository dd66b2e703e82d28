"""lvmhelpers.pbernoulli counterpart (reference: lvmhelpers/pbernoulli.pyx:10-22)."""
from ._factors import PFactorBernoulli  # noqa: F401
