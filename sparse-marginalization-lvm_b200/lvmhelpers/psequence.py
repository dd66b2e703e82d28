"""lvmhelpers.psequence counterpart (reference: lvmhelpers/psequence.pyx:10-37)."""
from ._factors import PFactorSequence, PFactorSequenceBinary  # noqa: F401
