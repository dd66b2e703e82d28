"""lvmhelpers.pbudget counterpart (reference: lvmhelpers/pbudget.pyx:10-22)."""
from ._factors import PFactorBudget  # noqa: F401
