"""Callers of the two-site update with a large local dimension (reference package `transmit/`): the bosonic
hopping chain `bhopping.BMPS` and the single-mode helpers of `cat` it uses.  Plotting (`transmit/paint.py`, the
matplotlib parts of `transmit/cat.py`) is outside the hot path and not provided."""
__all__ = ['bhopping', 'cat']
