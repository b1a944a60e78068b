"""Shim so the reference's `from props import getNode` resolves (generator only)."""
from imageanalysis_b200.propshim import PropertyNode, getNode, root  # noqa: F401
