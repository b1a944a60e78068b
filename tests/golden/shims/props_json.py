"""Shim for the reference's `import props_json` (generator only)."""
from imageanalysis_b200.propshim import load, save  # noqa: F401


def overlay(filename, node):
    return load(filename, node)
