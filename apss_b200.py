"""Importable alias of the package directory `all-pairs-similarity_b200/` (a hyphen is not a valid
Python identifier): `import apss_b200` gives the same module object."""
import importlib
import sys

_pkg = importlib.import_module("all-pairs-similarity_b200")
sys.modules[__name__] = _pkg
