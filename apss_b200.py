"""Importable alias of the package directory `all-pairs-similarity_b200/` (a hyphen is not a valid
Python identifier): `import apss_b200` gives the same module object."""
import importlib
import sys

_pkg = importlib.import_module("all-pairs-similarity_b200")
sys.modules[__name__] = _pkg
# the submodules the package imported are the same objects under the alias too: `from apss_b200.worker import X` must not
# load a second copy of worker.py (whose message classes would not be the ones `apss_b200.messages` exposes)
for _name, _mod in list(sys.modules.items()):
    if _name.startswith("all-pairs-similarity_b200."):
        sys.modules.setdefault(__name__ + _name[len("all-pairs-similarity_b200"):], _mod)
