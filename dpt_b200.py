"""Import alias: ``import dpt_b200`` == the package in ``decision-pretrained-transformer_b200/``
(whose directory name is not a valid Python identifier).  Submodules are registered under both names
so there is exactly one instance of each."""
import importlib
import sys

_REAL = "decision-pretrained-transformer_b200"
_pkg = importlib.import_module(_REAL)
for _name, _mod in list(sys.modules.items()):
    if _name.startswith(_REAL + "."):
        sys.modules["dpt_b200" + _name[len(_REAL):]] = _mod
sys.modules["dpt_b200"] = _pkg
