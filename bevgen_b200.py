"""Import alias: ``import bevgen_b200`` loads the package directory
``visiontransformer-intention-prediction_b200/`` (whose name is not a Python identifier)."""
import importlib.util as _ilu
import pathlib as _pl
import sys as _sys

_dir = _pl.Path(__file__).resolve().parent / "visiontransformer-intention-prediction_b200"
_spec = _ilu.spec_from_file_location(__name__, _dir / "__init__.py", submodule_search_locations=[str(_dir)])
_mod = _ilu.module_from_spec(_spec)
_sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
