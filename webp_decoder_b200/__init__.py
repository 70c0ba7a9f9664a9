"""Import shim: the product package lives in the directory `webp-decoder_b200/` (a name Python's import
statement cannot spell). `import webp_decoder_b200` loads it from there."""
import importlib.util
import sys
from pathlib import Path

_pkg_dir = Path(__file__).resolve().parent.parent / "webp-decoder_b200"
_spec = importlib.util.spec_from_file_location(__name__, _pkg_dir / "__init__.py", submodule_search_locations=[str(_pkg_dir)])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
