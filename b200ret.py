"""Import shim: the package directory name
`optimized-sparse-retrieval-for-high-performance-rag-pipelines_b200` is not a Python identifier, so
`import b200ret` loads it under this name (relative imports inside the package keep working)."""
import importlib.util as _u
import os as _os
import sys as _sys

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)),
                         "optimized-sparse-retrieval-for-high-performance-rag-pipelines_b200")
_spec = _u.spec_from_file_location("b200ret", _os.path.join(_PKG_DIR, "__init__.py"),
                                   submodule_search_locations=[_PKG_DIR])
_mod = _u.module_from_spec(_spec)
_sys.modules["b200ret"] = _mod
_spec.loader.exec_module(_mod)
