"""Import shim: ``import svdlstm`` loads the package that lives in the (non-importable, hyphenated)
directory ``lstm-acceleration-with-singular-value-decomposition_b200/`` required by the repo layout."""
import importlib.util
import os
import sys

_pkg_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                        "lstm-acceleration-with-singular-value-decomposition_b200")
_spec = importlib.util.spec_from_file_location("svdlstm", os.path.join(_pkg_dir, "__init__.py"),
                                               submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["svdlstm"] = _mod
_spec.loader.exec_module(_mod)
