"""Import shim: `import asph_b200` == the package in ./adaptive-sph_b200/ (a hyphen is not importable by name).
As a script it is the reference's command line, headless:  python asph_b200.py run <config> <scene> [-s secs] ..."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("adaptive-sph_b200")
if __name__ == "__main__":
    sys.exit(importlib.import_module("adaptive-sph_b200.cli").main())
sys.modules[__name__] = _pkg
