"""Import alias: ``import se_b200`` -> the package directory whose name (mandated by the build
contract) contains hyphens and therefore cannot appear in an ``import`` statement."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("sixty-years-of-frequency-domain-monaural-speech-enhancement_b200")
sys.modules[__name__] = _pkg
