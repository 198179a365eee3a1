"""Import alias: ``import se_b200`` -> the package directory whose name (mandated by the build
contract) contains hyphens and therefore cannot appear in an ``import`` statement."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_name = "sixty-years-of-frequency-domain-monaural-speech-enhancement_b200"
_pkg = importlib.import_module(_name)
# alias the package AND its already-imported submodules, so `from se_b200._lib import X` binds the
# same module objects (a second import under the alias name would duplicate classes and globals)
for _k, _m in list(sys.modules.items()):
    if _k.startswith(_name + "."):
        sys.modules[__name__ + _k[len(_name):]] = _m
sys.modules[__name__] = _pkg
