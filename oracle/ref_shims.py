"""Import the UNMODIFIED reference nn.Modules from /root/reference (build container only).

TEST INFRASTRUCTURE.  Used by ``oracle/make_golden.py`` and by the
``needs_reference`` tests to pin ``oracle.nets`` against the real thing.  The GPU
box has no ``/root/reference``; nothing that runs there may call this.

The reference model files import packages that are not installed here (librosa,
soundfile, h5py, ptflops ...) and ``config.py`` creates
directories at import (``LSTM/config.py:16-18``), so modules are imported with
empty ``sys.modules`` stubs and with the cwd in a scratch directory.
"""
from __future__ import annotations

import contextlib
import importlib
import os
import sys
import tempfile
import types

REFERENCE_ROOT = os.environ.get("SE_REFERENCE_ROOT", "/root/reference")

_STUBS = [
    "librosa", "librosa.filters", "librosa.util", "soundfile", "h5py", "ptflops",
    "ptflops.flops_counter", "torch_complex", "torch_complex.tensor", "show",
    "matplotlib", "matplotlib.pyplot", "pystoi", "pystoi.stoi", "conv_stft", "thop",
    "resampy", "pesq",
]
# names that collide between model directories and must be purged between imports
_COLLIDING = ["Backup", "config", "Step2_config", "Step1_config", "data", "istft", "misc", "loss"]


def have_reference() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "CRN"))


def _install_stubs():
    for name in _STUBS:
        if name in sys.modules:
            continue
        m = types.ModuleType(name)
        m.__dict__["__stub__"] = True
        sys.modules[name] = m
    sys.modules["ptflops"].get_model_complexity_info = lambda *a, **k: (0, 0)
    sys.modules["ptflops.flops_counter"].get_model_complexity_info = lambda *a, **k: (0, 0)
    sys.modules["show"].show_model = lambda *a, **k: None
    sys.modules["show"].show_params = lambda *a, **k: None
    sys.modules["pystoi"].stoi = lambda *a, **k: 0.0
    sys.modules["pystoi.stoi"].stoi = lambda *a, **k: 0.0
    sys.modules["torch_complex"].ComplexTensor = object
    sys.modules["torch_complex.tensor"].ComplexTensor = object
    sys.modules["librosa"].filters = sys.modules["librosa.filters"]
    sys.modules["librosa"].util = sys.modules["librosa.util"]
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    # DCCRN imports `complexnn`, which the reference does not vendor: inject the restatement
    from . import complexnn_restated
    sys.modules["complexnn"] = complexnn_restated
    sys.modules["conv_stft"].ConvSTFT = object
    sys.modules["conv_stft"].ConviSTFT = object


@contextlib.contextmanager
def _scratch_cwd():
    old = os.getcwd()
    with tempfile.TemporaryDirectory() as d:
        os.chdir(d)
        try:
            yield
        finally:
            os.chdir(old)


def import_reference(model_dir: str, module: str):
    """Import ``/root/reference/<model_dir>/<module>.py`` and return the module object."""
    if not have_reference():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    _install_stubs()
    path = os.path.join(REFERENCE_ROOT, model_dir)
    top = module.split(".")[0]
    for name in list(sys.modules):
        if name in _COLLIDING or name == top or name.startswith(top + "."):
            sys.modules.pop(name, None)
    sys.path.insert(0, path)
    try:
        with _scratch_cwd():
            mod = importlib.import_module(module)
    finally:
        sys.path.remove(path)
    # leave no colliding names behind for the next import
    for name in list(sys.modules):
        if name in _COLLIDING or name == top or name.startswith(top + "."):
            sys.modules.pop(name, None)
    return mod


def checkpoint_path(model_dir: str, name: str) -> str:
    return os.path.join(REFERENCE_ROOT, model_dir, "BEST_MODEL", name)
