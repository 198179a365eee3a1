"""se_b200 -- B200 (sm_100a) decode path for frequency-domain speech enhancement.

Package directory name follows the build contract
(``sixty-years-of-frequency-domain-monaural-speech-enhancement_b200``); it is importable as
``se_b200`` through the alias module at the repository root.
"""
from . import _lib, ops, packing, lstm_engine, decode, shard, plan, streaming   # noqa: F401
from .crn import crn_net                   # noqa: F401
from .lstm import lstm_net                 # noqa: F401
from . import fullsubnet                   # noqa: F401
from .dccrn import DCCRN                   # noqa: F401
from .uformer import Uformer               # noqa: F401
from .dpcrn import dpcrn                   # noqa: F401
from . import gcrn                         # noqa: F401  (gcrn.Net, as GCRN/GCRN_noncprs.py names it)
from . import ctsnet                       # noqa: F401  (ctsnet.Step1_net / ctsnet.Step2_net)
from .taylor import TaylorSENet            # noqa: F401
from . import g2net                        # noqa: F401  (g2net.gaf_base)

__all__ = ["crn_net", "lstm_net", "fullsubnet", "DCCRN", "Uformer", "gcrn", "dpcrn", "ctsnet", "TaylorSENet", "g2net", "ops", "decode", "packing", "shard", "plan", "streaming"]
