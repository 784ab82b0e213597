"""tspn_b200 — B200-native tracklet-pair stage of TSPN.

Host side: PyTorch modules with the reference's ``lib/modeling`` ``forward()`` signatures.
Compute: hand-written sm_100a CUDA behind the C-ABI of ``include/tspn_b200.h``
(``csrc/libtspn_b200.so``), loaded with ctypes.  There is no CPU fallback: importing the
package works anywhere, calling an op without the library or without a B200 raises.
"""
__version__ = "0.1.0"

from .config import CfgNode, cfg, get_default_cfg  # noqa: E402,F401
from .list_pair import PairList  # noqa: E402,F401


def __getattr__(name):
    # heavy modules (torch.nn mirrors) load on first use
    import importlib
    lazy = {"BaseModel": "model", "RelationPredictor": "model", "RelOIPool": "model",
            "PairStage": "pipeline", "StageConfig": "pipeline", "HostBatch": "batch", "DeviceBatch": "batch"}
    if name in lazy:
        return getattr(importlib.import_module(__name__ + "." + lazy[name]), name)
    raise AttributeError(name)
