"""dpmn_b200 -- B200-native (sm_100a) implementation of the DPMN hot path: the PGRM stack
(/root/reference/model/pgrm.py) and the Complementation Modulation Module (/root/reference/model/cmm.py).

    from dpmn_b200 import PGRM, ComplementationModulationModule

Both classes keep the reference's constructor / forward signatures and state_dict schema; all compute
runs in libdpmn_b200.so (hand-written CUDA behind the C-ABI of include/dpmn_b200.h).  There is no
CPU or PyTorch fallback.
"""
from .schema import PGRMConfig, cmm_schema, pgrm_schema  # noqa: F401


def __getattr__(name):   # lazy: `import dpmn_b200.schema` must not need torch or the CUDA library
    if name in ("PGRM", "window_attention"):
        from . import pgrm
        return getattr(pgrm, name)
    if name in ("ComplementationModulationModule", "CMM"):
        from . import cmm
        return cmm.ComplementationModulationModule
    raise AttributeError(name)
