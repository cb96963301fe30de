"""uhsdr_b200 -- B200-native batched receiver/transmitter DSP engine behind UHSDR's block contract.

Host-side mirror of the reference's block-path interface.  All arithmetic runs in the CUDA
library uhsdr_b200/csrc/libuhsdr_b200.so (C ABI in include/uhsdr_b200.h); there is no CPU fallback.
"""
from .config import ChanCfg, ChanStatus, default_cfg  # noqa: F401
