"""hrfuser_b200 -- B200-native HRFuser fusion-backbone hot path.

Public surface:
  HRFuserHRFormerBased   drop-in backbone (mmdet operator API)
  HRFPN                  drop-in neck (the backbone's consumer)
  pipeline.InputPrologue device-side Normalize / Pad / DefaultFormatBundle
  backbone_cfg, WORKLOADS   config dictionaries of the shipped variants
  ops                    thin ctypes wrappers over the C-ABI (include/hrfuser_b200.h)
"""
from .backbone import HRFuserHRFormerBased, register_with_mmdet
from .configs import WORKLOADS, backbone_cfg, tiny_cfg
from .neck import HRFPN
from .neck import register_with_mmdet as _register_neck

__all__ = ['HRFuserHRFormerBased', 'HRFPN', 'register_with_mmdet', 'backbone_cfg', 'tiny_cfg',
           'WORKLOADS']
__version__ = '0.1.0'
register_with_mmdet()
_register_neck()
