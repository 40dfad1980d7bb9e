"""humannerf_b200 - B200-native (sm_100a) implementation of HumanNeRF's per-ray rendering hot
path behind the reference's Network / component / config surface.  See DESIGN.md."""
from humannerf_b200.config import get_cfg, make_cfg  # noqa: F401

__all__ = ['get_cfg', 'make_cfg']
