"""Volume decoder alone (forward + backward) on cuda:0: for ncu launch lists and A/B timing.  usage: time_decoder.py [hn|cudnn] [tf32|fp32]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from humannerf_b200 import synthetic as syn                          # noqa: E402
from humannerf_b200.config import make_cfg                           # noqa: E402
from humannerf_b200.mweight_vol_decoders.deconv_vol_decoder import MotionWeightVolumeDecoder   # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else 'hn'
tf32 = (sys.argv[2] if len(sys.argv) > 2 else 'tf32') == 'tf32'
cfg = make_cfg()
cfg.b200.decoder = mode
sd = syn.fixture_state_dict(syn.REFERENCE_PARAM_SHAPES, dense=True)
m = MotionWeightVolumeDecoder(256, 32, 24)
m.load_state_dict({k[len('mweight_vol_decoder.'):]: v for k, v in sd.items() if k.startswith('mweight_vol_decoder.')})
m = m.to('cuda:0')
m.allow_tf32 = tf32
prior = torch.from_numpy(syn.make_frame(0)['motion_weights_priors']).to('cuda:0')[None]
g = torch.randn(1, 25, 32, 32, 32, device='cuda:0')
for rep in range(4):
    m.zero_grad(set_to_none=True)
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    vol = m(motion_weights_priors=prior)
    e1.record()
    (vol * g).sum().backward()
    e2.record()
    torch.cuda.synchronize()
    print(f'{mode} {"tf32" if tf32 else "fp32"}: forward {e0.elapsed_time(e1):.3f} ms, backward {e1.elapsed_time(e2):.3f} ms')
