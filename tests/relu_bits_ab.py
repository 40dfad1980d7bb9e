"""Helper of test_gpu_tc_train.py::test_relu_mask_images_match_activation_read (not a test module): gradients of the
bf16 field path on seeded inputs, written to an .npz.  `HN_TC_RELU_BITS` is read once per process by the library, so
the two arms of the comparison run in separate processes:  python tests/relu_bits_ab.py <P> <ignore 0|1> <out.npz>"""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def field_grads(P, ignore, dev='cuda:0'):
    from humannerf_b200 import _lib, ops, synthetic as syn
    sd = syn.fixture_state_dict(syn.REFERENCE_PARAM_SHAPES, dense=True)
    frame = syn.make_frame(0)
    gen = torch.Generator().manual_seed(1000 + P)
    x0 = torch.rand(P, 3, generator=gen) * 1.6 - 0.8
    cond = torch.from_numpy(frame['dst_posevec'])
    hann = torch.tensor([1.0, 1.0, 1.0, 0.5, 0.0, 0.0])
    g_raw, g_xyz, g_off = torch.randn(P, 4, generator=gen), torch.randn(P, 3, generator=gen), torch.randn(P, 3, generator=gen)
    nr = [sd[f'non_rigid_mlp.module.block_mlps.{i}.weight'] for i in range(0, 13, 2)]
    nrb = [sd[f'non_rigid_mlp.module.block_mlps.{i}.bias'] for i in range(0, 13, 2)]
    cn = [sd[f'cnl_mlp.module.pts_linears.{i}.weight'] for i in range(0, 15, 2)] + [sd['cnl_mlp.module.output_linear.0.weight']]
    cnb = [sd[f'cnl_mlp.module.pts_linears.{i}.bias'] for i in range(0, 15, 2)] + [sd['cnl_mlp.module.output_linear.0.bias']]
    mk = lambda ts: [t.clone().to(dev).requires_grad_(True) for t in ts]
    w = (mk(nr), mk(nrb), mk(cn), mk(cnb))
    spec = ops.FieldSpec()
    packed = ops.pack_field_bf16(spec, *w, cond.to(dev))
    xc = x0.to(dev).requires_grad_()
    raw, xyz, off = ops.canonical_field(spec, xc, cond.to(dev), hann.to(dev), *w, ignore_non_rigid=bool(ignore),
                                        precision=_lib.PRECISION_BF16, packed=packed)
    ((raw * g_raw.to(dev)).sum() + (xyz * g_xyz.to(dev)).sum() + (off * g_off.to(dev)).sum()).backward()
    torch.cuda.synchronize()
    out = {'raw': raw.detach().cpu().numpy(), 'd_x': xc.grad.cpu().numpy()}
    for gi, group in enumerate(w):
        for li, p in enumerate(group):
            if p.grad is not None:
                out[f'g{gi}_{li}'] = p.grad.cpu().numpy()
    return out


if __name__ == '__main__':
    np.savez(sys.argv[3], **field_grads(int(sys.argv[1]), int(sys.argv[2])))
