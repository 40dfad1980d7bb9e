"""Timeline of one pass of cluster 0 of the TS forward kernel (measurement build: HN_EXTRA_FLAGS=-DHN_TS_TRACE ./build.sh).
Renders 32768 rays on the bf16 path, then prints per weight chunk when the issuer had its operands and when it had issued,
and per layer when the two epilogue groups saw their accumulator half and when they handed the operand over."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from humannerf_b200 import _lib, ops, synthetic as syn          # noqa: E402

dev = 'cuda:0'
sd = syn.fixture_state_dict(syn.REFERENCE_PARAM_SHAPES, dense=True)
frame = syn.make_frame(0)
P = int(sys.argv[1]) if len(sys.argv) > 1 else 32768 * 128
train = len(sys.argv) > 2 and sys.argv[2] == 'train'
x = (torch.rand(P, 3) * 1.6 - 0.8).to(dev)
cond = torch.from_numpy(frame['dst_posevec']).to(dev)
hann = torch.ones(6, device=dev)
nr = [sd[f'non_rigid_mlp.module.block_mlps.{i}.weight'] for i in range(0, 13, 2)]
nrb = [sd[f'non_rigid_mlp.module.block_mlps.{i}.bias'] for i in range(0, 13, 2)]
cn = [sd[f'cnl_mlp.module.pts_linears.{i}.weight'] for i in range(0, 15, 2)] + [sd['cnl_mlp.module.output_linear.0.weight']]
cnb = [sd[f'cnl_mlp.module.pts_linears.{i}.bias'] for i in range(0, 15, 2)] + [sd['cnl_mlp.module.output_linear.0.bias']]
w = tuple([t.to(dev).requires_grad_(train) for t in ts] for ts in (nr, nrb, cn, cnb))
spec = ops.FieldSpec()
packed = ops.pack_field_bf16(spec, *w, cond)
for rep in range(3):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    with torch.set_grad_enabled(train):
        raw, xyz, off = ops.canonical_field(spec, x, cond, hann, *w, precision=_lib.PRECISION_BF16, packed=packed)
    ev1.record()
    torch.cuda.synchronize()
    print(f'forward of {P} samples ({"training" if train else "render"} variant): {ev0.elapsed_time(ev1):.3f} ms')
lib = _lib.load()
if not hasattr(lib, 'hn_ts_trace_read'):
    sys.exit('library built without -DHN_TS_TRACE')
buf = (ctypes.c_longlong * (3 * 128 * 2))()
lib.hn_ts_trace_read(buf)
v = np.array(list(buf), dtype=np.int64).reshape(3, 128, 2)
t0 = v[v > 0].min()
print('issuer: chunk  operands-ready  issued   (clk since the first event)')
for c in range(128):
    if v[0, c, 0]:
        print(f'  {c:3d}  {v[0, c, 0] - t0:7d}  {v[0, c, 1] - t0:7d}')
for g in (1, 2):
    print(f'epilogue group {g - 1}: slot  accumulator-seen  handed-over')
    for l in range(32):
        if v[g, l, 0]:
            print(f'  {l:3d}  {v[g, l, 0] - t0:7d}  {v[g, l, 1] - t0:7d}')
