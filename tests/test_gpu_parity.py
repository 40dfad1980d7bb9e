"""GPU parity tests: every CUDA entry point (through the C ABI / ops wrappers) against the CPU
oracle and the committed golden vectors of the reference.

Stated tolerances (north_star): ray / sample indexing bit-exact; rgb / alpha / depth 1e-4 abs on
the fp32 path and 2e-2 abs on the bf16 tensor-core path; gradients 5e-3 relative (max-norm) for
MLP / volume parameters and 2e-2 for the pose decoder (ill-conditioned through
x_skel = sum(w x) / clamp(sum w, 1e-4); two fp32 CPU evaluations already differ by 4e-3)."""
import os

import numpy as np
import pytest
import torch

from humannerf_b200 import synthetic as syn
from oracle import humannerf_oracle as orc

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
DEV = 'cuda:0'


def T(x, dev='cpu'):
    return torch.from_numpy(np.ascontiguousarray(x)).to(dev)


def gold(name):
    return np.load(os.path.join(GOLD, name), allow_pickle=False)


def report(name, got, want, atol, rtol=0.0):
    got = got.detach().double().cpu().numpy() if torch.is_tensor(got) else np.asarray(got, dtype=np.float64)
    want = want.detach().double().cpu().numpy() if torch.is_tensor(want) else np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, f'{name}: shape {got.shape} vs {want.shape}'
    if got.size == 0:
        return
    err = np.abs(got - want)
    lim = atol + rtol * np.abs(want)
    worst = int(np.argmax(err - lim))
    msg = (f'{name}: max abs err {err.max():.3e} (limit {atol:.1e} + {rtol:.1e}*|ref|), mean {err.mean():.3e}, '
           f'worst idx {np.unravel_index(worst, got.shape)} got {got.flat[worst]:.6g} want {want.flat[worst]:.6g}, '
           f'nan(got)={int(np.isnan(got).sum())}')
    print('[parity]', msg)
    assert np.all(err <= lim) and not np.isnan(got).any(), msg


def rel_max(name, got, want, rel, outlier=10.0):
    """Gradient check: relative L2 error <= rel and max-norm relative error <= outlier * rel.  The
    max-norm bound is looser because a ReLU whose pre-activation is within rounding of zero may
    switch between two fp32 evaluations, which moves single entries by far more than rounding."""
    got = got.detach().double().cpu().numpy()
    want = want.detach().double().cpu().numpy() if torch.is_tensor(want) else np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, f'{name}: shape {got.shape} vs {want.shape}'
    scale = np.abs(want).max() + 1e-20
    err = np.abs(got - want).max() / scale
    l2 = np.linalg.norm(got - want) / (np.linalg.norm(want) + 1e-20)
    print(f'[parity] {name}: rel L2 err {l2:.3e} (limit {rel:.1e}), max-norm rel err {err:.3e} '
          f'(limit {outlier * rel:.1e}), |ref|max {scale:.3e}')
    assert l2 <= rel and err <= outlier * rel and not np.isnan(got).any(), f'{name}: L2 {l2:.3e} max {err:.3e} > {rel:.1e}'


@pytest.fixture(scope='module')
def ops():
    from humannerf_b200 import ops as _ops
    return _ops


@pytest.fixture(scope='module')
def frame():
    return syn.make_frame(0)


@pytest.fixture(scope='module')
def sd():
    return syn.fixture_state_dict(syn.REFERENCE_PARAM_SHAPES, dense=True)


# --------------------------------------------------------------------------------------
# rays
# --------------------------------------------------------------------------------------
def test_gen_rays_and_bbox(ops):
    g = gold('rays_bbox.npz')
    H, W = int(g['H']), int(g['W'])
    o, d = ops.gen_rays(H, W, g['K'], g['E'], DEV)
    report('gen_rays.o', o, g['rays_o'], 1e-6)
    report('gen_rays.d', d, g['rays_d_in'], 1e-12)          # float64 camera; K^-1 / dot order differ from LAPACK/BLAS
    # a2 on identical float64 inputs: bit exact
    ro, rd = T(g['rays_o'], DEV), T(g['rays_d_in'], DEV).clone()
    near, far, mask = ops.rays_bbox(g['bounds'], ro, rd)
    assert np.array_equal(mask.cpu().numpy().astype(bool), g['mask'])
    assert np.array_equal(rd.cpu().numpy(), g['rays_d_out'])
    m = g['mask']
    assert np.array_equal(near.cpu().numpy()[m], g['near']), np.abs(near.cpu().numpy()[m] - g['near']).max()
    assert np.array_equal(far.cpu().numpy()[m], g['far'])
    # hand-made edge cases: axis aligned, tiny components (patched to 1e-5), corner grazers, inside origin
    xo, xd = T(g['x_o'], DEV), T(g['x_d_in'], DEV).clone()
    xn, xf, xm = ops.rays_bbox(g['bounds'], xo, xd)
    assert np.array_equal(xm.cpu().numpy().astype(bool), g['x_mask'])
    assert np.array_equal(xd.cpu().numpy(), g['x_d_out'])
    assert np.array_equal(xn.cpu().numpy()[g['x_mask']], g['x_near'])
    assert np.array_equal(xf.cpu().numpy()[g['x_mask']], g['x_far'])
    # compaction keeps pixel order and casts to float32
    co, cd, cn, cf, idx = ops.compact_rays(mask, ro, rd, near, far)
    assert np.array_equal(idx.cpu().numpy(), np.where(m)[0])
    assert np.array_equal(cn.cpu().numpy()[:, 0], g['near'].astype(np.float32))
    assert np.array_equal(cf.cpu().numpy()[:, 0], g['far'].astype(np.float32))
    assert np.array_equal(cd.cpu().numpy(), g['rays_d_out'][m].astype(np.float32))
    # empty input
    e = torch.zeros(0, 3, dtype=torch.float64, device=DEV)
    n0, f0, m0 = ops.rays_bbox(g['bounds'], e, e.clone())
    assert m0.numel() == 0


def test_compaction_large(ops):
    gen = torch.Generator().manual_seed(3)
    n = 1920 * 1080
    mask = (torch.rand(n, generator=gen) < 0.37).to(torch.uint8)
    vals = torch.arange(n, dtype=torch.float64)
    o = torch.stack([vals, vals + 0.5, -vals], 1)
    co, cd, cn, cf, idx = ops.compact_rays(mask.to(DEV), o.to(DEV), o.to(DEV), vals.to(DEV), (vals * 2).to(DEV))
    want = torch.nonzero(mask)[:, 0]
    assert torch.equal(idx.cpu().long(), want)
    assert torch.equal(cn.cpu()[:, 0], vals[want].float())
    assert torch.equal(co.cpu(), o[want].float())


def test_sample_z_bit_exact(ops):
    g = gold('stages.npz')
    near, far = T(g['near'], DEV), T(g['far'], DEV)
    z0 = ops.sample_z(near, far, 128)
    assert torch.equal(z0.cpu(), T(g['z_plain']))
    z1 = ops.sample_z(near, far, 128, T(g['t_rand'], DEV))
    assert torch.equal(z1.cpu(), T(g['z_jitter']))
    # other sample counts against the oracle, incl. ragged S
    for S in (1, 7, 64, 256):
        tr = torch.rand(near.shape[0], S, generator=torch.Generator().manual_seed(S))
        want0 = orc.z_samples(near.cpu(), far.cpu(), S)
        assert torch.equal(ops.sample_z(near, far, S).cpu(), want0)
        if S > 1:
            assert torch.equal(ops.sample_z(near, far, S, tr.to(DEV)).cpu(), orc.z_samples(near.cpu(), far.cpu(), S, tr))


# --------------------------------------------------------------------------------------
# motion field
# --------------------------------------------------------------------------------------
def _mf_inputs(frame):
    g = gold('stages.npz')
    rays = T(g['rays'])
    o, d = rays[0].float(), rays[1].float()
    z = T(g['z_jitter'])
    return g, o, d, z


def test_motion_field_forward_golden(ops, frame):
    g, o, d, z = _mf_inputs(frame)
    Rs, Ts, vol = T(g['motion_scale_Rs'], DEV), T(g['motion_Ts'], DEV), T(g['vol'], DEV)
    bmin, bsc = T(frame['cnl_bbox_min_xyz'], DEV), T(frame['cnl_bbox_scale_xyz'], DEV)
    x, m, w = ops.motion_field(Rs, Ts, vol[:24].contiguous(), bmin, bsc, rays_o=o.to(DEV), rays_d=d.to(DEV), z=z.to(DEV))
    report('motion_field.weights', w, g['bw'], 2e-6)
    report('motion_field.mask', m, g['fg_mask'], 5e-6)
    fg = g['fg_mask'][..., 0] > 1e-3
    report('motion_field.x_skel[fg]', x.cpu()[torch.from_numpy(fg)], g['x_skel'][fg], 1e-5)
    report('motion_field.x_skel', x, g['x_skel'], 2e-4)     # near-empty samples amplify rounding by 1e4
    # explicit-points mode gives the same answer, and weights can be skipped
    pts = (o[:, None, :] + d[:, None, :] * z[..., None]).to(DEV)
    x2, m2, w2 = ops.motion_field(Rs, Ts, vol[:24].contiguous(), bmin, bsc, pts=pts, want_weights=False)
    assert w2 is None
    assert torch.equal(x2, x) and torch.equal(m2, m)


def test_motion_field_known_answers(ops):
    V, J = 8, 3
    gen = torch.Generator().manual_seed(0)
    vol = torch.rand(J, V, V, V, generator=gen)
    Rs = torch.eye(3).expand(J, 3, 3).contiguous()
    Ts = torch.zeros(J, 3)
    bmin, bsc = torch.tensor([-1.0, -1.0, -1.0]), torch.tensor([1.0, 1.0, 1.0])   # g == pos
    # voxel centres return the voxel; a full cell outside returns exactly zero
    k = torch.tensor([[1.0, 2.0, 3.0], [0.0, 0.0, 0.0], [7.0, 7.0, 7.0]])
    pts = (2 * k / (V - 1) - 1)[None]                       # (1, 3, 3) as (rays=1, samples=3)
    far_out = torch.tensor([[[1.0 + 2.0 / (V - 1) + 1e-3, 0.0, 0.0], [0.0, -1.0 - 2.0 / (V - 1) - 1e-3, 0.0]]])
    x, m, w = ops.motion_field(Rs.to(DEV), Ts.to(DEV), vol.to(DEV), bmin.to(DEV), bsc.to(DEV), pts=pts.to(DEV))
    want = torch.stack([vol[:, 3, 2, 1], vol[:, 0, 0, 0], vol[:, 7, 7, 7]], 0)
    report('kat.voxel_centres', w[0], want, 2e-6)
    x, m, w = ops.motion_field(Rs.to(DEV), Ts.to(DEV), vol.to(DEV), bmin.to(DEV), bsc.to(DEV), pts=far_out.to(DEV))
    assert float(w.abs().max()) == 0.0 and float(m.abs().max()) == 0.0 and float(x.abs().max()) == 0.0
    # identity transforms: x_skel = x * W / clamp(W, 1e-4)
    p = (torch.rand(4, 16, 3, generator=gen) * 1.6 - 0.8)
    x, m, w = ops.motion_field(Rs.to(DEV), Ts.to(DEV), vol.to(DEV), bmin.to(DEV), bsc.to(DEV), pts=p.to(DEV))
    report('kat.identity', x.cpu(), p * (m.cpu() / m.cpu().clamp(min=1e-4)), 1e-6)


def test_motion_field_backward(ops, frame):
    g, o, d, z = _mf_inputs(frame)
    bmin, bsc = T(frame['cnl_bbox_min_xyz']), T(frame['cnl_bbox_scale_xyz'])
    Rs0, Ts0, vol0 = T(g['motion_scale_Rs']), T(g['motion_Ts']), T(g['vol'])[:24].contiguous()
    gen = torch.Generator().manual_seed(7)
    gx, gm, gw = torch.randn(24, 128, 3, generator=gen), torch.randn(24, 128, 1, generator=gen), torch.randn(24, 128, 24, generator=gen)
    # oracle autograd
    Rs, Ts, vol = Rs0.clone().requires_grad_(), Ts0.clone().requires_grad_(), vol0.clone().requires_grad_()
    pts = (o[:, None, :] + d[:, None, :] * z[..., None]).reshape(-1, 3)
    volp = torch.cat([vol, torch.zeros(1, 32, 32, 32)], 0)
    xs, ms, ws = orc.sample_motion_fields(pts, Rs, Ts, volp, bmin, bsc)
    loss = (xs.view(24, 128, 3) * gx).sum() + (ms.view(24, 128, 1) * gm).sum() + (ws.view(24, 128, 24) * gw).sum()
    loss.backward()
    # CUDA
    Rc, Tc, vc = Rs0.to(DEV).requires_grad_(), Ts0.to(DEV).requires_grad_(), vol0.to(DEV).requires_grad_()
    x, m, w = ops.motion_field(Rc, Tc, vc, bmin.to(DEV), bsc.to(DEV), rays_o=o.to(DEV), rays_d=d.to(DEV), z=z.to(DEV))
    ((x * gx.to(DEV)).sum() + (m * gm.to(DEV)).sum() + (w * gw.to(DEV)).sum()).backward()
    rel_max('motion_field.d_vol', vc.grad, vol.grad, 2e-3)
    rel_max('motion_field.d_Rs', Rc.grad, Rs.grad, 5e-3)
    rel_max('motion_field.d_Ts', Tc.grad, Ts.grad, 5e-3)
    # only x_skel upstream (what training uses besides the mask)
    Rc2, Tc2, vc2 = Rs0.to(DEV).requires_grad_(), Ts0.to(DEV).requires_grad_(), vol0.to(DEV).requires_grad_()
    x, m, w = ops.motion_field(Rc2, Tc2, vc2, bmin.to(DEV), bsc.to(DEV), rays_o=o.to(DEV), rays_d=d.to(DEV), z=z.to(DEV),
                               want_weights=False)
    (x * gx.to(DEV)).sum().backward()
    Rs, Ts, vol = Rs0.clone().requires_grad_(), Ts0.clone().requires_grad_(), vol0.clone().requires_grad_()
    xs, _, _ = orc.sample_motion_fields(pts, Rs, Ts, torch.cat([vol, torch.zeros(1, 32, 32, 32)], 0), bmin, bsc)
    (xs.view(24, 128, 3) * gx).sum().backward()
    rel_max('motion_field.d_vol(x only)', vc2.grad, vol.grad, 2e-3)
    rel_max('motion_field.d_Rs(x only)', Rc2.grad, Rs.grad, 5e-3)


def test_motion_field_backward_vector_scatter(ops, frame):
    """From 4096 samples on, the backward scatters the volume gradient as 16-byte vector reductions into a corner-table
    shaped scratch buffer and folds it into d_vol afterwards (hn_motion_field_bwd `grad_workspace`); below that it
    issues scalar atomics.  72 rays x 128 samples (the golden rays, three times with shifted depths) against the
    oracle's autograd, and the two scatter paths against each other through the C ABI."""
    from humannerf_b200 import _lib
    g, o, d, z = _mf_inputs(frame)
    o, d = o.repeat(3, 1), d.repeat(3, 1)
    z = torch.cat([z, z + 0.011, z - 0.007], 0)
    R, S = z.shape
    assert R * S >= 4096
    bmin, bsc = T(frame['cnl_bbox_min_xyz']), T(frame['cnl_bbox_scale_xyz'])
    Rs0, Ts0, vol0 = T(g['motion_scale_Rs']), T(g['motion_Ts']), T(g['vol'])[:24].contiguous()
    gen = torch.Generator().manual_seed(17)
    gx, gm = torch.randn(R, S, 3, generator=gen), torch.randn(R, S, 1, generator=gen)
    Rs, Ts, vol = Rs0.clone().requires_grad_(), Ts0.clone().requires_grad_(), vol0.clone().requires_grad_()
    pts = (o[:, None, :] + d[:, None, :] * z[..., None]).reshape(-1, 3)
    xs, ms, _ = orc.sample_motion_fields(pts, Rs, Ts, torch.cat([vol, torch.zeros(1, 32, 32, 32)], 0), bmin, bsc)
    ((xs.view(R, S, 3) * gx).sum() + (ms.view(R, S, 1) * gm).sum()).backward()
    Rc, Tc, vc = Rs0.to(DEV).requires_grad_(), Ts0.to(DEV).requires_grad_(), vol0.to(DEV).requires_grad_()
    x, m, _ = ops.motion_field(Rc, Tc, vc, bmin.to(DEV), bsc.to(DEV), rays_o=o.to(DEV), rays_d=d.to(DEV), z=z.to(DEV),
                               want_weights=False)
    ((x * gx.to(DEV)).sum() + (m * gm.to(DEV)).sum()).backward()
    rel_max('motion_field.d_vol (vector scatter)', vc.grad, vol.grad, 2e-3)
    rel_max('motion_field.d_Rs (vector scatter)', Rc.grad, Rs.grad, 5e-3)
    rel_max('motion_field.d_Ts (vector scatter)', Tc.grad, Ts.grad, 5e-3)
    # the scalar-atomic path on the same inputs, straight through the C ABI (no gradient workspace)
    lib, ptr = _lib.load(), _lib.ptr
    dv, dR, dT = torch.zeros_like(vc), torch.zeros_like(Rc), torch.zeros_like(Tc)
    od, dd, zd, gxd, gmd = o.to(DEV), d.to(DEV), z.to(DEV).contiguous(), gx.to(DEV).contiguous(), gm.to(DEV).contiguous()
    bmd, bsd = bmin.to(DEV), bsc.to(DEV)
    _lib.check(lib.hn_motion_field_bwd(R, S, 24, 32, None, ptr(od), ptr(dd), ptr(zd), ptr(Rc.detach()), ptr(Tc.detach()),
                                       ptr(vc.detach()), ptr(bmd), ptr(bsd), ptr(gxd), ptr(gmd), None, None, None,
                                       ptr(dv), ptr(dR), ptr(dT), None, None, 0, _lib.stream_ptr()), 'motion_field_bwd')
    torch.cuda.synchronize()
    rel_max('motion_field.d_vol (scalar vs vector)', dv, vc.grad, 1e-4)
    rel_max('motion_field.d_Rs (scalar vs vector)', dR, Rc.grad, 1e-4)


# --------------------------------------------------------------------------------------
# compositing
# --------------------------------------------------------------------------------------
def test_composite_forward_golden(ops):
    g = gold('stages.npz')
    d = T(g['rays'])[1].float().to(DEV)
    out = ops.composite(T(g['comp_raw'], DEV), T(g['comp_mask'], DEV), T(g['z_jitter'], DEV), d,
                        T(g['comp_xyz'], DEV), T(g['comp_bg'], DEV))
    names = ['comp_rgb', 'comp_acc', 'comp_depth', 'comp_w', 'comp_cnl_xyz', 'comp_cnl_rgb', 'comp_wmax', 'comp_rgbs']
    for t, n in zip(out, names):
        report('composite.' + n, t, g[n], 2e-6, 1e-5)


def test_composite_known_answers_and_shapes(ops):
    for S in (1, 5, 32, 33, 100, 128, 256, 300):
        R = 37
        gen = torch.Generator().manual_seed(S)
        z = torch.sort(torch.rand(R, S, generator=gen) * 2 + 1, dim=1)[0]
        d = torch.randn(R, 3, generator=gen)
        raw = torch.randn(R, S, 4, generator=gen) * 2
        raw[..., 3] *= 8
        mask = torch.rand(R, S, 1, generator=gen) * 1.1
        xyz = torch.randn(R, S, 3, generator=gen)
        bg = torch.tensor([10.0, 100.0, 250.0])
        want = orc.composite(raw, mask, z, d, xyz, bg)
        got = ops.composite(raw.to(DEV), mask.to(DEV), z.to(DEV), d.to(DEV), xyz.to(DEV), bg.to(DEV))
        for t, k in zip(got, ['rgb', 'alpha', 'depth', 'weights_on_rays', 'cnl_xyz', 'cnl_rgb', 'cnl_weight', 'rgb_on_rays']):
            report(f'composite[S={S}].{k}', t, want[k], 3e-6, 1e-5)
    # zero density: background colour, zero depth / acc
    R, S = 5, 128
    z = torch.linspace(1, 2, S).expand(R, S).contiguous()
    d = torch.tensor([[0.0, 0.0, 1.0]]).expand(R, 3).contiguous()
    raw = torch.zeros(R, S, 4)
    bg = torch.tensor([255.0, 0.0, 51.0])
    rgb, acc, depth, *_ = ops.composite(raw.to(DEV), torch.ones(R, S, 1, device=DEV), z.to(DEV), d.to(DEV),
                                        torch.zeros(R, S, 3, device=DEV), bg.to(DEV))
    report('kat.bg', rgb, np.tile([1.0, 0.0, 0.2], (R, 1)), 1e-7)
    assert float(acc.abs().max()) == 0 and float(depth.abs().max()) == 0
    raw[..., 3] = 2.0
    rgb, acc, *_ = ops.composite(raw.to(DEV), torch.ones(R, S, 1, device=DEV), z.to(DEV), d.to(DEV),
                                 torch.zeros(R, S, 3, device=DEV), torch.zeros(3, device=DEV))
    report('kat.acc=1', acc, np.ones(R), 1e-5)


def test_composite_backward(ops):
    for S, seed in ((128, 0), (256, 1), (40, 2)):
        R = 19
        gen = torch.Generator().manual_seed(seed)
        z = torch.sort(torch.rand(R, S, generator=gen) * 2 + 1, dim=1)[0]
        d = torch.randn(R, 3, generator=gen)
        raw0 = torch.randn(R, S, 4, generator=gen) * 2
        raw0[..., 3] *= 6
        mask0 = torch.rand(R, S, 1, generator=gen)
        xyz = torch.randn(R, S, 3, generator=gen)
        bg = torch.tensor([10.0, 100.0, 250.0])
        g_rgb, g_acc, g_dep, g_w = (torch.randn(R, 3, generator=gen), torch.randn(R, generator=gen),
                                    torch.randn(R, generator=gen), torch.randn(R, S, generator=gen))
        raw, mask = raw0.clone().requires_grad_(), mask0.clone().requires_grad_()
        o = orc.composite(raw, mask, z, d, xyz, bg)
        ((o['rgb'] * g_rgb).sum() + (o['alpha'] * g_acc).sum() + (o['depth'] * g_dep).sum()
         + (o['weights_on_rays'] * g_w).sum()).backward()
        rc, mc = raw0.to(DEV).requires_grad_(), mask0.to(DEV).requires_grad_()
        out = ops.composite(rc, mc, z.to(DEV), d.to(DEV), xyz.to(DEV), bg.to(DEV))
        ((out[0] * g_rgb.to(DEV)).sum() + (out[1] * g_acc.to(DEV)).sum() + (out[2] * g_dep.to(DEV)).sum()
         + (out[3] * g_w.to(DEV)).sum()).backward()
        report(f'composite.d_raw[S={S}]', rc.grad, raw.grad, 2e-5, 2e-4)
        report(f'composite.d_mask[S={S}]', mc.grad, mask.grad, 2e-5, 2e-4)
        # rgb-only upstream gradient (what the trainer produces)
        rc2, mc2 = raw0.to(DEV).requires_grad_(), mask0.to(DEV).requires_grad_()
        (ops.composite(rc2, mc2, z.to(DEV), d.to(DEV), xyz.to(DEV), bg.to(DEV))[0] * g_rgb.to(DEV)).sum().backward()
        raw, mask = raw0.clone().requires_grad_(), mask0.clone().requires_grad_()
        (orc.composite(raw, mask, z, d, xyz, bg)['rgb'] * g_rgb).sum().backward()
        report(f'composite.d_raw(rgb only)[S={S}]', rc2.grad, raw.grad, 2e-5, 2e-4)


# --------------------------------------------------------------------------------------
# canonical field (PE + non-rigid MLP + PE + canonical MLP)
# --------------------------------------------------------------------------------------
def _field_weights(sd, dev, requires_grad=False):
    nr = [sd[f'non_rigid_mlp.module.block_mlps.{i}.weight'] for i in range(0, 13, 2)]
    nrb = [sd[f'non_rigid_mlp.module.block_mlps.{i}.bias'] for i in range(0, 13, 2)]
    cn = [sd[f'cnl_mlp.module.pts_linears.{i}.weight'] for i in range(0, 15, 2)] + [sd['cnl_mlp.module.output_linear.0.weight']]
    cnb = [sd[f'cnl_mlp.module.pts_linears.{i}.bias'] for i in range(0, 15, 2)] + [sd['cnl_mlp.module.output_linear.0.bias']]
    mk = lambda ts: [t.clone().to(dev).requires_grad_(requires_grad) for t in ts]
    return mk(nr), mk(nrb), mk(cn), mk(cnb)


def _field_oracle(sd, x, cond, hann, ignore=False):
    if ignore:
        xyz, off = x, torch.zeros_like(x)
    else:
        xyz, off = orc.non_rigid_mlp(sd, orc.fourier_pe(x, 6, hann, False), x, cond[None])
    return orc.canonical_mlp(sd, orc.fourier_pe(xyz, 10)), xyz, off


@pytest.mark.parametrize('P', [1, 127, 128, 1000, 4099])
def test_field_fp32_forward(ops, sd, frame, P):
    gen = torch.Generator().manual_seed(P)
    x = torch.rand(P, 3, generator=gen) * 1.6 - 0.8
    cond = T(frame['dst_posevec'])
    hann = orc.hann_weights(30000 if P % 2 else 10000000, 6, 10000, 50000)
    spec = ops.FieldSpec()
    w = _field_weights(sd, DEV)
    for ignore in (False, True):
        raw_w, xyz_w, off_w = _field_oracle(sd, x, cond, hann, ignore)
        raw, xyz, off = ops.canonical_field(spec, x.to(DEV), cond.to(DEV), hann.to(DEV), *w, ignore_non_rigid=ignore)
        report(f'field32.offsets[P={P},ignore={ignore}]', off, off_w, 2e-6, 1e-5)
        report(f'field32.xyz[P={P},ignore={ignore}]', xyz, xyz_w, 2e-6, 1e-5)
        # the density row of the fixture is scaled x25 and the 2^9 encoding frequency amplifies the
        # 1e-7 differences in xyz; what has to hold to 1e-4 is the composited output (network tests)
        report(f'field32.raw[P={P},ignore={ignore}]', raw, raw_w, 3e-3, 5e-4)


def test_field_fp32_backward(ops, sd, frame):
    P = 700
    gen = torch.Generator().manual_seed(11)
    x0 = torch.rand(P, 3, generator=gen) * 1.6 - 0.8
    cond = T(frame['dst_posevec'])
    hann = orc.hann_weights(30000, 6, 10000, 50000)
    g_raw, g_xyz, g_off = torch.randn(P, 4, generator=gen), torch.randn(P, 3, generator=gen), torch.randn(P, 3, generator=gen)
    names = ([f'non_rigid_mlp.module.block_mlps.{i}.weight' for i in range(0, 13, 2)]
             + [f'non_rigid_mlp.module.block_mlps.{i}.bias' for i in range(0, 13, 2)]
             + [f'cnl_mlp.module.pts_linears.{i}.weight' for i in range(0, 15, 2)] + ['cnl_mlp.module.output_linear.0.weight']
             + [f'cnl_mlp.module.pts_linears.{i}.bias' for i in range(0, 15, 2)] + ['cnl_mlp.module.output_linear.0.bias'])
    for ignore in (False, True):
        sdo = {k: v.clone().requires_grad_(True) for k, v in sd.items() if k in names}
        x = x0.clone().requires_grad_()
        raw, xyz, off = _field_oracle(sdo, x, cond, hann, ignore)
        ((raw * g_raw).sum() + (xyz * g_xyz).sum() + (off * g_off).sum()).backward()
        w = _field_weights(sd, DEV, requires_grad=True)
        xc = x0.to(DEV).requires_grad_()
        rawc, xyzc, offc = ops.canonical_field(ops.FieldSpec(), xc, cond.to(DEV), hann.to(DEV), *w, ignore_non_rigid=ignore)
        ((rawc * g_raw.to(DEV)).sum() + (xyzc * g_xyz.to(DEV)).sum() + (offc * g_off.to(DEV)).sum()).backward()
        rel_max(f'field32.d_x[ignore={ignore}]', xc.grad, x.grad, 2e-3)
        flat = w[0] + w[1] + w[2] + w[3]
        for n, p in zip(names, flat):
            if ignore and n.startswith('non_rigid'):
                assert p.grad is None or float(p.grad.abs().max()) == 0.0
                continue
            rel_max(f'field32.grad {n}[ignore={ignore}]', p.grad, sdo[n].grad, 2e-3)


@pytest.mark.parametrize('P', [1, 128, 129, 5000, 128 * 300 + 17])
def test_field_bf16_forward(ops, sd, frame, P):
    """tcgen05 path: bf16 operands, fp32 accumulate.  The comparison point is the fp32 oracle; the
    tolerance is what bf16 rounding of 14 chained layers gives on these weights (sigma row x60)."""
    gen = torch.Generator().manual_seed(P)
    x = torch.rand(P, 3, generator=gen) * 1.6 - 0.8
    cond = T(frame['dst_posevec'])
    hann = orc.hann_weights(10000000, 6, 10000, 50000)
    spec = ops.FieldSpec()
    w = _field_weights(sd, DEV)
    from humannerf_b200 import _lib
    for ignore in (False, True):
        packed = ops.pack_field_bf16(spec, *w, cond.to(DEV))
        raw_w, xyz_w, off_w = _field_oracle(sd, x, cond, hann, ignore)
        raw, xyz, off = ops.canonical_field(spec, x.to(DEV), cond.to(DEV), hann.to(DEV), *w, ignore_non_rigid=ignore,
                                            precision=_lib.PRECISION_BF16, packed=packed)
        torch.cuda.synchronize()
        report(f'field16.offsets[P={P},ignore={ignore}]', off, off_w, 4e-3, 2e-2)
        report(f'field16.xyz[P={P},ignore={ignore}]', xyz, xyz_w, 4e-3, 2e-2)
        rgb_err = (torch.sigmoid(raw[:, :3].cpu()) - torch.sigmoid(raw_w[:, :3])).abs()
        print(f'[parity] field16 rgb(sigmoid) err max {rgb_err.max():.3e} mean {rgb_err.mean():.3e}; '
              f'sigma err mean {(raw[:, 3].cpu() - raw_w[:, 3]).abs().mean():.3e} '
              f'(|sigma| mean {raw_w[:, 3].abs().mean():.3e})')
        assert not torch.isnan(raw).any()
        if ignore:
            # without the non-rigid offsets the canonical input is identical: tight check
            assert float(rgb_err.mean()) < 1e-2
            assert float((raw[:, 3].cpu() - raw_w[:, 3]).abs().mean()) < 0.05 * float(raw_w[:, 3].abs().mean()) + 0.05
        else:
            assert float(rgb_err.mean()) < 3e-2


# --------------------------------------------------------------------------------------
# end to end: Network.forward vs the reference's golden outputs
# --------------------------------------------------------------------------------------
KEYS = ['rgb', 'alpha', 'depth', 'weights_on_rays', 'xyz_on_rays', 'rgb_on_rays', 'cnl_xyz', 'cnl_rgb',
        'cnl_weight', 'backward_motion_weights', 'offsets']
PER_SAMPLE = ('xyz_on_rays', 'rgb_on_rays', 'offsets')


def _network(sd, **b200):
    from humannerf_b200.config import make_cfg
    from humannerf_b200.network import Network
    cfg = make_cfg()
    for k, v in b200.items():
        cfg.b200[k] = v
    net = Network()
    net.load_state_dict(sd, strict=True)
    return net.to(DEV), cfg


def _e2e_data(g, frame):
    st = gold('stages.npz')
    fr = {k: T(v, DEV) for k, v in frame.items()}
    data = dict(rays=T(st['rays'], DEV), near=T(st['near'], DEV), far=T(st['far'], DEV), dst_Rs=fr['dst_Rs'],
                dst_Ts=fr['dst_Ts'], cnl_gtfms=fr['cnl_gtfms'], motion_weights_priors=fr['motion_weights_priors'],
                dst_posevec=fr['dst_posevec'], cnl_bbox_min_xyz=fr['cnl_bbox_min_xyz'],
                cnl_bbox_max_xyz=fr['cnl_bbox_max_xyz'], cnl_bbox_scale_xyz=fr['cnl_bbox_scale_xyz'],
                bgcolor=T(g['bgcolor'], DEV), head_id=torch.tensor(0, device=DEV),
                dir_idx=torch.zeros(1, dtype=torch.long, device=DEV), frame_id=torch.tensor(0, device=DEV),
                # junk keys the trainer also passes; must be accepted and ignored
                patch_masks=torch.zeros(1, device=DEV), target_patches=torch.zeros(1, device=DEV))
    return data, T(st['t_rand'], DEV)


def _check_outputs(out, g, prefix, atol, atol_xyz):
    fg = g[prefix + 'backward_motion_weights'].sum(-1) > 1e-2
    assert sorted(out.keys()) == sorted(KEYS)
    for k in KEYS:
        a, b = out[k].detach().cpu().numpy(), g[prefix + k]
        assert a.shape == b.shape, (k, a.shape, b.shape)
        if k in PER_SAMPLE:
            a, b = a[fg], b[fg]
        if k in ('cnl_xyz', 'cnl_rgb'):        # arg-max picks: compare where the winner is unambiguous
            sel = g[prefix + 'cnl_weight'] > 1e-3
            a, b = a[sel], b[sel]
        report(prefix + k, a, b, atol_xyz if k in ('xyz_on_rays', 'cnl_xyz') else atol, 1e-4)


def test_network_eval_fp32(sd, frame):
    g = gold('network_e2e.npz')
    net, cfg = _network(sd, mlp_precision='fp32')
    data, _ = _e2e_data(g, frame)
    cfg.perturb = 0.
    net.eval()
    with torch.no_grad():
        out = net(**data, iter_val=cfg.eval_iter)
    _check_outputs(out, g, 'eval_', 1e-4, 2e-4)
    # chunking must not change anything: 24 rays in chunks of 7
    cfg.chunk = 7
    with torch.no_grad():
        out2 = net(**data, iter_val=cfg.eval_iter)
    for k in KEYS:      # (the cuDNN volume decoder is not bit-reproducible between calls, so no torch.equal here)
        report('chunked ' + k, out2[k], out[k], 2e-5, 1e-4)
    # light-weight dict
    cfg.b200.heavy_outputs = False
    with torch.no_grad():
        out3 = net(**data, iter_val=cfg.eval_iter)
    assert 'backward_motion_weights' not in out3 and 'rgb_on_rays' not in out3
    report('light rgb', out3['rgb'], out['rgb'], 2e-5, 1e-4)


def test_network_train_fp32_with_grads(sd, frame):
    g = gold('network_e2e.npz')
    net, cfg = _network(sd, mlp_precision='fp32')
    data, t_rand = _e2e_data(g, frame)
    cfg.perturb = 1.
    net.train()
    out = net(**data, iter_val=torch.full((1,), 30000, device=DEV), t_rand=t_rand)
    _check_outputs(out, g, 'train_', 1e-4, 2e-4)
    loss = (out['rgb'] * T(g['train_loss_w'], DEV)).sum() + 0.3 * out['alpha'].sum() + 0.1 * out['depth'].sum()
    assert abs(loss.item() - float(g['train_loss'])) < 2e-3
    net.zero_grad()
    loss.backward()
    grads = {k: p.grad for k, p in net.named_parameters()}
    names = [str(n) for n in g['grad_names']]
    assert sorted(grads) == names
    print('[parity] d(out bias) ours', grads['cnl_mlp.module.output_linear.0.bias'].cpu().numpy())
    bad = []
    for n, ref_l2 in zip(names, g['grad_l2']):
        assert grads[n] is not None, n
        ours = float(grads[n].double().norm())
        rel = 2e-2 if n.startswith('pose_decoder') else 5e-3
        err = abs(ours - ref_l2) / max(ref_l2, 1e-12)
        print(f'[parity] grad L2 {n}: ours {ours:.6g} ref {ref_l2:.6g} rel {err:.2e}')
        if abs(ours - ref_l2) > rel * max(ref_l2, 1e-6) + 1e-7:
            bad.append((n, ours, float(ref_l2)))
    for key in g.files:
        if key.startswith('grad/'):
            try:
                rel_max(key, grads[key[5:]], g[key], 2e-2 if 'pose_decoder' in key else 5e-3)
            except AssertionError as e:
                bad.append((key, str(e)))
    assert not bad, bad


def test_composite_backward_on_network_tensors(ops, sd, frame):
    """Compositing backward on the tensors the network actually produces (zeros in the mask, large
    densities, the 1e10 last interval), upstream gradient = the e2e test loss."""
    g = gold('network_e2e.npz')
    st = gold('stages.npz')
    fr = {k: T(v) for k, v in frame.items()}
    fr['bgcolor'] = T(g['bgcolor'])
    with torch.no_grad():
        tr = orc.network_forward(sd, T(st['rays']), T(st['near']), T(st['far']), fr, torch.full((1,), 30000),
                                 t_rand=T(st['t_rand']), return_intermediates=True)
    raw0, mask0, z, xyz = tr['raw'], tr['mask'], tr['z_vals'], tr['xyz_on_rays']
    d = T(st['rays'])[1].float()
    lw = T(g['train_loss_w'])
    raw, mask = raw0.clone().requires_grad_(), mask0.clone().requires_grad_()
    o = orc.composite(raw, mask, z, d, xyz, fr['bgcolor'])
    ((o['rgb'] * lw).sum() + 0.3 * o['alpha'].sum() + 0.1 * o['depth'].sum()).backward()
    rc, mc = raw0.to(DEV).requires_grad_(), mask0.to(DEV).requires_grad_()
    out = ops.composite(rc, mc, z.to(DEV), d.to(DEV), xyz.to(DEV), fr['bgcolor'].to(DEV))
    ((out[0] * lw.to(DEV)).sum() + 0.3 * out[1].sum() + 0.1 * out[2].sum()).backward()
    print('[parity] sum d_raw per channel: ours', rc.grad.sum((0, 1)).cpu().numpy(), 'oracle', raw.grad.sum((0, 1)).numpy())
    report('composite(e2e).d_raw', rc.grad, raw.grad, 1e-6, 1e-3)
    report('composite(e2e).d_mask', mc.grad, mask.grad, 1e-6, 1e-3)
    # and the canonical-field backward driven by exactly this d_raw
    xs = tr['x_skel'].reshape(-1, 3)
    cond, hann = fr['dst_posevec'], orc.hann_weights(30000, 6, 10000, 50000)
    sdo = {k: v.clone().requires_grad_(True) for k, v in sd.items() if k.startswith(('cnl_mlp', 'non_rigid'))}
    x = xs.clone().requires_grad_()
    raw_o, xyz_o, off_o = _field_oracle(sdo, x, cond, hann)
    (raw_o * raw.grad.reshape(-1, 4)).sum().backward()
    w = _field_weights(sd, DEV, requires_grad=True)
    xc = xs.to(DEV).requires_grad_()
    rawc, xyzc, offc = ops.canonical_field(ops.FieldSpec(), xc, cond.to(DEV), hann.to(DEV), *w)
    (rawc * raw.grad.reshape(-1, 4).to(DEV)).sum().backward()
    flat_names = ([f'non_rigid_mlp.module.block_mlps.{i}.weight' for i in range(0, 13, 2)]
                  + [f'non_rigid_mlp.module.block_mlps.{i}.bias' for i in range(0, 13, 2)]
                  + [f'cnl_mlp.module.pts_linears.{i}.weight' for i in range(0, 15, 2)] + ['cnl_mlp.module.output_linear.0.weight']
                  + [f'cnl_mlp.module.pts_linears.{i}.bias' for i in range(0, 15, 2)] + ['cnl_mlp.module.output_linear.0.bias'])
    flat = w[0] + w[1] + w[2] + w[3]
    for n, p in zip(flat_names, flat):
        rel_max(f'field32(e2e d_raw).grad {n}', p.grad, sdo[n].grad, 5e-3)   # sparse upstream: few samples dominate
    rel_max('field32(e2e d_raw).d_x', xc.grad, x.grad, 5e-3)


def test_network_eval_bf16(sd, frame):
    g = gold('network_e2e.npz')
    net, cfg = _network(sd, mlp_precision='bf16')
    data, _ = _e2e_data(g, frame)
    cfg.perturb = 0.
    net.eval()
    with torch.no_grad():
        out = net(**data, iter_val=cfg.eval_iter)
    for k in ('rgb', 'alpha', 'depth'):
        report('bf16.eval_' + k, out[k], g['eval_' + k], 2e-2 if k != 'depth' else 6e-2)


@pytest.mark.parametrize('precision,eps', [('fp32', 1e-7), ('bf16', 1e-5)])
def test_empty_space_skipping_bound(sd, frame, precision, eps):
    """cfg.b200.cull_mask_eps (opt-in): samples whose foreground likelihood is below eps
    skip the MLPs.  Stated bound: every per-ray output within 2 * N_samples * eps of the dense render -
    2.6e-5 at eps = 1e-7, i.e. inside the 1e-4 parity bound against the reference's golden render."""
    g = gold('network_e2e.npz')
    net, cfg = _network(sd, mlp_precision=precision)
    data, _ = _e2e_data(g, frame)
    cfg.perturb = 0.
    net.eval()
    with torch.no_grad():
        dense = net(**data, iter_val=cfg.eval_iter)
        cfg.b200.cull_mask_eps = eps
        try:
            culled = net(**data, iter_val=cfg.eval_iter)
        finally:
            cfg.b200.cull_mask_eps = 0.0
    frac = net.last_cull_fraction
    print(f'[parity] culling eps={eps:g} ({precision}): {100 * frac:.1f} % of the samples skipped')
    assert 0.0 < frac < 1.0
    bound = 2 * cfg.N_samples * eps
    for k in ('rgb', 'alpha'):
        err = float((culled[k] - dense[k]).abs().max())
        print(f'[parity] culled vs dense {k}: max abs {err:.3e} (bound {bound:.1e})')
        assert err <= bound
    if precision == 'fp32':
        for k in ('rgb', 'alpha', 'depth'):
            report('culled.eval_' + k, culled[k], g['eval_' + k], 1e-4, 1e-4)
    # a culled sample keeps xyz = x_skel, offset = 0 and gets no compositing weight
    w = culled['weights_on_rays']
    assert float(w.min()) >= 0.0 and torch.isfinite(culled['xyz_on_rays']).all()


def test_full_size_properties(ops, sd, frame):
    """BASELINE-size chunk (32768 rays x 128 samples) through the whole kernel chain; checked through
    size-independent properties instead of the (too slow) oracle."""
    R, S = 32768, 128
    K, E = syn.look_at_camera(512, 512, 1250.0)
    o64, d64 = ops.gen_rays(512, 512, K, E, DEV)
    bounds = np.stack([frame['dst_bbox_min_xyz'], frame['dst_bbox_max_xyz']])
    near, far, mask = ops.rays_bbox(bounds, o64, d64)
    o, d, nr, fr_, idx = ops.compact_rays(mask, o64, d64, near, far)
    assert o.shape[0] > 150000                      # ~196k valid rays with this camera
    assert bool((idx[1:] > idx[:-1]).all())         # pixel order kept
    o, d, nr, fr_ = o[:R].contiguous(), d[:R].contiguous(), nr[:R], fr_[:R]
    # seeded jitter: the bf16-vs-fp32 maximum over 32768 rays below varies between 1.5e-2 and 2.0e-2 from draw to draw
    z = ops.sample_z(nr, fr_, S, torch.rand(R, S, device=DEV, generator=torch.Generator(device=DEV).manual_seed(0)))
    assert bool((z[:, 1:] >= z[:, :-1]).all()) and bool((z[:, :1] >= nr).all()) and bool((z[:, -1:] <= fr_).all())
    net, cfg = _network(sd, mlp_precision='fp32')
    fr = {k: T(v, DEV) for k, v in frame.items()}
    with torch.no_grad():
        Rs, Ts = net.motion_basis_computer(fr['dst_Rs'][None], fr['dst_Ts'][None], fr['cnl_gtfms'][None])
        vol = net.mweight_vol_decoder(motion_weights_priors=fr['motion_weights_priors'][None])[0]
        x, m, w = ops.motion_field(Rs[0], Ts[0], vol[:24], fr['cnl_bbox_min_xyz'], fr['cnl_bbox_scale_xyz'],
                                   rays_o=o, rays_d=d, z=z)
        assert torch.allclose(w.sum(-1, keepdim=True), m, atol=1e-5)
        assert float(w.min()) >= 0.0 and float(m.max()) < 1.5
        spec = ops.FieldSpec()
        wts = net._field_weights()
        cond = fr['dst_posevec']
        hann = torch.ones(6, device=DEV)
        raw, xyz, off = ops.canonical_field(spec, x.view(-1, 3), cond, hann, *wts)
        # sub-ranges give bit-identical results (no cross-sample coupling, chunk invariant)
        raw_b, xyz_b, off_b = ops.canonical_field(spec, x.view(-1, 3)[4096 * 128:4096 * 128 + 1000].contiguous(), cond, hann, *wts)
        assert torch.equal(raw_b, raw[4096 * 128:4096 * 128 + 1000])
        assert torch.allclose(xyz, x.view(-1, 3) + off, atol=1e-6)
        bg = torch.tensor([255.0, 128.0, 0.0], device=DEV)
        rgb, acc, depth, wts_r, *_ = ops.composite(raw.view(R, S, 4), m, z, d, xyz.view(R, S, 3), bg)
        rgb0, acc0, *_ = ops.composite(raw.view(R, S, 4), m, z, d, xyz.view(R, S, 3), torch.zeros(3, device=DEV))
        assert torch.allclose(wts_r.sum(-1), acc, atol=1e-5)
        assert torch.allclose(rgb - rgb0, (1 - acc)[:, None] * bg[None] / 255.0, atol=1e-5)
        assert torch.equal(acc, acc0) and float(acc.max()) <= 1.0 + 1e-4
        # bf16 tensor-core path at the same size: image-level agreement with fp32
        packed = ops.pack_field_bf16(spec, *wts, cond)
        from humannerf_b200 import _lib
        raw16, xyz16, off16 = ops.canonical_field(spec, x.view(-1, 3), cond, hann, *wts, precision=_lib.PRECISION_BF16,
                                                  packed=packed)
        rgb16, acc16, depth16, *_ = ops.composite(raw16.view(R, S, 4), m, z, d, xyz16.view(R, S, 3), bg)
        report('full.bf16 rgb vs fp32', rgb16, rgb, 2e-2)
        report('full.bf16 alpha vs fp32', acc16, acc, 2e-2)


def test_pose_fk_kernel_against_torch_modules(ops):
    """a5 + a6: hn_pose_fk_fwd / bwd (Rodrigues + corrected rotations + forward kinematics in one launch)
    against the torch restatement of the reference modules (nn_util.RodriguesModule / MotionBasisComputer,
    themselves pinned to the reference by the golden tests).  Forward: 2e-6 abs; gradient of the axis-angle
    corrections: 1e-4 relative."""
    from humannerf_b200.nn_util import SMPL_PARENT, MotionBasisComputer, RodriguesModule
    fr = syn.make_frame(3)
    dst_Rs, dst_Ts = T(fr['dst_Rs'], DEV), T(fr['dst_Ts'], DEV)
    cnl = T(fr['cnl_gtfms'], DEV)
    g = torch.Generator().manual_seed(5)
    for scale in (1e-5, 0.05):                         # the refiner's initial output scale, and a trained one
        rvec0 = (torch.randn(23, 3, generator=g) * scale).to(DEV)
        w_G = torch.randn(24, 4, 4, generator=g).to(DEV)
        w_G[:, 3, :] = 0.0
        # torch modules
        rv = rvec0.clone().requires_grad_()
        dR = RodriguesModule()(rv).view(23, 3, 3)
        Rs = torch.cat([dst_Rs[0:1], torch.matmul(dst_Rs[1:], dR)], dim=0)
        local = torch.zeros(24, 4, 4, device=DEV)
        local[:, :3, :3], local[:, :3, 3], local[:, 3, 3] = Rs, dst_Ts, 1.0
        glob = [local[0]]
        for i in range(1, 24):
            glob.append(torch.matmul(glob[SMPL_PARENT[i]], local[i]))
        G_ref = torch.stack(glob, 0)
        (G_ref * w_G).sum().backward()
        # kernel
        rk = rvec0.clone().requires_grad_()
        Rs_k, G_k = ops.pose_fk(rk, dst_Rs, dst_Ts, SMPL_PARENT)
        (G_k * w_G).sum().backward()
        torch.cuda.synchronize()
        report(f'pose_fk.Rs[scale={scale:g}]', Rs_k, Rs.detach(), 2e-6)
        report(f'pose_fk.G[scale={scale:g}]', G_k, G_ref.detach(), 2e-6)
        rel = float((rk.grad - rv.grad).norm() / rv.grad.norm())
        print(f'[parity] pose_fk.d_rvec[scale={scale:g}]: rel L2 {rel:.3e}')
        assert rel < 1e-4
        # and through the whole motion-basis computation
        sR, sT = MotionBasisComputer(24)(Rs.detach()[None], dst_Ts[None], cnl[None])
        inv = torch.linalg.inv_ex(G_k.detach()).inverse
        f = torch.matmul(cnl, inv)
        report(f'pose_fk.scale_Rs[scale={scale:g}]', f[:, :3, :3], sR[0], 5e-6)
        report(f'pose_fk.Ts[scale={scale:g}]', f[:, :3, 3], sT[0], 5e-6)
    # no refinement: rvec = None
    _, G0 = ops.pose_fk(None, dst_Rs, dst_Ts, SMPL_PARENT)
    assert torch.isfinite(G0).all()


# --------------------------------------------------------------------------------------
# a3: patch selection on the device
# --------------------------------------------------------------------------------------
def _sampler_inputs(ray_mask, H, W, seed):
    gen = torch.Generator().manual_seed(seed)
    R = int(ray_mask.sum())
    ro, rd = torch.randn(R, 3, generator=gen), torch.randn(R, 3, generator=gen)
    near, far = torch.rand(R, generator=gen) + 4.0, torch.rand(R, generator=gen) + 6.0
    image = torch.rand(H, W, 3, generator=gen)
    return ro, rd, near, far, image


def test_device_patch_selection_bit_exact_against_reference():
    """hn_patch_prepare + hn_select_patches against the reference's Dataset.get_patch_ray_indices
    (tests/golden/patches.npz: 6 patches of 16 x 16 on a 64 x 64 frame, the reference's logged host randoms)."""
    from humannerf_b200.data import DevicePatchSampler
    g = gold('patches.npz')
    H, W, P = int(g['H']), int(g['W']), int(g['patch_size'])
    ray_mask, subject = g['ray_mask'].astype(bool).reshape(-1), g['subject_mask'].astype(bool)
    ro, rd, near, far, image = _sampler_inputs(ray_mask, H, W, 3)
    smp = DevicePatchSampler(T(ray_mask.astype(np.uint8), DEV), T(subject.astype(np.uint8), DEV), ro.to(DEV), rd.to(DEV),
                             near.to(DEV), far.to(DEV), image.to(DEV), H, W, 6, P,
                             sample_subject_ratio=float(g['sample_subject_ratio']))
    assert smp.n_rays_frame == int(ray_mask.sum()) and smp.n_subject == int(subject.sum())
    assert smp.n_background == int((ray_mask.reshape(H, W) & ~subject).sum())
    # the candidate counts the reference drew from are the ones the device computed
    for u, n in zip(g['u_subject'], g['choice_n']):
        assert int(n) == (smp.n_subject if u < smp.ratio else smp.n_background)
    out = smp.select(T(g['u_subject'].astype(np.float64), DEV), T(g['choice_idx'].astype(np.int64), DEV))
    torch.cuda.synchronize()
    n = int(out['n_rays'])
    assert n == int(g['div_indices'][-1]) == g['select_inds'].shape[0]
    assert np.array_equal(out['select_inds'][:n].cpu().numpy(), g['select_inds'])
    assert np.array_equal(out['patch_masks'].cpu().numpy().astype(bool), g['patch_masks'])
    assert np.array_equal(out['patch_div_indices'].cpu().numpy(), g['div_indices'])
    assert np.array_equal(out['xy_min'].cpu().numpy(), g['xy_min'])
    sel = torch.from_numpy(g['select_inds'].astype(np.int64))
    assert torch.equal(out['rays'][0, :n].cpu(), ro[sel]) and torch.equal(out['rays'][1, :n].cpu(), rd[sel])
    assert torch.equal(out['rays'][2, :n].cpu(), rd[sel])
    assert torch.equal(out['near'][:n, 0].cpu(), near[sel]) and torch.equal(out['far'][:n, 0].cpu(), far[sel])
    assert np.array_equal(out['scatter_index'][:n].cpu().numpy(), np.nonzero(g['patch_masks'].reshape(-1))[0])
    if n < smp.cap:
        assert bool((out['scatter_index'][n:] == -1).all()) and float(out['rays'][:, n:].abs().max()) == 0.0
    for i, (x, y) in enumerate(g['xy_min']):
        assert torch.equal(out['target_patches'][i].cpu(), image[y:y + P, x:x + P])
    print(f'[parity] device patch selection: {n} rays in 6 patches, select_inds / masks / div / xy_min bit-exact')


@pytest.mark.parametrize('P,H,W', [(32, 512, 512), (20, 512, 512), (16, 37, 53), (16, 16, 16)])
def test_device_patch_selection_matches_host(P, H, W):
    """... and against the host implementation (data.sample_patches / make_train_batch, itself pinned to the reference by
    tests/test_host.py) at the benchmarked geometry and on ragged images, 25 steps of seeded draws each: windows
    clipped at every border, patches without any ray, the k = 0 and k = n - 1 candidates."""
    from humannerf_b200.data import DevicePatchSampler, make_train_batch
    rs = np.random.RandomState(P * 1000 + H)
    yy, xx = np.mgrid[0:H, 0:W]
    ray_mask = ((xx - W * 0.45) ** 2 / (0.4 * W) ** 2 + (yy - H * 0.5) ** 2 / (0.47 * H) ** 2 < 1.0)
    ray_mask |= rs.rand(H, W) < 0.02
    ray_mask[0, 0] = ray_mask[H - 1, W - 1] = True
    subject = ((xx - W * 0.5) ** 2 / (0.15 * W) ** 2 + (yy - H * 0.5) ** 2 / (0.35 * H) ** 2 < 1.0)
    subject[0, 0] = True                      # first candidate in row-major order
    subject[H - 1, W - 1] = True
    ro, rd, near, far, image = _sampler_inputs(ray_mask.reshape(-1), H, W, 5)
    smp = DevicePatchSampler(T(ray_mask.reshape(-1).astype(np.uint8), DEV), T(subject.astype(np.uint8), DEV), ro.to(DEV),
                             rd.to(DEV), near.to(DEV), far.to(DEV), image.to(DEV), H, W, 6, P)
    buf = smp.alloc()
    for step in range(25):
        rng_a, rng_b = np.random.RandomState(step), np.random.RandomState(step)
        draws = smp.draw(rng_a)
        if step == 0:
            draws[:, 0] = (0.1, 0)                                   # first subject pixel
            draws[:, 1] = (0.1, smp.n_subject - 1)                   # last subject pixel
            draws[:, 2] = (0.9, smp.n_background - 1)                # last background pixel
        out = smp.select(T(draws[0], DEV), T(draws[1].astype(np.int64), DEV), out=buf)
        if step == 0:
            continue                                                 # hand-made draws: checked by invariants below
        b = make_train_batch(ro.numpy(), rd.numpy(), near.numpy(), far.numpy(), ray_mask.reshape(-1), subject,
                             image.numpy(), H, W, 6, P, rng_b, pad_rays=True)
        n = int(out['n_rays'])
        assert n == int(b['n_rays'])
        assert np.array_equal(out['patch_masks'].cpu().numpy().astype(bool), b['patch_masks'])
        assert np.array_equal(out['patch_div_indices'].cpu().numpy(), b['patch_div_indices'])
        assert np.array_equal(out['scatter_index'].cpu().numpy(), b['scatter_index'])
        assert np.array_equal(out['rays'].cpu().numpy(), b['rays'])
        assert np.array_equal(out['near'].cpu().numpy(), b['near']) and np.array_equal(out['far'].cpu().numpy(), b['far'])
        assert np.array_equal(out['target_patches'].cpu().numpy(), b['target_patches'])
    # invariants of SURVEY 8c (viii) on the hand-made draws of step 0
    out = smp.select(T(np.array([0.1, 0.1, 0.9, 0.5, 0.5, 0.5]), DEV),
                     T(np.array([0, smp.n_subject - 1, smp.n_background - 1, 1, 2, 3], dtype=np.int64), DEV), out=buf)
    xy = out['xy_min'].cpu().numpy()
    assert (xy >= 0).all() and (xy[:, 0] <= W - P).all() and (xy[:, 1] <= H - P).all()
    assert tuple(xy[0]) == (0, 0) and tuple(xy[1]) == (W - P, H - P)
    assert int(out['patch_div_indices'][-1]) == int(out['n_rays']) == int(out['patch_masks'].sum())
    print(f'[parity] device patch selection vs host, P={P} on {H}x{W}: 24 seeded steps bit-exact')


# --------------------------------------------------------------------------------------
# a7: hand-written volume decoder
# --------------------------------------------------------------------------------------
def test_vol_decoder_against_reference_and_torch(ops, sd, frame):
    """hn_vol_decoder_fwd / _bwd (csrc/vol_decoder.cu, 3xTF32 products = fp32-level accuracy) against (1) the volume the
    unmodified reference decoded (tests/golden/stages.npz 'vol', 1e-5) and (2) the PyTorch module on cuDNN in fp32 (TF32
    off) for the forward and the gradients of all 14 parameter tensors under a random upstream gradient (relative L2 1e-4).
    Also: the weight gradients come back in the parameters' channels_last_3d strides, and two runs are bit-identical."""
    from humannerf_b200.config import make_cfg
    from humannerf_b200.mweight_vol_decoders.deconv_vol_decoder import MotionWeightVolumeDecoder
    cfg = make_cfg()
    prefix = 'mweight_vol_decoder.'
    dec_sd = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    mods = {}
    for mode in ('hn', 'cudnn'):
        m = MotionWeightVolumeDecoder(embedding_size=256, volume_size=32, total_bones=24)
        m.load_state_dict(dec_sd, strict=True)
        mods[mode] = m.to(DEV)
    prior = T(frame['motion_weights_priors'], DEV)[None]
    g_vol = torch.randn(1, 25, 32, 32, 32, generator=torch.Generator().manual_seed(4)).to(DEV)
    outs, grads = {}, {}
    for mode, m in mods.items():
        cfg.b200.decoder = mode
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            vol = m(motion_weights_priors=prior)
            (vol * g_vol).sum().backward()
        outs[mode] = vol.detach()
        grads[mode] = {k: p.grad.detach().clone() for k, p in m.named_parameters()}
    cfg.b200.decoder = 'hn'
    st = gold('stages.npz')
    report('vol_decoder vs reference golden', outs['hn'][0], st['vol'], 1e-5, 1e-5)
    report('vol_decoder vs cuDNN fp32', outs['hn'], outs['cudnn'], 1e-5, 1e-5)
    assert abs(float(outs['hn'].sum(1).mean()) - 1.0) < 1e-5
    for k, g in grads['hn'].items():
        ref = grads['cudnn'][k]
        assert g.shape == ref.shape and g.stride() == dict(mods['hn'].named_parameters())[k].stride(), (k, g.stride())
        rel = float((g.double() - ref.double()).norm() / (ref.double().norm() + 1e-30))
        print(f'[parity] vol_decoder grad {k}: rel L2 vs cuDNN fp32 {rel:.2e}, |g| {float(ref.norm()):.3e}')
        assert rel <= 1e-4, (k, rel)
    # plain-TF32 products (what the bf16 path uses, like cuDNN's TF32 engines): 10-bit mantissa operands
    m = mods['hn']
    m.zero_grad()
    m.allow_tf32 = True
    vol_t = m(motion_weights_priors=prior)
    (vol_t * g_vol).sum().backward()
    m.allow_tf32 = False
    report('vol_decoder TF32 products vs fp32', vol_t.detach(), outs['cudnn'], 5e-4, 5e-3)
    worst = max(float((p.grad.double() - grads['cudnn'][k].double()).norm() / grads['cudnn'][k].double().norm())
                for k, p in m.named_parameters())
    print(f'[parity] vol_decoder TF32 products: worst gradient rel L2 vs fp32 {worst:.2e}')
    assert worst <= 5e-2, worst                      # measured 2.9e-2 (10-bit mantissa operands through five layers)
    # determinism: no atomics anywhere
    m = mods['hn']
    m.zero_grad()
    vol2 = m(motion_weights_priors=prior)
    (vol2 * g_vol).sum().backward()
    assert torch.equal(vol2.detach(), outs['hn'])
    for k, p in m.named_parameters():
        assert torch.equal(p.grad, grads['hn'][k]), k


@pytest.mark.parametrize('volume_size,embedding', [(16, 256), (8, 64)])
def test_vol_decoder_other_geometries(ops, volume_size, embedding):
    """The decoder kernels on the reference module at other `cfg.mweight_volume` sizes (network_util.py:25-38 builds
    log2(V) - 1 hidden transposed convolutions): V = 16 puts the 25-channel layer on the 128-row kernel with 8^3 inputs,
    V = 8 runs every layer on the 64-row kernel.  Forward and all parameter gradients against PyTorch / cuDNN in fp32
    (TF32 off): 1e-5 abs on the softmax volume, 1e-4 relative L2 per gradient tensor; plain-TF32 mode within 5e-2."""
    from humannerf_b200.config import make_cfg
    from humannerf_b200.mweight_vol_decoders.deconv_vol_decoder import MotionWeightVolumeDecoder
    cfg = make_cfg()
    torch.manual_seed(volume_size)
    ref = MotionWeightVolumeDecoder(embedding_size=embedding, volume_size=volume_size, total_bones=24)
    with torch.no_grad():
        for p in ref.parameters():                       # the module's own init leaves the last layers tiny
            p.copy_(torch.randn_like(p) * (0.5 / max(p.shape[0] if p.dim() > 1 else 1, 1) ** 0.5))
    mods = {'cudnn': ref.to(DEV)}
    mods['hn'] = MotionWeightVolumeDecoder(embedding_size=embedding, volume_size=volume_size, total_bones=24).to(DEV)
    mods['hn'].load_state_dict(ref.state_dict())
    V = volume_size
    gen = torch.Generator().manual_seed(7)
    prior = torch.rand(1, 25, V, V, V, generator=gen).to(DEV) + 0.05
    prior = prior / prior.sum(1, keepdim=True)
    g_vol = torch.randn(1, 25, V, V, V, generator=gen).to(DEV)
    outs, grads = {}, {}
    try:
        for mode, m in mods.items():
            cfg.b200.decoder = mode
            with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
                vol = m(motion_weights_priors=prior)
                (vol * g_vol).sum().backward()
            outs[mode] = vol.detach()
            grads[mode] = {k: p.grad.detach().clone() for k, p in m.named_parameters()}
        report(f'vol_decoder V={V} vs cuDNN fp32', outs['hn'], outs['cudnn'], 1e-5, 1e-5)
        for k, g in grads['hn'].items():
            r = grads['cudnn'][k]
            rel = float((g.double() - r.double()).norm() / (r.double().norm() + 1e-30))
            assert rel <= 1e-4, (k, rel)
        cfg.b200.decoder = 'hn'
        m = mods['hn']
        m.zero_grad()
        m.allow_tf32 = True
        vol_t = m(motion_weights_priors=prior)
        (vol_t * g_vol).sum().backward()
        report(f'vol_decoder V={V} TF32 products vs fp32', vol_t.detach(), outs['cudnn'], 1e-3, 1e-2)
        worst = max(float((p.grad.double() - grads['cudnn'][k].double()).norm() / grads['cudnn'][k].double().norm())
                    for k, p in m.named_parameters())
        print(f'[parity] vol_decoder V={V} TF32 products: worst gradient rel L2 vs fp32 {worst:.2e}')
        assert worst <= 5e-2, worst
    finally:
        cfg.b200.decoder = 'auto'
