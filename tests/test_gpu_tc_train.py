"""GPU tests of the tensor-core (bf16) training building blocks: T16 tile images, the MN-major
tcgen05 weight-gradient kernel, and the bf16 forward-with-store / backward-data chain."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def bf16(x):
    return x.to(torch.bfloat16).float()


@pytest.fixture(scope='module')
def ops():
    from humannerf_b200 import ops as _ops
    return _ops


@pytest.mark.parametrize('P,F,pad', [(1, 3, 16), (300, 40, 48), (128 * 5, 256, 256), (1000, 63, 64)])
def test_tile_roundtrip(ops, P, F, pad):
    g = torch.Generator().manual_seed(P + F)
    x = torch.randn(P, F, generator=g)
    img = ops.tile_pack(x.to(DEV), pad)
    assert img.numel() == ((P + 127) // 128) * 128 * pad * 2
    back = ops.tile_unpack(img, P, F, pad).cpu()
    assert torch.equal(back, bf16(x))
    # explicit layout check on a few elements
    raw = img.cpu().numpy().view(np.uint16)
    for s, f in [(0, 0), (P - 1, F - 1), (P // 2, F // 3)]:
        off = ((s // 128) * 128 * pad * 2 + ((s % 128) // 16) * 16 * pad * 2 + (f // 8) * 256 + (s % 16) * 16 + (f % 8) * 2) // 2
        want = x[s, f].to(torch.bfloat16).view(torch.int16).item() & 0xFFFF
        assert int(raw[off]) == want, (s, f)


def test_tile_wgrad_against_matmul(ops):
    P = 1000
    g = torch.Generator().manual_seed(0)
    dz = torch.randn(P, 256, generator=g) * 0.5
    x = torch.randn(P, 256, generator=g)
    pe = torch.randn(P, 63, generator=g)
    pe_ones = torch.cat([pe, torch.ones(P, 1)], 1)                     # encoding block with its constant-1 column
    dz_img, x_img, pe_img = ops.tile_pack(dz.to(DEV), 256), ops.tile_pack(x.to(DEV), 256), ops.tile_pack(pe_ones.to(DEV), 64)
    dz3 = torch.randn(P, 3, generator=g)
    dz3_img = ops.tile_pack(dz3.to(DEV), 16)
    dzb, xb, peb, dz3b = bf16(dz), bf16(x), bf16(pe_ones), bf16(dz3)

    dw_full = torch.zeros(256, 319, device=DEV)            # skip layer: [pe 63 | h 256]
    db_full = torch.zeros(256, device=DEV)
    dw_bias_only = torch.zeros(256, 16, device=DEV)
    db2 = torch.zeros(256, device=DEV)
    dw_out = torch.zeros(3, 256, device=DEV)
    dw_half = torch.zeros(128, 128, device=DEV)
    jobs = [
        dict(dz=dz_img, x=x_img, dw=dw_full, f_dz=256, f_x=256, m=256, x_f0=0, n=256, n_valid=256, dw_ld=319, dw_col0=63),
        dict(dz=dz_img, x=pe_img, dw=dw_full, db=db_full, f_dz=256, f_x=64, m=256, x_f0=0, n=64, n_valid=63, dw_ld=319,
             dw_col0=0, bias_col=63),
        # bias-only job: last 16 columns of the encoding block, nothing written to dW
        dict(dz=dz_img, x=pe_img, dw=dw_bias_only, db=db2, f_dz=256, f_x=64, m=256, x_f0=48, n=16, n_valid=0, dw_ld=16,
             dw_col0=0, bias_col=15),
        dict(dz=dz3_img, x=x_img, dw=dw_out, f_dz=16, f_x=256, m=3, x_f0=0, n=256, n_valid=256, dw_ld=256, dw_col0=0),
        # one M tile, window in the middle of X
        dict(dz=dz_img, x=x_img, dw=dw_half, f_dz=256, f_x=256, m=128, x_f0=64, n=128, n_valid=128, dw_ld=128, dw_col0=0),
    ]
    ops.tile_wgrad(jobs, P)
    torch.cuda.synchronize()

    def check(name, got, want, tol=2e-3):
        err = (got.cpu() - want).abs().max().item() / (want.abs().max().item() + 1e-12)
        print(f'[parity] tile_wgrad {name}: max-norm rel err {err:.3e}')
        assert err < tol, (name, err)
    check('dW[:, 63:]', dw_full[:, 63:], dzb.t() @ xb)
    check('dW[:, :63]', dw_full[:, :63], dzb.t() @ peb[:, :63])
    check('db (ones column)', db_full, dzb.sum(0))
    check('db (bias-only job)', db2, dzb.sum(0))
    assert float(dw_bias_only.abs().max()) == 0.0
    check('dW out (m=3)', dw_out, dz3b.t() @ xb)
    check('dW window', dw_half, dzb[:, :128].t() @ xb[:, 64:192])
    # accumulation: running the same jobs again doubles the result
    ops.tile_wgrad(jobs[:1], P)
    check('accumulate', dw_full[:, 63:], 2 * (dzb.t() @ xb))


# --------------------------------------------------------------------------------------
# bf16 tensor-core training path: forward with stored activations + backward chains + wgrad
# --------------------------------------------------------------------------------------
def _weights(sd, dev, requires_grad=False):
    nr = [sd[f'non_rigid_mlp.module.block_mlps.{i}.weight'] for i in range(0, 13, 2)]
    nrb = [sd[f'non_rigid_mlp.module.block_mlps.{i}.bias'] for i in range(0, 13, 2)]
    cn = [sd[f'cnl_mlp.module.pts_linears.{i}.weight'] for i in range(0, 15, 2)] + [sd['cnl_mlp.module.output_linear.0.weight']]
    cnb = [sd[f'cnl_mlp.module.pts_linears.{i}.bias'] for i in range(0, 15, 2)] + [sd['cnl_mlp.module.output_linear.0.bias']]
    mk = lambda ts: [t.clone().to(dev).requires_grad_(requires_grad) for t in ts]
    return mk(nr), mk(nrb), mk(cn), mk(cnb)


NAMES = ([f'non_rigid_mlp.module.block_mlps.{i}.weight' for i in range(0, 13, 2)]
         + [f'non_rigid_mlp.module.block_mlps.{i}.bias' for i in range(0, 13, 2)]
         + [f'cnl_mlp.module.pts_linears.{i}.weight' for i in range(0, 15, 2)] + ['cnl_mlp.module.output_linear.0.weight']
         + [f'cnl_mlp.module.pts_linears.{i}.bias' for i in range(0, 15, 2)] + ['cnl_mlp.module.output_linear.0.bias'])


@pytest.mark.parametrize('P,ignore', [(700, False), (128 * 7 + 5, True), (4099, False)])
def test_field_bf16_backward(ops, P, ignore):
    """Gradients of the bf16 path.

    Checker 1 (0.12 relative L2 per tensor; measured 1.5e-2 ... 8e-2): `tests/bf16_emu.py`, the oracle's
    graph with the kernel's rounding points (bf16 weights / activations / pre-activation gradients, fp32
    accumulate).  The residual is `__sincosf` vs `torch.sin` moving ~4 % of the encoding entries by one
    bf16 ulp, which flips a few ReLU masks.
    Checker 2: the fp32 oracle - the kernel must be as close to it as the emulation is (within 15 % +
    0.02), and below 0.6 absolute.  The distance between the two CHECKERS is 0.13 (last
    hidden layer) to 0.43 (first non-rigid layers) on this fixture: a 2 % forward error of bf16
    activations flips ~2 % of the ReLU masks per layer, and every flipped unit changes its gradient by
    100 % - the bf16 gradient is the exact gradient of the bf16 forward function, not a noisy copy of
    the fp32 one.  Measured values are printed."""
    import os, sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import bf16_emu as emu
    from humannerf_b200 import _lib, synthetic as syn
    from oracle import humannerf_oracle as orc
    sd = syn.fixture_state_dict(syn.REFERENCE_PARAM_SHAPES, dense=True)
    frame = syn.make_frame(0)
    gen = torch.Generator().manual_seed(P)
    x0 = torch.rand(P, 3, generator=gen) * 1.6 - 0.8
    cond = torch.from_numpy(frame['dst_posevec'])
    hann = orc.hann_weights(30000, 6, 10000, 50000)
    g_raw, g_xyz, g_off = torch.randn(P, 4, generator=gen), torch.randn(P, 3, generator=gen), torch.randn(P, 3, generator=gen)

    def cpu_grads(emulate):
        sdo = {k: v.clone().requires_grad_(True) for k, v in sd.items() if k in NAMES}
        x = x0.clone().requires_grad_()
        if emulate:
            raw_o, xyz_o, off_o = emu.field(sdo, x, cond, hann, ignore_non_rigid=ignore)
        elif ignore:
            xyz_o, off_o = x, torch.zeros_like(x)
            raw_o = orc.canonical_mlp(sdo, orc.fourier_pe(xyz_o, 10))
        else:
            xyz_o, off_o = orc.non_rigid_mlp(sdo, orc.fourier_pe(x, 6, hann, False), x, cond[None])
            raw_o = orc.canonical_mlp(sdo, orc.fourier_pe(xyz_o, 10))
        ((raw_o * g_raw).sum() + (xyz_o * g_xyz).sum() + (off_o * g_off).sum()).backward()
        return x.grad, {k: v.grad for k, v in sdo.items()}, raw_o.detach()
    gx32, gw32, raw32 = cpu_grads(False)
    gx16, gw16, raw16 = cpu_grads(True)
    # CUDA bf16
    spec = ops.FieldSpec()
    w = _weights(sd, DEV, requires_grad=True)
    packed = ops.pack_field_bf16(spec, *w, cond.to(DEV))
    xc = x0.to(DEV).requires_grad_()
    raw, xyz, off = ops.canonical_field(spec, xc, cond.to(DEV), hann.to(DEV), *w, ignore_non_rigid=ignore,
                                        precision=_lib.PRECISION_BF16, packed=packed)
    ((raw * g_raw.to(DEV)).sum() + (xyz * g_xyz.to(DEV)).sum() + (off * g_off.to(DEV)).sum()).backward()
    torch.cuda.synchronize()

    def rel_l2(a, b):
        return float((a.detach().double().cpu() - b.double()).norm() / (b.double().norm() + 1e-20))
    print(f'[parity] field16.raw[P={P},ignore={ignore}]: rel L2 vs bf16-emu {rel_l2(raw, raw16):.3e}  vs fp32 {rel_l2(raw, raw32):.3e}')
    bad = []

    def check(name, got, want16, want32):
        e16, e32 = rel_l2(got, want16), rel_l2(got, want32)
        print(f'[parity] field16.grad {name}[P={P},ignore={ignore}]: rel L2 vs bf16-emu {e16:.3e}  vs fp32 {e32:.3e}  '
              f'(emu vs fp32 {rel_l2(want16, want32):.3e})')
        emu32 = rel_l2(want16, want32)
        if not (e16 < 1.2e-1 and e32 < 0.6 and e32 < 1.15 * emu32 + 0.02):
            bad.append((name, e16, e32, emu32))
    check('d_x', xc.grad, gx16, gx32)
    for n, p in zip(NAMES, w[0] + w[1] + w[2] + w[3]):
        if ignore and n.startswith('non_rigid'):
            assert p.grad is None or float(p.grad.abs().max()) == 0.0
            continue
        check(n, p.grad, gw16[n], gw32[n])
    assert not bad, bad


def test_network_train_bf16(ops, monkeypatch):
    """End to end: Network in train mode on the bf16 path.  Outputs against the reference's golden
    render (2e-2 abs).  Gradients of all 55 parameter tensors (i) against the oracle run with the
    bf16-emulated MLPs of tests/bf16_emu.py - stated tolerance 0.15 relative L2 - and (ii) against the
    reference's fp32 golden gradients: norms within 30 %, relative L2 < 0.6 (see
    test_field_bf16_backward for why bf16 and fp32 gradients of a ReLU MLP differ by tens of percent)."""
    import os, sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import bf16_emu as emu
    from humannerf_b200 import synthetic as syn
    from humannerf_b200.config import make_cfg
    from humannerf_b200.network import Network
    from oracle import humannerf_oracle as orc
    GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
    g = np.load(os.path.join(GOLD, 'network_e2e.npz'))
    st = np.load(os.path.join(GOLD, 'stages.npz'))
    Tc = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    T = lambda a: Tc(a).to(DEV)
    cfg = make_cfg()
    cfg.b200.mlp_precision = 'bf16'
    cfg.perturb = 1.0
    sd = syn.fixture_state_dict(syn.REFERENCE_PARAM_SHAPES, dense=True)
    net = Network()
    net.load_state_dict(sd)
    net = net.to(DEV).train()
    frame = syn.make_frame(0)
    fr = {k: T(v) for k, v in frame.items()}
    data = dict(rays=T(st['rays']), near=T(st['near']), far=T(st['far']), dst_Rs=fr['dst_Rs'], dst_Ts=fr['dst_Ts'],
                cnl_gtfms=fr['cnl_gtfms'], motion_weights_priors=fr['motion_weights_priors'], dst_posevec=fr['dst_posevec'],
                cnl_bbox_min_xyz=fr['cnl_bbox_min_xyz'], cnl_bbox_max_xyz=fr['cnl_bbox_max_xyz'],
                cnl_bbox_scale_xyz=fr['cnl_bbox_scale_xyz'], bgcolor=T(g['bgcolor']))
    out = net(**data, iter_val=torch.full((1,), 30000, device=DEV), t_rand=T(st['t_rand']))
    for k in ('rgb', 'alpha'):
        err = float((out[k].detach().cpu() - torch.from_numpy(g['train_' + k])).abs().max())
        print(f'[parity] bf16 train_{k}: max abs err {err:.3e}')
        assert err < 2e-2
    loss = (out['rgb'] * T(g['train_loss_w'])).sum() + 0.3 * out['alpha'].sum() + 0.1 * out['depth'].sum()
    net.zero_grad()
    loss.backward()
    grads = {k: p.grad for k, p in net.named_parameters()}

    # the same step through the oracle with bf16-emulated MLPs (CPU)
    monkeypatch.setattr(orc, 'non_rigid_mlp', emu.non_rigid_mlp)
    monkeypatch.setattr(orc, 'canonical_mlp', emu.canonical_mlp)
    sde = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    frc = {k: Tc(v) for k, v in frame.items()}
    frc['bgcolor'] = Tc(g['bgcolor'])
    oe = orc.network_forward(sde, Tc(st['rays']), Tc(st['near']), Tc(st['far']), frc, torch.full((1,), 30000),
                             t_rand=Tc(st['t_rand']))
    ((oe['rgb'] * Tc(g['train_loss_w'])).sum() + 0.3 * oe['alpha'].sum() + 0.1 * oe['depth'].sum()).backward()
    bad = []
    for n in sorted(grads):
        ref = sde[n].grad.double()
        e = float((grads[n].double().cpu() - ref).norm() / (ref.norm() + 1e-30))
        print(f'[parity] bf16 e2e grad vs bf16-emu oracle {n}: rel L2 {e:.3e}')
        if not e < 0.15:
            bad.append((n, 'emu', e))
    for n, ref_l2 in zip([str(x) for x in g['grad_names']], g['grad_l2']):
        ours = float(grads[n].double().norm())
        err = abs(ours - ref_l2) / max(ref_l2, 1e-12)
        print(f'[parity] bf16 grad L2 norm vs fp32 reference {n}: ours {ours:.5g} ref {ref_l2:.5g} rel {err:.2e}')
        if err > 0.3:
            bad.append((n, 'norm', err))
    for key in g.files:
        if key.startswith('grad/'):
            ref = torch.from_numpy(g[key]).double()
            e = float((grads[key[5:]].double().cpu() - ref).norm() / ref.norm())
            print(f'[parity] bf16 {key} vs fp32 reference: rel L2 err {e:.3e}')
            if e > 0.6:
                bad.append((key, 'fp32', e))
    assert not bad, bad
