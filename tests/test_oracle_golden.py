"""The CPU oracle (oracle/humannerf_oracle.py) replayed against golden vectors
that were produced by the unmodified reference (tests/golden/make_golden.py).
This is what pins the oracle; the GPU tests then compare CUDA against it."""
import json
import os

import numpy as np
import pytest
import torch

from humannerf_b200 import synthetic as syn
from oracle import humannerf_oracle as orc


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def T(x):
    return torch.from_numpy(np.ascontiguousarray(x))


def test_synthetic_generators_match_reference_helpers(golden_dir):
    g = _load(golden_dir, 'synthetic_inputs.npz')
    fr = syn.make_frame(0)
    np.testing.assert_array_equal(fr['dst_Rs'], g['dst_Rs'])
    np.testing.assert_array_equal(fr['dst_Ts'], g['dst_Ts'])
    np.testing.assert_array_equal(fr['cnl_gtfms'], g['cnl_gtfms'])
    pr = fr['motion_weights_priors']
    np.testing.assert_allclose(pr[:, ::3, ::3, ::3], g['prior_sub'], rtol=0, atol=1e-6)
    np.testing.assert_allclose(pr.reshape(25, -1).sum(1), g['prior_chan_sum'], rtol=1e-6)
    K, E = syn.look_at_camera(512, 512, 1250.0)
    np.testing.assert_array_equal(K, g['K'])
    np.testing.assert_allclose(E, g['E'], atol=1e-7)
    np.testing.assert_allclose(syn.rotate_camera_y(E, 2 * np.pi * 37 / 360), g['E_rot37'], atol=1e-6)


def test_rays_and_bbox_bit_exact(golden_dir):
    g = _load(golden_dir, 'rays_bbox.npz')
    H, W = int(g['H']), int(g['W'])
    K64, E64 = g['K'].astype(np.float64), g['E'].astype(np.float64)     # float64 camera (training path)
    ro, rd = orc.rays_from_camera(H, W, K64, E64[:3, :3], E64[:3, 3])
    np.testing.assert_array_equal(ro.reshape(-1, 3), g['rays_o'])
    np.testing.assert_array_equal(rd.reshape(-1, 3), g['rays_d_in'])
    near, far, mask, d_out = orc.rays_bbox_near_far(g['bounds'], g['rays_o'], g['rays_d_in'])
    np.testing.assert_array_equal(mask, g['mask'])
    np.testing.assert_array_equal(near, g['near'])
    np.testing.assert_array_equal(far, g['far'])
    np.testing.assert_array_equal(d_out, g['rays_d_out'])
    xn, xf, xm, xd = orc.rays_bbox_near_far(g['bounds'], g['x_o'], g['x_d_in'])
    np.testing.assert_array_equal(xm, g['x_mask'])
    np.testing.assert_array_equal(xn, g['x_near'])
    np.testing.assert_array_equal(xf, g['x_far'])
    np.testing.assert_array_equal(xd, g['x_d_out'])
    # a host-side helper used by bench/tests must agree with the oracle too
    o2, d2 = syn.camera_rays_np(H, W, g['K'], g['E'])
    np.testing.assert_allclose(d2, g['rays_d_in'], atol=1e-12)


def test_patch_selection_bit_exact(golden_dir):
    g = _load(golden_dir, 'patches.npz')
    u_center = (g['choice_idx'] + 0.5) / g['choice_n']
    sel, masks, div, xy = orc.patch_ray_indices(
        g['ray_mask'], g['subject_mask'], int(g['H']), int(g['W']), int(g['patch_size']),
        g['u_subject'], u_center, float(g['sample_subject_ratio']))
    np.testing.assert_array_equal(sel, g['select_inds'])
    np.testing.assert_array_equal(masks, g['patch_masks'])
    np.testing.assert_array_equal(div, g['div_indices'])
    np.testing.assert_array_equal(xy, g['xy_min'])
    assert div[-1] == sel.size and masks.sum() == sel.size


def test_stage_vectors(golden_dir):
    g = _load(golden_dir, 'stages.npz')
    sd = syn.fixture_state_dict(syn.REFERENCE_PARAM_SHAPES, dense=True)
    fr = {k: T(v) for k, v in syn.make_frame(0).items()}
    near, far = T(g['near']), T(g['far'])
    z0 = orc.z_samples(near, far, 128)
    assert torch.equal(z0, T(g['z_plain']))
    z1 = orc.z_samples(near, far, 128, T(g['t_rand']))
    assert torch.equal(z1, T(g['z_jitter']))
    assert torch.equal(z0[:, 0], near[:, 0]) and torch.equal(z0[:, -1], far[:, 0])

    np.testing.assert_allclose(orc.rodrigues(T(g['rod_in'])).numpy(), g['rod_out'], atol=1e-7)
    Rs = orc.refine_pose(sd, fr['dst_Rs'], fr['dst_posevec'])
    np.testing.assert_allclose(Rs.numpy(), g['refined_Rs'], atol=2e-6)
    mRs, mTs = orc.motion_bases(T(g['refined_Rs']), fr['dst_Ts'], fr['cnl_gtfms'])
    np.testing.assert_allclose(mRs.numpy(), g['motion_scale_Rs'], atol=2e-6)
    np.testing.assert_allclose(mTs.numpy(), g['motion_Ts'], atol=2e-6)
    vol = orc.decode_motion_weight_volume(sd, fr['motion_weights_priors'])
    np.testing.assert_allclose(vol.numpy(), g['vol'], atol=2e-6)

    rays = T(g['rays'])
    o, d = rays[0].float(), rays[1].float()
    pts = o[:, None, :] + d[:, None, :] * z1[..., None]
    xs, m, bw = orc.sample_motion_fields(pts.reshape(-1, 3), T(g['motion_scale_Rs']), T(g['motion_Ts']),
                                         T(g['vol']), fr['cnl_bbox_min_xyz'], fr['cnl_bbox_scale_xyz'])
    np.testing.assert_allclose(bw.view(-1, 128, 24).numpy(), g['bw'], atol=2e-6)
    np.testing.assert_allclose(m.view(-1, 128, 1).numpy(), g['fg_mask'], atol=5e-6)
    np.testing.assert_allclose(xs.view(-1, 128, 3).numpy(), g['x_skel'], atol=1e-4)
    big = g['fg_mask'][..., 0] > 1e-3
    np.testing.assert_allclose(xs.view(-1, 128, 3).numpy()[big], g['x_skel'][big], atol=5e-6)

    xf = T(g['x_skel']).reshape(-1, 3)[:64]
    np.testing.assert_allclose(orc.fourier_pe(xf, 10).numpy(), g['pe10'], atol=1e-6)
    w_mid = orc.hann_weights(torch.full((1,), 30000), 6, 10000, 50000)
    np.testing.assert_allclose(w_mid.numpy(), [1, 1, 1, 0, 0, 0], atol=1e-6)
    np.testing.assert_allclose(orc.fourier_pe(xf, 6, w_mid, False).numpy(), g['pe6_mid'], atol=1e-6)
    w_full = orc.hann_weights(10000000, 6, 10000, 50000)
    np.testing.assert_allclose(orc.fourier_pe(xf, 6, w_full, False).numpy(), g['pe6_full'], atol=1e-6)
    xyz, off = orc.non_rigid_mlp(sd, T(g['pe6_full']), xf, fr['dst_posevec'][None])
    np.testing.assert_allclose(off.numpy(), g['nr_off'], atol=2e-6)
    np.testing.assert_allclose(xyz.numpy(), g['nr_xyz'], atol=2e-6)
    raw = orc.canonical_mlp(sd, orc.fourier_pe(T(g['nr_xyz']), 10))
    np.testing.assert_allclose(raw.numpy(), g['raw64'], rtol=1e-5, atol=2e-4)

    c = orc.composite(T(g['comp_raw']), T(g['comp_mask']), T(g['z_jitter']), d, T(g['comp_xyz']),
                      T(g['comp_bg']))
    for ours, ref in (('rgb', 'comp_rgb'), ('alpha', 'comp_acc'), ('weights_on_rays', 'comp_w'),
                      ('depth', 'comp_depth'), ('cnl_xyz', 'comp_cnl_xyz'), ('cnl_rgb', 'comp_cnl_rgb'),
                      ('cnl_weight', 'comp_wmax'), ('rgb_on_rays', 'comp_rgbs')):
        np.testing.assert_allclose(c[ours].numpy(), g[ref], rtol=1e-5, atol=1e-6, err_msg=ours)


def _e2e_inputs(g):
    fr = {k: T(v) for k, v in syn.make_frame(0).items()}
    fr['bgcolor'] = T(g['bgcolor'])
    st = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'stages.npz'))
    return fr, T(st['rays']), T(st['near']), T(st['far']), T(st['t_rand'])


KEYS = ['rgb', 'alpha', 'depth', 'weights_on_rays', 'xyz_on_rays', 'rgb_on_rays', 'cnl_xyz',
        'cnl_rgb', 'cnl_weight', 'backward_motion_weights', 'offsets']



PER_SAMPLE = ('xyz_on_rays', 'rgb_on_rays', 'offsets')


def _compare_outputs(out, g, prefix, atol=3e-5):
    """Per-sample keys are compared where the foreground likelihood is not
    negligible: x_skel = sum(w_i x_i) / clamp(sum w, 1e-4) amplifies fp32
    rounding by up to 1e4 for near-empty samples, and the 2^9 positional
    encoding frequency then decorrelates the MLP output.  Those samples carry
    weight <= mask ~ 1e-4 in every composited key, which are compared in full."""
    fg = g[prefix + 'backward_motion_weights'].sum(-1) > 1e-2
    for k in KEYS:
        a, b = out[k].numpy(), g[prefix + k]
        if k in PER_SAMPLE:
            a, b = a[fg], b[fg]
        tol = 2e-4 if k in ('xyz_on_rays', 'cnl_xyz') else atol
        np.testing.assert_allclose(a, b, rtol=1e-4, atol=tol, err_msg=k)


def test_network_forward_eval_and_train_with_grads(golden_dir):
    g = _load(golden_dir, 'network_e2e.npz')
    fr, rays, near, far, t_rand = _e2e_inputs(g)
    sd = {k: v.clone().requires_grad_(True)
          for k, v in syn.fixture_state_dict(syn.REFERENCE_PARAM_SHAPES, dense=True).items()}
    with torch.no_grad():
        ev = orc.network_forward(sd, rays, near, far, fr, 10000000)
    _compare_outputs(ev, g, 'eval_')
    tr = orc.network_forward(sd, rays, near, far, fr, torch.full((1,), 30000), t_rand=t_rand)
    _compare_outputs({k: v.detach() for k, v in tr.items()}, g, 'train_')
    loss = (tr['rgb'] * T(g['train_loss_w'])).sum() + 0.3 * tr['alpha'].sum() + 0.1 * tr['depth'].sum()
    assert abs(loss.item() - float(g['train_loss'])) < 1e-3
    loss.backward()
    names = [str(n) for n in g['grad_names']]
    for n, ref_l2 in zip(names, g['grad_l2']):
        ours = float(sd[n].grad.double().norm())
        # stated gradient tolerance: 2e-3 relative (L2) for the MLPs and the volume decoder,
        # 1e-2 for the pose decoder, whose gradient runs through x_skel / clamp(sum w, 1e-4)
        # and differs by 3.7e-3 between two fp32 CPU evaluations of the same graph
        rel = 1e-2 if n.startswith('pose_decoder') else 2e-3
        assert abs(ours - ref_l2) <= rel * max(ref_l2, 1e-6) + 1e-7, (n, ours, ref_l2)
    for key in g.files:
        if key.startswith('grad/'):
            ref = g[key]
            ours = sd[key[5:]].grad.numpy()
            scale = np.abs(ref).max() + 1e-12
            assert np.abs(ours - ref).max() <= (1e-2 if 'pose_decoder' in key else 4e-3) * scale, key


def test_analytic_known_answers():
    # compositing: zero density -> background colour, zero depth/acc
    R, S = 3, 16
    z = torch.linspace(1, 2, S).expand(R, S).contiguous()
    d = torch.tensor([[0.0, 0.0, 1.0]]).expand(R, 3)
    raw = torch.zeros(R, S, 4)
    c = orc.composite(raw, torch.ones(R, S, 1), z, d, torch.zeros(R, S, 3), torch.tensor([255.0, 0.0, 51.0]))
    np.testing.assert_allclose(c['rgb'].numpy(), np.tile([1.0, 0.0, 0.2], (R, 1)), atol=1e-7)
    assert c['alpha'].abs().max() == 0 and c['depth'].abs().max() == 0
    # constant positive density, mask 1 -> acc == 1 (last delta is 1e10)
    raw[..., 3] = 2.0
    c = orc.composite(raw, torch.ones(R, S, 1), z, d, torch.zeros(R, S, 3), torch.zeros(3))
    np.testing.assert_allclose(c['alpha'].numpy(), 1.0, atol=1e-5)
    # trilinear at voxel centres returns the voxel; far outside returns 0
    vol = torch.rand(2, 4, 4, 4)
    k = torch.tensor([[1.0, 2.0, 3.0]])
    gcoord = (2 * k / 3 - 1)[None].expand(2, 1, 3)
    np.testing.assert_allclose(orc.trilinear_zero_pad(vol, gcoord).numpy()[:, 0],
                               vol[:, 3, 2, 1].numpy(), atol=1e-6)
    out = orc.trilinear_zero_pad(vol, torch.full((2, 1, 3), 1.0 + 2.0 / 3 + 1e-3))
    assert out.abs().max() == 0
    # PE of zero: sin terms 0, cos terms 1
    pe = orc.fourier_pe(torch.zeros(1, 3), 10)
    assert pe.shape == (1, 63) and pe[0, :3].abs().max() == 0
    assert torch.equal(pe[0, 3:].view(10, 2, 3)[:, 0], torch.zeros(10, 3))
    assert torch.equal(pe[0, 3:].view(10, 2, 3)[:, 1], torch.ones(10, 3))
    # hann window end points
    assert orc.hann_weights(10000, 6, 10000, 50000).abs().max() == 0
    np.testing.assert_allclose(orc.hann_weights(50000, 6, 10000, 50000).numpy(), 1.0, atol=1e-6)
    # rodrigues(0) is not the identity: theta = sqrt(1e-5)
    r0 = orc.rodrigues(torch.zeros(1, 3))[0]
    np.testing.assert_allclose(torch.diagonal(r0).numpy(), np.cos(np.sqrt(1e-5)), atol=1e-7)
    # identity pose -> motion bases are identity
    fr = syn.make_frame(0)
    Rs0 = torch.eye(3).expand(24, 3, 3).contiguous()
    mRs, mTs = orc.motion_bases(Rs0, T(fr['dst_Ts']), T(fr['cnl_gtfms']))
    np.testing.assert_allclose(mRs.numpy(), np.tile(np.eye(3), (24, 1, 1)), atol=1e-6)
    np.testing.assert_allclose(mTs.numpy(), 0.0, atol=1e-6)


def test_train_step_cfg2_p20_oracle_and_lpips_against_reference():
    """The oracle (and `training.PerceptualLoss` on CPU tensors) against one full TRAIN STEP of the unmodified
    reference at cfg2's single_gpu.yaml size (6 x 20 x 20 patches = 2400 rays x 128 samples, iter 30000, perturb 1):
    forward outputs, the loss through the reference's Trainer.get_loss with LPIPS(net='vgg', pnet_rand=True), and
    the gradients of all 55 tensors.  Stated tolerances: outputs 1e-5 + 1e-5 |ref|, loss terms 1e-5 relative, gradients 2e-3
    relative L2 (pose decoder 1e-2, see test_network_forward_eval_and_train_with_grads)."""
    import step_fixture as sf
    st = sf.load_step('p20')
    fr = {k: T(v) for k, v in st['frame'].items()}
    fr['bgcolor'] = T(st['bgcolor'])
    sd = {k: v.clone().requires_grad_(True)
          for k, v in syn.fixture_state_dict(syn.REFERENCE_PARAM_SHAPES, dense=True).items()}
    out = orc.network_forward(sd, T(st['rays']), T(st['near']), T(st['far']), fr, torch.full((1,), st['iter']),
                              t_rand=st['t_rand'])
    for k in ('rgb', 'alpha', 'depth'):
        err = float(((out[k].detach() - T(st[k])).abs() - 1e-5 * T(st[k]).abs()).max())
        assert err < 1e-5, (k, err)
    lp = sf.make_perceptual(st)
    img = orc.unpack_patches(out['rgb'], T(st['patch_masks']), T(st['bgcolor']) / 255.0, st['div_indices'])
    tg = T(st['targets'])
    mse = torch.mean((img - tg) ** 2)
    lpips = torch.mean(lp(img.permute(0, 3, 1, 2) * 2 - 1, tg.permute(0, 3, 1, 2) * 2 - 1))
    assert abs(float(mse) * 0.2 - st['loss_mse']) <= 1e-5 * st['loss_mse']
    assert abs(float(lpips) - st['loss_lpips']) <= 1e-5 * st['loss_lpips'], (float(lpips), st['loss_lpips'])
    loss = 0.2 * mse + 1.0 * lpips
    assert abs(float(loss) - st['loss']) <= 1e-5 * st['loss']
    loss.backward()
    rows = sf.compare_grads(st, {k: v.grad for k, v in sd.items()}, 2e-3, pose_tol=1e-2, label='oracle p20', f64_tag='p20')
    worst = max(rows, key=lambda r: r[1])
    print(f'[parity] oracle vs reference train step p20: loss {float(loss):.6f} vs {st["loss"]:.6f}, '
          f'worst gradient rel L2 {worst[1]:.2e} ({worst[0]})')
