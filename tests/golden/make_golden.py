"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED
reference (ChenYutongTHU/humannerf, mounted read-only at /root/reference).

Run in the build container only (`python tests/golden/make_golden.py`); the GPU
box has no /root/reference, it only replays the committed .npz files.

Three harness shims, none of which touches the reference
tree: a fake `imp` module (removed in Python 3.12), `torch.cuda.device_count`
patched to 1 while `configs` is imported, and `cfg.primary_gpus =
cfg.secondary_gpus = ['cpu']`.  Missing optional deps of the *data* layer
(`termcolor`, `skimage`, `imageio`, the dataset-preparation tool) are stubbed
with empty modules; `np.bool` is aliased to `bool` for train.py:244.
"""
import importlib.util
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.path.insert(0, REPO)

from humannerf_b200 import synthetic as syn  # noqa: E402


def import_reference():
    imp = types.ModuleType('imp')

    def load_source(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod
    imp.load_source = load_source
    sys.modules['imp'] = imp
    for missing in ('termcolor', 'skimage', 'skimage.metrics', 'imageio'):
        sys.modules.setdefault(missing, types.ModuleType(missing))
    sys.modules['termcolor'].colored = lambda s, *a, **k: s
    prep = types.ModuleType('tools.prepare_zju_mocap.prepare_dataset')
    prep.get_mask = lambda *a, **k: None
    sys.modules['tools.prepare_zju_mocap.prepare_dataset'] = prep
    if not hasattr(np, 'bool'):
        np.bool = bool
    os.chdir(REF)
    sys.path.insert(0, REF)
    sys.argv = ['make_golden', '--cfg', 'configs/human_nerf/zju_mocap/387/adventure.yaml']
    real = torch.cuda.device_count
    torch.cuda.device_count = lambda: 1
    try:
        from configs import cfg
    finally:
        torch.cuda.device_count = real
    cfg.primary_gpus = ['cpu']
    cfg.secondary_gpus = ['cpu']
    return cfg


def t(x):
    return torch.from_numpy(np.ascontiguousarray(x))


def main():
    cfg = import_reference()
    torch.set_num_threads(8)
    from core.nets import create_network
    from core.nets.human_nerf.network import Network as RefNetwork
    from core.utils import body_util, camera_util
    from core.utils.network_util import MotionBasisComputer, RodriguesModule
    from core.data.human_nerf import tpose as ref_tpose

    out_dir = HERE

    # ------------------------------------------------------------------ 1
    # input generators vs the reference's numpy helpers
    joints = syn.TPOSE_JOINTS
    poses = syn.make_pose(0)
    Rs, Ts = body_util.body_pose_to_body_RTs(poses, joints)
    gt = body_util.get_canonical_global_tfms(joints)
    bmin, bmax = joints.min(0) - cfg.bbox_offset, joints.max(0) + cfg.bbox_offset
    prior = body_util.approx_gaussian_bone_volumes(joints, bmin, bmax, 32).astype('float32')
    K, E = ref_tpose.Dataset.setup_camera(img_size=512, radius=6.0, focal=1250.)
    E_rot = camera_util.rotate_camera_by_frame_idx(extrinsics=E, frame_idx=37, period=360)
    np.savez_compressed(
        os.path.join(out_dir, 'synthetic_inputs.npz'),
        dst_Rs=Rs, dst_Ts=Ts, cnl_gtfms=gt, prior_sub=prior[:, ::3, ::3, ::3],
        prior_chan_sum=prior.reshape(25, -1).sum(1).astype(np.float64),
        prior_full_l1=np.float64(np.abs(prior.astype(np.float64)).sum()),
        K=K, E=E, E_rot37=E_rot)

    # ------------------------------------------------------------------ 2
    # rays + bbox intersection (float64 numpy path of the reference)
    frame = syn.make_frame(0)
    H = W = 48
    Ks = K.copy()
    Ks[:2] *= (W / 512.0)
    # float64 camera, like the training path (E comes out of apply_global_tfm_to_camera as float64)
    ro, rd = camera_util.get_rays_from_KRT(H, W, Ks.astype(np.float64), E[:3, :3].astype(np.float64),
                                           E[:3, 3].astype(np.float64))
    ro = ro.reshape(-1, 3).copy()
    rd = rd.reshape(-1, 3).copy()
    rd_in = rd.copy()
    bounds = np.stack([frame['dst_bbox_min_xyz'], frame['dst_bbox_max_xyz']], 0)
    near, far, mask = camera_util.rays_intersect_3d_bbox(bounds, ro, rd)   # mutates rd
    # extra hand-made rays: axis-aligned through the centre, edge grazers, tiny components
    c = 0.5 * (bounds[0] + bounds[1]).astype(np.float64)
    xo = np.array([c + [0, 0, 5.0], c + [0, 0, 5.0], c + [5.0, 0, 0], bounds[1] + [0.01, 0.01, 5.0],
                   c + [0, 3.0, 5.0], c + [0, 0, 0.0]], dtype=np.float64)
    xd = np.array([[0, 0, -1.0], [1e-7, -1e-6, -1.0], [-1.0, 0, 0], [0, 0, -1.0],
                   [0, 0, -1.0], [0.3, 0.2, -1.0]], dtype=np.float64)
    xd_in = xd.copy()
    xn, xf, xm = camera_util.rays_intersect_3d_bbox(bounds, xo, xd)
    np.savez_compressed(
        os.path.join(out_dir, 'rays_bbox.npz'),
        H=H, W=W, K=Ks, E=E, rays_o=ro, rays_d_in=rd_in, rays_d_out=rd, bounds=bounds,
        near=near, far=far, mask=mask,
        x_o=xo, x_d_in=xd_in, x_d_out=xd, x_near=xn, x_far=xf, x_mask=xm)

    # ------------------------------------------------------------------ 3
    # patch selection (integer index math + host RNG draws, logged)
    from core.data.human_nerf.train import Dataset as RefTrainDataset
    Hp = Wp = 64
    Kp = K.copy()
    Kp[:2] *= (Wp / 512.0)
    po, pd = camera_util.get_rays_from_KRT(Hp, Wp, Kp, E[:3, :3], E[:3, 3])
    po, pd = po.reshape(-1, 3).copy(), pd.reshape(-1, 3).copy()
    _, _, pmask = camera_util.rays_intersect_3d_bbox(bounds, po, pd)
    yy, xx = np.mgrid[0:Hp, 0:Wp]
    subject = (((xx - 32) / 9.0) ** 2 + ((yy - 30) / 20.0) ** 2) < 1.0
    subject &= pmask.reshape(Hp, Wp)
    draws = {'rand': [], 'choice': [], 'n': []}
    real_rand, real_choice = np.random.rand, np.random.choice

    def log_rand(*a):
        v = real_rand(*a)
        draws['rand'].append(float(v[0]))
        return v

    def log_choice(n, size=None, replace=True):
        v = real_choice(n, size=size, replace=replace)
        draws['choice'].append(int(v[0]))
        draws['n'].append(int(n))
        return v
    np.random.seed(123)
    np.random.rand, np.random.choice = log_rand, log_choice
    try:
        cfg_patch = 16
        dummy = types.SimpleNamespace()
        dummy._get_patch_ray_indices = lambda *a, **k: RefTrainDataset._get_patch_ray_indices(dummy, *a, **k)
        sel, info, div = RefTrainDataset.get_patch_ray_indices(
            dummy, N_patch=6, ray_mask=pmask, subject_mask=subject,
            bbox_mask=pmask.reshape(Hp, Wp), patch_size=cfg_patch, H=Hp, W=Wp)
    finally:
        np.random.rand, np.random.choice = real_rand, real_choice
    np.savez_compressed(
        os.path.join(out_dir, 'patches.npz'),
        H=Hp, W=Wp, patch_size=cfg_patch, ray_mask=pmask, subject_mask=subject,
        u_subject=np.array(draws['rand']), choice_idx=np.array(draws['choice']),
        choice_n=np.array(draws['n']), select_inds=sel, patch_masks=info['mask'],
        xy_min=info['xy_min'], div_indices=div,
        sample_subject_ratio=float(cfg.patch.sample_subject_ratio))

    # ------------------------------------------------------------------ 4
    # per-stage vectors from the reference's static methods / modules
    torch.manual_seed(0)
    net = create_network()
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    assert shapes == syn.REFERENCE_PARAM_SHAPES, 'reference parameter set changed'
    # init statistics of the reference constructor under manual_seed(0)
    init_stats = {k: [float(v.double().sum()), float(v.double().abs().sum()),
                      [float(x) for x in v.reshape(-1)[:4]]] for k, v in net.state_dict().items()}
    with open(os.path.join(out_dir, 'init_stats_seed0.json'), 'w') as f:
        json.dump(init_stats, f, indent=0)

    sd = syn.fixture_state_dict(shapes, dense=True)
    net.load_state_dict(sd, strict=True)
    net.eval()

    ft = {k: t(v) for k, v in frame.items()}
    R_small = 24
    # rays for the stage / e2e fixtures: a deterministic subset of the valid 48x48 rays
    valid = np.where(mask)[0]
    pick = valid[np.linspace(0, valid.shape[0] - 1, R_small).astype(int)]
    pick_c = np.searchsorted(valid, pick)
    rays_o = ro[pick]
    rays_d = rd[pick]            # after the in-place 1e-5 patch, as the Dataset passes them on
    near_s = near[pick_c][:, None].astype('float32')
    far_s = far[pick_c][:, None].astype('float32')
    rays = np.stack([rays_o, rays_d, rays_d], 0)

    with torch.no_grad():
        # z samples
        cfg.N_samples = 128
        z0 = RefNetwork._get_samples_along_ray(R_small, t(near_s), t(far_s))
        torch.manual_seed(1)
        t_rand = torch.rand(R_small, 128)
        torch.manual_seed(1)
        z1 = RefNetwork._stratified_sampling(z0)
        # pose refine + motion bases + volume
        pose_out = net.pose_decoder(ft['dst_posevec'][None])
        Rs_ref = torch.cat([ft['dst_Rs'][None][:, :1],
                            RefNetwork._multiply_corrected_Rs(ft['dst_Rs'][None][:, 1:], pose_out['Rs'])], 1)
        mRs, mTs = MotionBasisComputer(24)(Rs_ref, ft['dst_Ts'][None], ft['cnl_gtfms'][None])
        vol = net.mweight_vol_decoder(motion_weights_priors=ft['motion_weights_priors'][None])[0]
        rod_in = torch.tensor([[0.0, 0.0, 0.0], [0.3, -0.2, 0.1], [1e-3, 2e-3, -1e-3]])
        rod_out = RodriguesModule()(rod_in)
        # motion field
        ro_t, rd_t = t(rays_o).float(), t(rays_d).float()
        pts = ro_t[:, None, :] + rd_t[:, None, :] * z1[..., None]
        mv = RefNetwork._sample_motion_fields(
            pts=pts, motion_scale_Rs=mRs[0], motion_Ts=mTs[0], motion_weights_vol=vol,
            cnl_bbox_min_xyz=ft['cnl_bbox_min_xyz'], cnl_bbox_scale_xyz=ft['cnl_bbox_scale_xyz'],
            output_list=['x_skel', 'fg_likelihood_mask', 'backward_motion_weights'])
        # embedders
        from core.nets.human_nerf.embedders import fourier as ref_fourier
        from core.nets.human_nerf.embedders import hannw_fourier as ref_hann
        x_flat = mv['x_skel'].reshape(-1, 3)
        pe10 = ref_fourier.get_embedder(10, 0)[0](x_flat[:64])
        hann_fn = ref_hann.get_embedder(6, torch.full((1,), 30000), 0)[0]
        pe6_mid = hann_fn(x_flat[:64])
        pe6_full = ref_hann.get_embedder(6, 10000000, 0)[0](x_flat[:64])
        # MLPs
        cond = ft['dst_posevec'][None]
        nr = net.non_rigid_mlp.module(pos_embed=pe6_full, pos_xyz=x_flat[:64], condition_code=cond)
        raw64 = net.cnl_mlp.module(pos_embed=ref_fourier.get_embedder(10, 0)[0](nr['xyz']))
        # compositing on synthetic raw with a wide alpha range
        g = torch.Generator().manual_seed(5)
        raw_c = torch.randn(R_small, 128, 4, generator=g) * 3.0
        raw_c[..., 3] = raw_c[..., 3] * 10.0
        mask_c = torch.rand(R_small, 128, 1, generator=g) * 1.2
        mask_c[:, :20] = 0.0
        xyz_c = torch.randn(R_small, 128, 3, generator=g)
        bg = torch.tensor([30.0, 200.0, 90.0])
        comp = RefNetwork._raw2outputs(raw_c, mask_c, z1, rd_t, xyz_c, bg)
    np.savez_compressed(
        os.path.join(out_dir, 'stages.npz'),
        rays=rays, near=near_s, far=far_s, z_plain=z0.numpy(), t_rand=t_rand.numpy(),
        z_jitter=z1.numpy(), refined_Rs=Rs_ref[0].numpy(), motion_scale_Rs=mRs[0].numpy(),
        motion_Ts=mTs[0].numpy(), vol=vol.numpy().astype(np.float32),
        rod_in=rod_in.numpy(), rod_out=rod_out.numpy(),
        x_skel=mv['x_skel'].numpy(), fg_mask=mv['fg_likelihood_mask'].numpy(),
        bw=mv['backward_motion_weights'].numpy(),
        pe10=pe10.numpy(), pe6_mid=pe6_mid.numpy(), pe6_full=pe6_full.numpy(),
        nr_xyz=nr['xyz'].numpy(), nr_off=nr['offsets'].numpy(), raw64=raw64.numpy(),
        comp_raw=raw_c.numpy(), comp_mask=mask_c.numpy(), comp_xyz=xyz_c.numpy(), comp_bg=bg.numpy(),
        comp_rgb=comp[0].numpy(), comp_acc=comp[1].numpy(), comp_w=comp[2].numpy(),
        comp_depth=comp[3].numpy(), comp_cnl_xyz=comp[4].numpy(), comp_cnl_rgb=comp[5].numpy(),
        comp_wmax=comp[6].numpy(), comp_rgbs=comp[7].numpy())

    # ------------------------------------------------------------------ 5
    # end-to-end Network.forward: eval (perturb 0, iter 1e7) and train (perturb 1, iter 30000)
    data = dict(rays=t(rays), near=t(near_s), far=t(far_s), dst_Rs=ft['dst_Rs'], dst_Ts=ft['dst_Ts'],
                cnl_gtfms=ft['cnl_gtfms'], motion_weights_priors=ft['motion_weights_priors'],
                dst_posevec=ft['dst_posevec'], cnl_bbox_min_xyz=ft['cnl_bbox_min_xyz'],
                cnl_bbox_max_xyz=ft['cnl_bbox_max_xyz'], cnl_bbox_scale_xyz=ft['cnl_bbox_scale_xyz'],
                bgcolor=torch.tensor([255.0, 128.0, 0.0]), head_id=torch.tensor(0),
                dir_idx=torch.zeros(1, dtype=torch.long), frame_id=torch.tensor(0))
    e2e = {}
    cfg.perturb = 0.
    with torch.no_grad():
        ev = net(**data, iter_val=cfg.eval_iter)
    for k, v in ev.items():
        e2e['eval_' + k] = v.numpy()
    # train mode with gradients
    cfg.perturb = 1.
    net.train()
    torch.manual_seed(1)
    tr = net(**data, iter_val=torch.full((1,), 30000))
    gw = torch.Generator().manual_seed(9)
    loss_w = torch.rand(R_small, 3, generator=gw)
    loss = (tr['rgb'] * loss_w).sum() + 0.3 * tr['alpha'].sum() + 0.1 * tr['depth'].sum()
    net.zero_grad()
    loss.backward()
    for k, v in tr.items():
        e2e['train_' + k] = v.detach().numpy()
    e2e['train_loss_w'] = loss_w.numpy()
    e2e['train_loss'] = np.float64(loss.item())
    grads = {k: p.grad for k, p in net.named_parameters()}
    keep = ['non_rigid_mlp.module.block_mlps.12.weight', 'non_rigid_mlp.module.block_mlps.0.weight',
            'non_rigid_mlp.module.block_mlps.8.bias', 'cnl_mlp.module.output_linear.0.weight',
            'cnl_mlp.module.pts_linears.0.weight', 'cnl_mlp.module.pts_linears.10.bias',
            'pose_decoder.block_mlps.8.weight', 'mweight_vol_decoder.const_embedding']
    for k in keep:
        e2e['grad/' + k] = grads[k].numpy()
    e2e['grad_names'] = np.array(sorted(grads))
    e2e['grad_l2'] = np.array([float(grads[k].double().norm()) for k in sorted(grads)])
    e2e['bgcolor'] = data['bgcolor'].numpy()
    np.savez_compressed(os.path.join(out_dir, 'network_e2e.npz'), **e2e)
    print('alpha eval: min %.4f max %.4f mean %.4f' % (ev['alpha'].min(), ev['alpha'].max(), ev['alpha'].mean()))
    print('offset max abs: %.4f' % ev['offsets'].abs().max())
    print('written to', out_dir)


if __name__ == '__main__':
    main()
