"""Golden TRAIN STEPS at the benchmarked configuration (BASELINE.json configs[1]), produced by the
UNMODIFIED reference mounted at /root/reference:

  * 512x512 synthetic frame, rays + bbox from the reference's `camera_util`, 6 patches picked by the reference's
    `Dataset.get_patch_ray_indices` under `np.random.seed`;
  * `Network.forward` in train mode (perturb = 1, the stratified jitter drawn by the reference itself on the CPU
    generator under `torch.manual_seed`) on the dense fixture weights;
  * the loss through the reference's own `Trainer.get_loss` / `get_img_rebuild_loss` / `_unpack_imgs`
    (trainer.py:28-37, 97-175) with `LPIPS(net='vgg', pnet_rand=True)` constructed under `torch.manual_seed` - the
    ImageNet trunk cannot be downloaded offline, the vendored 1x1 `lin` weights ARE loaded;
  * `backward()`: gradient norms of all 55 parameter tensors, full gradients of the small tensors, a seeded
    4096-element sample of the large ones.

Two steps are written to tests/golden/train_step_cfg2.npz: `p32_*` (6 x 32 x 32 patches = what BASELINE.json's text
and default.yaml say, iter 60000 = full hann band) and `p20_*` (6 x 20 x 20 = what zju_mocap/387/single_gpu.yaml
sets, iter 30000 = mid hann window).  Rays are not stored (a 512^2 float64 frame): the replaying tests regenerate them
with functions that are pinned bit-exact elsewhere and check the stored float64 checksums.

Run in the build container only:  python tests/golden/make_golden_step.py   (about two minutes of CPU)."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference, t  # noqa: E402

from humannerf_b200 import synthetic as syn  # noqa: E402

H = W = 512
SAMPLE = 4096          # elements kept of every gradient tensor larger than FULL
FULL = 20000


def sample_index(name: str, numel: int) -> np.ndarray:
    """The element sample of a large gradient tensor: seeded by the tensor's name, sorted, without replacement."""
    seed = int.from_bytes(name.encode()[-8:].rjust(8, b'\0'), 'little') % (2 ** 31)
    return np.sort(np.random.RandomState(seed).choice(numel, SAMPLE, replace=False))


def main():
    cfg = import_reference()
    sys.modules.setdefault('wandb', types.ModuleType('wandb'))
    sys.modules['skimage.metrics'].structural_similarity = None       # metrics only (run.py); never called here
    torch.set_num_threads(os.cpu_count() or 8)
    from core.nets import create_network
    from core.utils import camera_util
    from core.data.human_nerf import tpose as ref_tpose
    from core.data.human_nerf.train import Dataset as RefTrainDataset
    from core.train.trainers.human_nerf import trainer as ref_trainer
    from third_parties.lpips import LPIPS

    frame = syn.make_frame(0)
    ft = {k: t(v) for k, v in frame.items()}
    K, E = ref_tpose.Dataset.setup_camera(img_size=H, radius=6.0, focal=1250.)
    ro, rd = camera_util.get_rays_from_KRT(H, W, K.astype(np.float64), E[:3, :3].astype(np.float64),
                                           E[:3, 3].astype(np.float64))
    ro, rd = ro.reshape(-1, 3).copy(), rd.reshape(-1, 3).copy()
    bounds = np.stack([frame['dst_bbox_min_xyz'], frame['dst_bbox_max_xyz']], 0)
    near, far, ray_mask = camera_util.rays_intersect_3d_bbox(bounds, ro, rd)      # patches rd in place
    ro_c, rd_c = ro[ray_mask], rd[ray_mask]
    near, far = near.astype('float32'), far.astype('float32')
    yy, xx = np.mgrid[0:H, 0:W]
    subject = ((((xx - W / 2) / (0.16 * W)) ** 2 + ((yy - H / 2) / (0.38 * H)) ** 2) < 1.0) & ray_mask.reshape(H, W)
    image = np.random.RandomState(1).rand(H, W, 3).astype(np.float32)
    bgcolor = np.array([30.0, 200.0, 90.0], dtype=np.float32)

    torch.manual_seed(0)
    net = create_network()
    net.load_state_dict(syn.fixture_state_dict({k: tuple(v.shape) for k, v in net.state_dict().items()}, dense=True),
                        strict=True)
    net.train()
    cfg.perturb = 1.
    cfg.N_samples = 128

    # the reference's loss object: trainer.py:70 builds LPIPS(net='vgg', ...) with ImageNet weights (a download); the
    # same module with pnet_rand=True under a seed is what both sides can construct offline
    LPIPS_SEED = 0
    torch.manual_seed(LPIPS_SEED)
    lp = LPIPS(net='vgg', lpips=cfg.lpips.lpips, layers=cfg.lpips.layers, pnet_rand=True, verbose=False)
    ref_trainer.set_requires_grad(lp, requires_grad=False)
    fake_trainer = types.SimpleNamespace(lpips=lp)
    fake_trainer.get_img_rebuild_loss = lambda *a, **k: ref_trainer.Trainer.get_img_rebuild_loss(fake_trainer, *a, **k)

    out = {'H': H, 'W': W, 'K': K, 'E': E, 'bgcolor': bgcolor, 'lpips_seed': LPIPS_SEED,
           'n_valid_rays': int(ray_mask.sum()),
           'rays_o_sum': np.float64(ro_c.sum()), 'rays_d_sum': np.float64(rd_c.sum()),
           'near_sum': np.float64(near.astype(np.float64).sum()), 'far_sum': np.float64(far.astype(np.float64).sum()),
           'image_sum': np.float64(image.astype(np.float64).sum()),
           'lossweights': np.array([cfg.train.lossweights.lpips, cfg.train.lossweights.mse])}
    for k in range(5):
        out[f'lpips_lin{k}'] = lp.lins[k].model[-1].weight.detach().numpy().reshape(-1).copy()
    # spot values of the seeded trunk (the replaying side must build bit-identical weights from the seed alone)
    tw = lp.net.slice1[0].weight.detach().reshape(-1)
    out['lpips_trunk_first8'] = tw[:8].numpy().copy()
    out['lpips_trunk_abs_sum'] = np.float64(sum(float(p.double().abs().sum()) for p in lp.net.parameters()))

    for tag, patch, it, np_seed, th_seed in (('p32', 32, 60000, 7, 1), ('p20', 20, 30000, 8, 2)):
        np.random.seed(np_seed)
        dummy = types.SimpleNamespace()
        dummy._get_patch_ray_indices = lambda *a, **k: RefTrainDataset._get_patch_ray_indices(dummy, *a, **k)
        sel, info, div = RefTrainDataset.get_patch_ray_indices(
            dummy, N_patch=6, ray_mask=ray_mask, subject_mask=subject, bbox_mask=ray_mask.reshape(H, W),
            patch_size=patch, H=H, W=W)
        # what Dataset.__getitem__ hands over (train.py:592-650): selected rays, near / far, target crops
        r_o, r_d = ro_c[sel], rd_c[sel]
        rays = np.stack([r_o, r_d, r_d], 0)
        nr, fr = near[sel][:, None], far[sel][:, None]
        targets = np.stack([image[y:y + patch, x:x + patch] for x, y in info['xy_min']], 0)
        data = dict(rays=t(rays), near=t(nr), far=t(fr), dst_Rs=ft['dst_Rs'], dst_Ts=ft['dst_Ts'],
                    cnl_gtfms=ft['cnl_gtfms'], motion_weights_priors=ft['motion_weights_priors'],
                    dst_posevec=ft['dst_posevec'], cnl_bbox_min_xyz=ft['cnl_bbox_min_xyz'],
                    cnl_bbox_max_xyz=ft['cnl_bbox_max_xyz'], cnl_bbox_scale_xyz=ft['cnl_bbox_scale_xyz'],
                    bgcolor=t(bgcolor), head_id=torch.tensor(0), dir_idx=torch.zeros(1, dtype=torch.long),
                    frame_id=torch.tensor(0))
        R = rays.shape[1]
        torch.manual_seed(th_seed)
        t_rand = torch.rand(R, 128)
        torch.manual_seed(th_seed)                       # the forward's first draw IS this jitter (network.py:468)
        res = net(**data, iter_val=torch.full((1,), it))
        # trainer.py:209-215: loss on the re-assembled patches, bgcolor scaled to [0, 1]
        loss, parts = ref_trainer.Trainer.get_loss(
            fake_trainer, net_output=res, patch_masks=t(info['mask']), bgcolor=t(bgcolor) / 255.,
            targets=t(targets), div_indices=div)
        net.zero_grad()
        loss.backward()
        grads = {k: p.grad.detach() for k, p in net.named_parameters()}
        names = sorted(grads)
        p = tag + '_'
        out.update({p + 'patch': patch, p + 'iter': it, p + 'np_seed': np_seed, p + 'torch_seed': th_seed,
                    p + 'select_inds': sel.astype(np.int32), p + 'patch_masks': info['mask'],
                    p + 'xy_min': np.asarray(info['xy_min']), p + 'div_indices': np.asarray(div),
                    p + 't_rand_sum': np.float64(t_rand.double().sum()),
                    p + 'rgb': res['rgb'].detach().numpy(), p + 'alpha': res['alpha'].detach().numpy(),
                    p + 'depth': res['depth'].detach().numpy(),
                    p + 'loss': np.float64(loss.item()), p + 'loss_mse': np.float64(float(parts['mse'])),
                    p + 'loss_lpips': np.float64(float(parts['lpips'])),
                    p + 'grad_names': np.array(names),
                    p + 'grad_l2': np.array([float(grads[k].double().norm()) for k in names]),
                    p + 'grad_sum': np.array([float(grads[k].double().sum()) for k in names])})
        for k in names:
            g = grads[k].reshape(-1).numpy()
            out[p + 'grad/' + k] = g.copy() if g.size <= FULL else g[sample_index(k, g.size)].copy()
        print(f'{tag}: R = {R}, loss {loss.item():.6f} (mse {float(parts["mse"]):.6f}, lpips {float(parts["lpips"]):.6f}), '
              f'alpha max {float(res["alpha"].max()):.3f}')
    path = os.path.join(HERE, 'train_step_cfg2.npz')
    np.savez_compressed(path, **out)
    print('written', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
