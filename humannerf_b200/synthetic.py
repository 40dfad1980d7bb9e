"""Synthetic scene inputs for parity tests, smoke() and bench.py.

The licensed SMPL model and the ZJU-Mocap frames are not available offline, so
every test and benchmark runs on a synthetic 24-joint skeleton, a seeded random
pose and a pin-hole camera at the ZJU-387 `adventure.yaml` shape (SURVEY.md
section 8d).  These are *input generators* (what the reference's Dataset does on
the host before the hot path starts), written with plain numpy; they follow the
behaviour of

* reference `core/utils/body_util.py:222-249`  (pose -> per-joint local R, T)
* reference `core/utils/body_util.py:251-272`  (canonical global transforms)
* reference `core/utils/body_util.py:140-171, 274-348` (Gaussian bone priors)
* reference `core/data/human_nerf/tpose.py:65-84` + `core/utils/camera_util.py:49-81`
  (look-at camera)

and are pinned against the reference's own helpers by
`tests/test_synthetic_golden.py` (golden values in `tests/golden/`).
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np

TOTAL_BONES = 24

# parent[i] for joints 1..23 (SMPL kinematic tree; joint 0 is the root)
SMPL_PARENT = np.array(
    [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21],
    dtype=np.int64)

# joints that carry the wider torso Gaussians
_TORSO = (0, 3, 6, 9, 13, 14)
_HEAD = 15

_BONE_STDS = np.array([0.03, 0.06, 0.03])
_HEAD_STDS = np.array([0.06, 0.06, 0.06])
_JOINT_STDS = np.array([0.02, 0.02, 0.02])

# Fixed T-pose table in metres.
TPOSE_JOINTS = np.array([
    [0.00, 0.00, 0.00], [0.07, -0.09, 0.00], [-0.07, -0.09, 0.00], [0.00, 0.11, 0.00],
    [0.10, -0.47, 0.00], [-0.10, -0.47, 0.00], [0.00, 0.25, 0.00], [0.09, -0.87, -0.03],
    [-0.09, -0.87, -0.03], [0.00, 0.30, 0.00], [0.11, -0.93, 0.09], [-0.11, -0.93, 0.09],
    [0.00, 0.51, 0.00], [0.08, 0.42, 0.00], [-0.08, 0.42, 0.00], [0.00, 0.60, 0.03],
    [0.17, 0.45, 0.00], [-0.17, 0.45, 0.00], [0.43, 0.45, 0.00], [-0.43, 0.45, 0.00],
    [0.68, 0.45, 0.00], [-0.68, 0.45, 0.00], [0.77, 0.45, 0.00], [-0.77, 0.45, 0.00],
], dtype=np.float32)


def _skew(v: np.ndarray) -> np.ndarray:
    x, y, z = np.asarray(v).ravel()
    return np.array([[0.0, -z, y], [z, 0.0, -x], [-y, x, 0.0]])


def rvec_to_rmtx(rvec: np.ndarray) -> np.ndarray:
    """Axis-angle -> rotation, with the reference's `norm + 1e-5` guard
    (body_util.py:200-219): R = cos(t) I + sin(t) [r]x + (1 - cos t) r r^T."""
    rvec = np.asarray(rvec).reshape(3, 1)
    theta = np.linalg.norm(rvec)
    r = rvec / (theta + 1e-5)
    return (math.cos(theta) * np.eye(3) + math.sin(theta) * _skew(r)
            + (1.0 - math.cos(theta)) * r.dot(r.T))


def body_pose_to_body_RTs(jangles: np.ndarray, tpose_joints: np.ndarray):
    """Per-joint local rotation and parent-relative translation (float32)."""
    jangles = np.asarray(jangles).reshape(-1, 3)
    n = jangles.shape[0]
    Rs = np.zeros((n, 3, 3), dtype=np.float32)
    Ts = np.zeros((n, 3), dtype=np.float32)
    for i in range(n):
        Rs[i] = rvec_to_rmtx(jangles[i])
        Ts[i] = tpose_joints[i] if i == 0 else tpose_joints[i] - tpose_joints[SMPL_PARENT[i]]
    return Rs, Ts


def get_canonical_global_tfms(canonical_joints: np.ndarray) -> np.ndarray:
    """Global 4x4 transforms of the rest pose: pure translations chained down
    the tree, accumulated in float32 like the reference."""
    n = canonical_joints.shape[0]
    g = np.zeros((n, 4, 4), dtype=np.float32)
    for i in range(n):
        local = np.eye(4, dtype=np.float32)
        if i == 0:
            local[:3, 3] = canonical_joints[0]
            g[0] = local
        else:
            p = SMPL_PARENT[i]
            local[:3, 3] = canonical_joints[i] - canonical_joints[p]
            g[i] = g[p].dot(local)
    return g


def _align_rotation(v1: np.ndarray, v2: np.ndarray) -> np.ndarray:
    """Rotation taking unit(v1) onto unit(v2) (Rodrigues on the cross product)."""
    a = v1 / np.clip(np.linalg.norm(v1), 1e-5, None)
    b = v2 / np.clip(np.linalg.norm(v2), 1e-5, None)
    n = np.cross(a, b)
    c = float(a.dot(b))
    k = _skew(n).astype(np.float32)
    return (np.eye(3) + k + k.dot(k) * (1.0 / (1.0 + c))).astype(np.float32)


def _gaussian_volume(grid_size, bmin, bmax, center, S, R) -> np.ndarray:
    """exp(-x^T (R S S R^T) x) on a [z][y][x]-indexed grid spanning the bbox."""
    sigma = R.dot(S).dot(S).dot(R.T)
    zs = np.linspace(bmin[2], bmax[2], grid_size)
    ys = np.linspace(bmin[1], bmax[1], grid_size)
    xs = np.linspace(bmin[0], bmax[0], grid_size)
    zg, yg, xg = np.meshgrid(zs, ys, xs, indexing='ij')
    g = np.stack([xg - center[0], yg - center[1], zg - center[2]], axis=-1)
    d = np.einsum('abci,abci->abc', np.einsum('abci,ij->abcj', g, sigma), g)
    return np.exp(-d)


def _inv_std_matrix(stds) -> np.ndarray:
    m = np.eye(3, dtype=np.float32)
    for a in range(3):
        m[a, a] = 1.0 / stds[a]
    return m


def approx_gaussian_bone_volumes(tpose_joints, bbox_min_xyz, bbox_max_xyz,
                                 grid_size: int = 32) -> np.ndarray:
    """(25, V, V, V) prior: one Gaussian blob per bone that starts at a joint
    (summed over its children), an isotropic blob at leaf joints, plus a
    background channel; channels normalised to sum to one."""
    tj = np.asarray(tpose_joints, dtype=np.float32)
    n = tj.shape[0]
    y_axis = np.array([0.0, 1.0, 0.0], dtype=np.float32)
    vols = []
    for j in range(n):
        vol = np.zeros((grid_size,) * 3, dtype=np.float32)
        children = [c for c in range(1, n) if SMPL_PARENT[c] == j]
        for c in children:
            S = _inv_std_matrix(_BONE_STDS * 2.0)
            if j in _TORSO:
                S[0, 0] *= 1 / 1.5
                S[2, 2] *= 1 / 1.5
            start, end = tj[j], tj[c]
            R = _align_rotation(y_axis, end - start)
            vol = vol + _gaussian_volume(grid_size, bbox_min_xyz, bbox_max_xyz,
                                         (start + end) / 2.0, S, R)
        if not children:
            stds = _HEAD_STDS if j == _HEAD else _JOINT_STDS
            vol = _gaussian_volume(grid_size, bbox_min_xyz, bbox_max_xyz, tj[j],
                                   _inv_std_matrix(stds * 2.0), np.eye(3, dtype=np.float32))
        vols.append(vol)
    vols = np.stack(vols, axis=0)
    bg = 1.0 - np.sum(vols, axis=0, keepdims=True).clip(min=0.0, max=1.0)
    vols = np.concatenate([vols, bg], axis=0)
    return vols / np.sum(vols, axis=0, keepdims=True).clip(min=0.001)


def look_at_camera(img_w: int, img_h: int, focal: float, radius: float = 6.0,
                   y: float = -0.25):
    """K (3,3) and E (4,4), float32: camera at (0, y, radius) looking at
    (0, y, 0) with the image y axis pointing down (tpose.py `inv_camera=True`)."""
    campos = np.array([0.0, y, radius], dtype=np.float32)
    up = np.array([0.0, -1.0, 0.0], dtype=np.float32)
    fwd = np.array([0.0, y, 0.0], dtype=np.float32) - campos
    fwd /= np.linalg.norm(fwd)
    right = np.cross(up, fwd)
    right /= np.linalg.norm(right)
    up = np.cross(fwd, right)
    up /= np.linalg.norm(up)
    rot = np.array([right, up, fwd], dtype=np.float32)
    E = np.eye(4, dtype=np.float32)
    E[:3, :3] = rot
    E[:3, 3] = -rot.dot(campos)
    K = np.eye(3, dtype=np.float32)
    K[0, 0] = focal
    K[1, 1] = focal
    K[0, 2] = img_w / 2.0
    K[1, 2] = img_h / 2.0
    return K, E


def rotate_camera_y(E: np.ndarray, angle: float) -> np.ndarray:
    """Orbit the camera about the world y axis (freeview; camera_util.py:84-109,
    `trans=None`).  Returns a float64 4x4 like the reference."""
    inv_E = np.linalg.inv(E)
    camrot, campos = inv_E[:3, :3], inv_E[:3, 3]
    if camrot.T[1, 1] < 0.0:
        angle = -angle
    c, s = math.cos(angle), math.sin(angle)
    g = np.array([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]], dtype=np.float32)
    new_rot = g.dot(camrot)
    new_pos = g.dot(campos)
    out = np.identity(4)
    out[:3, :3] = new_rot.T
    out[:3, 3] = -new_rot.T.dot(new_pos)
    return out


def make_pose(seed: int = 0, scale: float = 0.2) -> np.ndarray:
    """72-d axis-angle pose, root rotation zero."""
    rng = np.random.RandomState(seed)
    poses = (scale * rng.randn(72)).astype(np.float32)
    poses[:3] = 0.0
    return poses


def make_frame(seed: int = 0, bbox_offset: float = 0.3, volume_size: int = 32,
               pose_scale: float = 0.2) -> Dict[str, np.ndarray]:
    """All per-frame tensors `Network.forward` takes besides the rays
    (same keys the reference Dataset emits, train.py:656-745)."""
    joints = TPOSE_JOINTS
    poses = make_pose(seed, pose_scale)
    dst_Rs, dst_Ts = body_pose_to_body_RTs(poses, joints)
    cnl_gtfms = get_canonical_global_tfms(joints)
    cnl_min = joints.min(axis=0) - bbox_offset
    cnl_max = joints.max(axis=0) + bbox_offset
    prior = approx_gaussian_bone_volumes(joints, cnl_min, cnl_max, volume_size).astype(np.float32)
    # posed joints -> destination bbox used for near/far (train.py:125-132)
    g = np.zeros((TOTAL_BONES, 4, 4))
    for i in range(TOTAL_BONES):
        local = np.eye(4)
        local[:3, :3] = dst_Rs[i]
        local[:3, 3] = dst_Ts[i]
        g[i] = local if i == 0 else g[SMPL_PARENT[i]].dot(local)
    posed = g[:, :3, 3].astype(np.float32)
    return {
        'poses': poses,
        'dst_Rs': dst_Rs, 'dst_Ts': dst_Ts, 'cnl_gtfms': cnl_gtfms,
        'motion_weights_priors': prior,
        'dst_posevec': (poses[3:] + 1e-2).astype(np.float32),
        'cnl_bbox_min_xyz': cnl_min.astype(np.float32),
        'cnl_bbox_max_xyz': cnl_max.astype(np.float32),
        'cnl_bbox_scale_xyz': (2.0 / (cnl_max - cnl_min)).astype(np.float32),
        'dst_bbox_min_xyz': (posed.min(axis=0) - bbox_offset).astype(np.float32),
        'dst_bbox_max_xyz': (posed.max(axis=0) + bbox_offset).astype(np.float32),
        'bgcolor': np.zeros(3, dtype=np.float32),
    }


# ---------------------------------------------------------------------------
# Deterministic fixture weights (independent of module construction order)
# ---------------------------------------------------------------------------
def fixture_state_dict(shapes, dense: bool = True, scale: float = 1.0):
    """Name-seeded weights for every parameter in `shapes` ({name: shape}).

    Random-init HumanNeRF renders an almost empty image (alpha_max ~ 8e-3,
    offsets ~ 1e-4: SURVEY.md section 8c), which makes every tolerance vacuous.
    `dense=True` therefore rescales the density row of the canonical output
    layer, the non-rigid output layer and the pose-decoder output layer so that
    alpha spans (0,1), offsets are centimetres (as in a trained model) and the
    pose refinement is non-trivial.  The same dict is loaded into the reference (when generating
    golden vectors), the oracle and the CUDA implementation."""
    import zlib
    import torch

    out = {}
    for name in sorted(shapes):
        shape = tuple(shapes[name])
        g = torch.Generator().manual_seed(zlib.crc32(name.encode()) & 0x7FFFFFFF)
        if name.endswith('const_embedding'):
            t = torch.randn(shape, generator=g)
        elif name.endswith('.bias'):
            t = (torch.rand(shape, generator=g) * 2 - 1) * 0.05
        else:
            if len(shape) == 5:      # ConvTranspose3d (in, out, 4,4,4), stride 2
                fan = (shape[0] + shape[1]) * 8
            else:
                fan = shape[0] + shape[1]
            bound = scale * math.sqrt(2.0) * math.sqrt(6.0 / fan)
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        out[name] = t.float()
    if dense:
        k = 'cnl_mlp.module.output_linear.0'
        if k + '.weight' in out:
            out[k + '.weight'][3] *= 25.0
            out[k + '.bias'][3] += 6.0
        k = 'non_rigid_mlp.module.block_mlps.12.weight'
        if k in out:
            out[k] *= 0.012
        k = 'pose_decoder.block_mlps.8.weight'
        if k in out:
            out[k] *= 0.1
    return out


REFERENCE_PARAM_SHAPES = {
    'mweight_vol_decoder.const_embedding': (256,),
    'mweight_vol_decoder.decoder.block_mlp.0.weight': (1024, 256),
    'mweight_vol_decoder.decoder.block_mlp.0.bias': (1024,),
    'mweight_vol_decoder.decoder.block_conv.0.weight': (1024, 512, 4, 4, 4),
    'mweight_vol_decoder.decoder.block_conv.0.bias': (512,),
    'mweight_vol_decoder.decoder.block_conv.2.weight': (512, 512, 4, 4, 4),
    'mweight_vol_decoder.decoder.block_conv.2.bias': (512,),
    'mweight_vol_decoder.decoder.block_conv.4.weight': (512, 256, 4, 4, 4),
    'mweight_vol_decoder.decoder.block_conv.4.bias': (256,),
    'mweight_vol_decoder.decoder.block_conv.6.weight': (256, 256, 4, 4, 4),
    'mweight_vol_decoder.decoder.block_conv.6.bias': (256,),
    'mweight_vol_decoder.decoder.block_conv.8.weight': (256, 25, 4, 4, 4),
    'mweight_vol_decoder.decoder.block_conv.8.bias': (25,),
    'non_rigid_mlp.module.block_mlps.0.weight': (128, 105),
    'non_rigid_mlp.module.block_mlps.0.bias': (128,),
    'non_rigid_mlp.module.block_mlps.2.weight': (128, 128),
    'non_rigid_mlp.module.block_mlps.2.bias': (128,),
    'non_rigid_mlp.module.block_mlps.4.weight': (128, 128),
    'non_rigid_mlp.module.block_mlps.4.bias': (128,),
    'non_rigid_mlp.module.block_mlps.6.weight': (128, 128),
    'non_rigid_mlp.module.block_mlps.6.bias': (128,),
    'non_rigid_mlp.module.block_mlps.8.weight': (128, 164),
    'non_rigid_mlp.module.block_mlps.8.bias': (128,),
    'non_rigid_mlp.module.block_mlps.10.weight': (128, 128),
    'non_rigid_mlp.module.block_mlps.10.bias': (128,),
    'non_rigid_mlp.module.block_mlps.12.weight': (3, 128),
    'non_rigid_mlp.module.block_mlps.12.bias': (3,),
    'cnl_mlp.module.pts_linears.0.weight': (256, 63),
    'cnl_mlp.module.pts_linears.0.bias': (256,),
    'cnl_mlp.module.pts_linears.2.weight': (256, 256),
    'cnl_mlp.module.pts_linears.2.bias': (256,),
    'cnl_mlp.module.pts_linears.4.weight': (256, 256),
    'cnl_mlp.module.pts_linears.4.bias': (256,),
    'cnl_mlp.module.pts_linears.6.weight': (256, 256),
    'cnl_mlp.module.pts_linears.6.bias': (256,),
    'cnl_mlp.module.pts_linears.8.weight': (256, 256),
    'cnl_mlp.module.pts_linears.8.bias': (256,),
    'cnl_mlp.module.pts_linears.10.weight': (256, 319),
    'cnl_mlp.module.pts_linears.10.bias': (256,),
    'cnl_mlp.module.pts_linears.12.weight': (256, 256),
    'cnl_mlp.module.pts_linears.12.bias': (256,),
    'cnl_mlp.module.pts_linears.14.weight': (256, 256),
    'cnl_mlp.module.pts_linears.14.bias': (256,),
    'cnl_mlp.module.output_linear.0.weight': (4, 256),
    'cnl_mlp.module.output_linear.0.bias': (4,),
    'pose_decoder.block_mlps.0.weight': (256, 69),
    'pose_decoder.block_mlps.0.bias': (256,),
    'pose_decoder.block_mlps.2.weight': (256, 256),
    'pose_decoder.block_mlps.2.bias': (256,),
    'pose_decoder.block_mlps.4.weight': (256, 256),
    'pose_decoder.block_mlps.4.bias': (256,),
    'pose_decoder.block_mlps.6.weight': (256, 256),
    'pose_decoder.block_mlps.6.bias': (256,),
    'pose_decoder.block_mlps.8.weight': (69, 256),
    'pose_decoder.block_mlps.8.bias': (69,),
}


def camera_rays_np(H: int, W: int, K: np.ndarray, E: np.ndarray):
    """Host-side float64 pixel rays (test/bench input helper; the device
    version is `hn_gen_rays`).  Returns rays_o, rays_d of shape (H*W, 3)."""
    R = np.asarray(E[:3, :3], dtype=np.float64)
    T = np.asarray(E[:3, 3], dtype=np.float64)
    Kinv = np.linalg.inv(np.asarray(K, dtype=np.float64))
    origin = -(R.T @ T)
    xs, ys = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64),
                         indexing='xy')
    pix = np.stack([xs, ys, np.ones_like(xs)], axis=-1).reshape(-1, 3)
    cam = pix @ Kinv.T
    world = (cam - T) @ R
    d = world - origin
    return np.broadcast_to(origin, d.shape).copy(), d
