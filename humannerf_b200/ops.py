"""Operators of the rendering hot path: thin `torch.autograd.Function` wrappers over
the C ABI (`humannerf_b200/_lib.py`).  PyTorch allocates; the kernels compute.

Operator <-> reference code (paths relative to the reference root)
  gen_rays / rays_bbox / compact_rays   core/utils/camera_util.py:132-208, train.py:592-616
  sample_z                              core/nets/human_nerf/network.py:455-471
  motion_field                          network.py:392-444 (_sample_motion_fields) + :499
  canonical_field                       network.py:233-327 (_apply_mlp_kernals) with the embedders,
                                        mlp_offset.py:74-114 and mlp_rgb_sigma.py:132-206
  composite                             network.py:355-388 (_raw2outputs)
  scatter_rows / gather_rows            trainer.py:28-37 (_unpack_imgs), run.py:49-65
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import FieldGrads, FieldParams, check, ptr, stream_ptr


# ---------------------------------------------------------------------------
# optional per-operator device timing (CUDA events on the launching stream; used by bench.py)
# ---------------------------------------------------------------------------
_TIMING = {'on': False, 'events': {}}


class _timed:
    def __init__(self, name: str):
        self.name = name

    def __enter__(self):
        if _TIMING['on']:
            self.s = torch.cuda.Event(enable_timing=True)
            self.e = torch.cuda.Event(enable_timing=True)
            self.s.record()
        return self

    def __exit__(self, *exc):
        if _TIMING['on']:
            self.e.record()
            _TIMING['events'].setdefault(self.name, []).append((self.s, self.e))
        return False


def timing_start() -> None:
    _TIMING['on'] = True
    _TIMING['events'] = {}


def timing_stop():
    """{op name: (total milliseconds, calls)}; synchronises the device."""
    _TIMING['on'] = False
    torch.cuda.synchronize()
    out = {k: (sum(s.elapsed_time(e) for s, e in v), len(v)) for k, v in _TIMING['events'].items()}
    _TIMING['events'] = {}
    return out


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError('humannerf_b200: expected a CUDA tensor - the rendering hot path has no CPU fallback')
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


# ---------------------------------------------------------------------------
# ray set-up
# ---------------------------------------------------------------------------
def gen_rays(H: int, W: int, K: np.ndarray, E: np.ndarray, device) -> Tuple[torch.Tensor, torch.Tensor]:
    """Pixel rays in float64: rays_o, rays_d of shape (H*W, 3) on `device`."""
    lib = _lib.load()
    Kh = np.ascontiguousarray(K, dtype=np.float64)[:3, :3].copy()
    Eh = np.ascontiguousarray(E, dtype=np.float64)
    if Eh.shape == (3, 4):
        Eh = np.concatenate([Eh, [[0, 0, 0, 1.0]]], 0)
    Eh = np.ascontiguousarray(Eh)
    o = torch.empty(H * W, 3, dtype=torch.float64, device=device)
    d = torch.empty(H * W, 3, dtype=torch.float64, device=device)
    with torch.cuda.device(o.device):
        check(lib.hn_gen_rays(H, W, Kh.ctypes.data, Eh.ctypes.data, ptr(o), ptr(d), stream_ptr()), 'gen_rays')
    return o, d


def rays_bbox(bounds: np.ndarray, rays_o: torch.Tensor, rays_d: torch.Tensor):
    """near, far (float64, dense) and mask (uint8) per ray; patches |d| < 1e-5 in `rays_d` IN PLACE."""
    lib = _lib.load()
    assert rays_o.dtype == torch.float64 and rays_d.dtype == torch.float64
    n = rays_o.shape[0]
    b = np.ascontiguousarray(np.asarray(bounds, dtype=np.float64).reshape(2, 3))
    near = torch.empty(n, dtype=torch.float64, device=rays_o.device)
    far = torch.empty(n, dtype=torch.float64, device=rays_o.device)
    mask = torch.empty(n, dtype=torch.uint8, device=rays_o.device)
    with torch.cuda.device(rays_o.device):
        check(lib.hn_rays_bbox(n, b.ctypes.data, ptr(rays_o), ptr(rays_d), ptr(near), ptr(far), ptr(mask),
                               stream_ptr()), 'rays_bbox')
    return near, far, mask


def compact_rays(mask, rays_o, rays_d, near, far):
    """Keep the rays with mask = 1 in pixel order; returns float32 o, d, near (R,1), far (R,1), pixel_index (R)."""
    lib = _lib.load()
    n = mask.shape[0]
    dev = mask.device
    o = torch.empty(n, 3, dtype=torch.float32, device=dev)
    d = torch.empty(n, 3, dtype=torch.float32, device=dev)
    nr = torch.empty(n, dtype=torch.float32, device=dev)
    fr = torch.empty(n, dtype=torch.float32, device=dev)
    idx = torch.empty(n, dtype=torch.int32, device=dev)
    count = torch.zeros(1, dtype=torch.int64, device=dev)
    ws = torch.empty(int(lib.hn_compact_rays_workspace_bytes(n)), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(lib.hn_compact_rays(n, ptr(mask), ptr(rays_o), ptr(rays_d), ptr(near), ptr(far), ptr(o), ptr(d),
                                  ptr(nr), ptr(fr), ptr(idx), ptr(count), ptr(ws), stream_ptr()), 'compact_rays')
    r = int(count.item())
    return o[:r], d[:r], nr[:r, None], fr[:r, None], idx[:r]


_T_TABLES = {}


def _t_table(n_samples: int, device) -> torch.Tensor:
    """torch.linspace(0, 1, S) evaluated on the CPU (bit-exact with the reference's table) and cached."""
    key = (n_samples, str(device))
    if key not in _T_TABLES:
        _T_TABLES[key] = torch.linspace(0.0, 1.0, steps=n_samples).to(device)
    return _T_TABLES[key]


def sample_z(near: torch.Tensor, far: torch.Tensor, n_samples: int,
             t_rand: Optional[torch.Tensor] = None) -> torch.Tensor:
    lib = _lib.load()
    near, far = _f32c(near).reshape(-1), _f32c(far).reshape(-1)
    r = near.shape[0]
    z = torch.empty(r, n_samples, dtype=torch.float32, device=near.device)
    if t_rand is not None:
        t_rand = _f32c(t_rand)
        assert t_rand.shape == (r, n_samples)
    with torch.cuda.device(near.device), _timed('sample_z'):
        check(lib.hn_sample_z(r, n_samples, ptr(near), ptr(far), ptr(_t_table(n_samples, near.device)),
                              ptr(t_rand), ptr(z), stream_ptr()), 'sample_z')
    return z


# ---------------------------------------------------------------------------
# motion field
# ---------------------------------------------------------------------------
class _MotionField(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rays_o, rays_d, z, pts, Rs, Ts, vol, bbox_min, bbox_scale, want_weights):
        lib = _lib.load()
        J = Rs.shape[0]
        V = vol.shape[-1]
        if pts is not None:
            n_rays, S = pts.shape[0], pts.shape[1]
            dev = pts.device
        else:
            n_rays, S = z.shape
            dev = z.device
        x_skel = torch.empty(n_rays, S, 3, dtype=torch.float32, device=dev)
        mask = torch.empty(n_rays, S, 1, dtype=torch.float32, device=dev)
        weights = torch.empty(n_rays, S, J, dtype=torch.float32, device=dev) if want_weights else None
        # corner table (27.6 MB at J = 24, V = 32): built by the forward call, reused by the backward
        tbytes = int(lib.hn_motion_field_workspace_bytes(J, V))
        table = torch.empty(tbytes, dtype=torch.uint8, device=dev) if n_rays * S >= 4096 else None
        with torch.cuda.device(dev), _timed('motion_field_fwd'):
            check(lib.hn_motion_field_fwd(n_rays, S, J, V, ptr(pts), ptr(rays_o), ptr(rays_d), ptr(z), ptr(Rs),
                                          ptr(Ts), ptr(vol), ptr(bbox_min), ptr(bbox_scale), ptr(x_skel),
                                          ptr(mask), ptr(weights), ptr(table), tbytes if table is not None else 0,
                                          stream_ptr()), 'motion_field_fwd')
        ctx.table = table
        ctx.save_for_backward(rays_o, rays_d, z, pts, Rs, Ts, vol, bbox_min, bbox_scale, x_skel, mask)
        ctx.dims = (n_rays, S, J, V)
        if weights is None:
            weights = torch.empty(0, device=dev)
            ctx.mark_non_differentiable(weights)
        return x_skel, mask, weights

    @staticmethod
    def backward(ctx, g_x, g_m, g_w):
        lib = _lib.load()
        rays_o, rays_d, z, pts, Rs, Ts, vol, bbox_min, bbox_scale, x_skel, mask = ctx.saved_tensors
        n_rays, S, J, V = ctx.dims
        dev = Rs.device
        g_x = _f32c(g_x) if g_x is not None else None
        g_m = _f32c(g_m) if g_m is not None else None
        g_w = _f32c(g_w) if (g_w is not None and g_w.numel() > 0) else None
        d_vol = torch.zeros_like(vol)
        d_Rs = torch.zeros_like(Rs)
        d_Ts = torch.zeros_like(Ts)
        # scratch for the vector-reduction scatter of the volume gradient (same size as the corner table)
        gbytes = int(lib.hn_motion_field_workspace_bytes(J, V)) if n_rays * S >= 4096 else 0
        gtable = torch.empty(gbytes, dtype=torch.uint8, device=dev) if gbytes > 0 else None
        with torch.cuda.device(dev), _timed('motion_field_bwd'):
            check(lib.hn_motion_field_bwd(n_rays, S, J, V, ptr(pts), ptr(rays_o), ptr(rays_d), ptr(z), ptr(Rs),
                                          ptr(Ts), ptr(vol), ptr(bbox_min), ptr(bbox_scale), ptr(g_x), ptr(g_m),
                                          ptr(g_w), ptr(x_skel), ptr(mask), ptr(d_vol), ptr(d_Rs), ptr(d_Ts),
                                          ptr(ctx.table), ptr(gtable), gbytes, stream_ptr()),
                  'motion_field_bwd')
        return None, None, None, None, d_Rs, d_Ts, d_vol, None, None, None


def motion_field(Rs, Ts, vol, bbox_min, bbox_scale, *, pts=None, rays_o=None, rays_d=None, z=None,
                 want_weights=True):
    """x_skel (R,S,3), mask (R,S,1), weights (R,S,J) [or None].  `vol` is the first J channels of the
    motion-weight volume, (J,V,V,V) contiguous; Rs (J,3,3); Ts (J,3).  Differentiable in Rs, Ts, vol."""
    Rs, Ts, vol = _f32c(Rs), _f32c(Ts), _f32c(vol)
    bbox_min, bbox_scale = _f32c(bbox_min), _f32c(bbox_scale)
    if pts is not None:
        pts = _f32c(pts)
        assert pts.dim() == 3 and pts.shape[-1] == 3
    else:
        rays_o, rays_d, z = _f32c(rays_o), _f32c(rays_d), _f32c(z)
    x, m, w = _MotionField.apply(rays_o, rays_d, z, pts, Rs, Ts, vol, bbox_min, bbox_scale, bool(want_weights))
    return x, m, (w if want_weights else None)


# ---------------------------------------------------------------------------
# per-frame front end: Rodrigues + corrected rotations + forward kinematics
# ---------------------------------------------------------------------------
class _PoseFK(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rvec, dst_Rs, dst_Ts, parents):
        lib = _lib.load()
        J = dst_Rs.shape[0]
        par = (C.c_int32 * J)(*[int(v) for v in parents])
        Rs_out = torch.empty(J, 3, 3, dtype=torch.float32, device=dst_Rs.device)
        G = torch.empty(J, 4, 4, dtype=torch.float32, device=dst_Rs.device)
        with torch.cuda.device(dst_Rs.device):
            check(lib.hn_pose_fk_fwd(J, par, ptr(rvec), ptr(dst_Rs), ptr(dst_Ts), ptr(Rs_out), ptr(G), stream_ptr()),
                  'pose_fk_fwd')
        ctx.par = par
        if rvec is not None:
            ctx.save_for_backward(rvec, dst_Rs, dst_Ts, Rs_out, G)
        ctx.mark_non_differentiable(Rs_out)
        return Rs_out, G

    @staticmethod
    def backward(ctx, _g_rs, g_G):
        lib = _lib.load()
        rvec, dst_Rs, dst_Ts, Rs_out, G = ctx.saved_tensors
        J = dst_Rs.shape[0]
        d_rvec = torch.empty_like(rvec)
        with torch.cuda.device(dst_Rs.device):
            check(lib.hn_pose_fk_bwd(J, ctx.par, ptr(rvec), ptr(dst_Rs), ptr(Rs_out), ptr(dst_Ts), ptr(G), ptr(_f32c(g_G)),
                                     None, ptr(d_rvec), stream_ptr()), 'pose_fk_bwd')
        return d_rvec, None, None, None


def pose_fk(rvec: Optional[torch.Tensor], dst_Rs: torch.Tensor, dst_Ts: torch.Tensor, parents: Sequence[int]):
    """Corrected local rotations (J,3,3) and global joint transforms G (J,4,4) from the axis-angle corrections
    rvec (J-1,3) (None: no refinement), the frame's local rotations dst_Rs (J,3,3) and translations dst_Ts (J,3).
    Differentiable in rvec."""
    if rvec is not None:
        rvec = _f32c(rvec).reshape(-1, 3)
    return _PoseFK.apply(rvec, _f32c(dst_Rs), _f32c(dst_Ts), tuple(parents))


# ---------------------------------------------------------------------------
# a7: motion-weight volume decoder (Linear + LeakyReLU, transposed convolutions, softmax with the prior)
# ---------------------------------------------------------------------------
class _VolDecoder(torch.autograd.Function):
    """inputs: embedding (E), lin_w (H, E), lin_b (H), prior (C, V, V, V), then conv weights and biases
    [w0, b0, w1, b1, ...] - weights (C_in, C_out, 4, 4, 4) in channels_last_3d strides."""

    @staticmethod
    def forward(ctx, precision, embedding, lin_w, lin_b, prior, *wb):
        lib = _lib.load()
        ws_, bs_ = list(wb[0::2]), list(wb[1::2])
        n_conv = len(ws_)
        for w in ws_:
            if not (w.dim() == 5 and w.is_contiguous(memory_format=torch.channels_last_3d)) or w.dtype != torch.float32:
                raise RuntimeError('humannerf_b200.vol_decoder: transposed-convolution weights must be fp32 in '
                                   'channels_last_3d strides ([C_in][kz][ky][kx][C_out])')
        channels = [int(ws_[0].shape[0])] + [int(w.shape[1]) for w in ws_]
        ch = (C.c_int * (n_conv + 1))(*channels)
        E, H = int(embedding.numel()), int(lin_w.shape[0])
        embedding, lin_w, lin_b, prior = _f32c(embedding).reshape(-1), _f32c(lin_w), _f32c(lin_b), _f32c(prior)
        bs_ = [_f32c(b) for b in bs_]
        dev = embedding.device
        side = 1 << n_conv
        assert tuple(prior.shape) == (channels[-1], side, side, side), (prior.shape, channels, side)
        wbytes = int(lib.hn_vol_decoder_workspace_bytes(E, H, n_conv, ch))
        sbytes = int(lib.hn_vol_decoder_scratch_bytes(E, H, n_conv, ch))
        if wbytes < 0 or sbytes < 0:
            check(-1, 'vol_decoder_workspace_bytes')
        ws = torch.empty(wbytes, dtype=torch.uint8, device=dev)
        scratch = torch.empty(sbytes, dtype=torch.uint8, device=dev)
        vol = torch.empty_like(prior)
        wp = (C.c_void_p * n_conv)(*[w.data_ptr() for w in ws_])
        bp = (C.c_void_p * n_conv)(*[b.data_ptr() for b in bs_])
        with torch.cuda.device(dev), _timed('vol_decoder_fwd'):
            check(lib.hn_vol_decoder_fwd(E, H, n_conv, ch, ptr(embedding), ptr(lin_w), ptr(lin_b), wp, bp, ptr(prior), ptr(vol),
                                         ptr(ws), wbytes, ptr(scratch), sbytes, int(precision), stream_ptr()), 'vol_decoder_fwd')
        ctx.geom = (E, H, n_conv, channels)
        ctx.precision = int(precision)
        ctx.ws = ws
        ctx.save_for_backward(embedding, lin_w, vol, *ws_)
        return vol

    @staticmethod
    def backward(ctx, g_vol):
        lib = _lib.load()
        E, H, n_conv, channels = ctx.geom
        embedding, lin_w, vol, *ws_ = ctx.saved_tensors
        ch = (C.c_int * (n_conv + 1))(*channels)
        dev = vol.device
        g_vol = _f32c(g_vol)
        d_e, d_lw = torch.empty_like(embedding), torch.empty_like(lin_w)
        d_lb = torch.empty(H, dtype=torch.float32, device=dev)
        d_w = [torch.empty_like(w) for w in ws_]                 # preserve_format: the parameter's NDHWC strides
        d_b = [torch.empty(int(w.shape[1]), dtype=torch.float32, device=dev) for w in ws_]
        sbytes = int(lib.hn_vol_decoder_scratch_bytes(E, H, n_conv, ch))
        scratch = torch.empty(sbytes, dtype=torch.uint8, device=dev)
        wp = (C.c_void_p * n_conv)(*[w.data_ptr() for w in ws_])
        dwp = (C.c_void_p * n_conv)(*[w.data_ptr() for w in d_w])
        dbp = (C.c_void_p * n_conv)(*[b.data_ptr() for b in d_b])
        with torch.cuda.device(dev), _timed('vol_decoder_bwd'):
            check(lib.hn_vol_decoder_bwd(E, H, n_conv, ch, ptr(embedding), ptr(lin_w), wp, ptr(vol), ptr(g_vol), ptr(ctx.ws),
                                         ctx.ws.numel(), ptr(scratch), sbytes, ptr(d_e), ptr(d_lw), ptr(d_lb), dwp, dbp,
                                         ctx.precision, stream_ptr()), 'vol_decoder_bwd')
        ctx.ws = None
        out = [None, d_e, d_lw, d_lb, None]
        for w, b in zip(d_w, d_b):
            out += [w, b]
        return tuple(out)


def vol_decoder(embedding, lin_w, lin_b, prior, conv_weights, conv_biases, precision=_lib.PRECISION_FP32):
    """softmax_c(ConvDecoder3D(embedding) + log prior) -> (C, V, V, V) (deconv_vol_decoder.py:25-33, network_util.py:12-50)
    on the hand-written kernels of csrc/vol_decoder.cu; differentiable in everything but the prior.
    precision: PRECISION_FP32 = 3xTF32 products (fp32-level accuracy), PRECISION_BF16 = plain TF32 products."""
    wb = []
    for w, b in zip(conv_weights, conv_biases):
        wb += [w, b]
    return _VolDecoder.apply(int(precision), embedding, lin_w, lin_b, prior, *wb)


# ---------------------------------------------------------------------------
# canonical field (hann-PE -> non-rigid MLP -> PE -> canonical MLP)
# ---------------------------------------------------------------------------
class FieldSpec:
    """Static description of the two MLPs (layer counts, widths, skip positions, encodings)."""

    def __init__(self, nr_layers=7, nr_width=128, nr_skip=4, nr_freqs=6, cond_dim=69,
                 cn_layers=9, cn_width=256, cn_skip=5, cn_freqs=10):
        self.nr_layers, self.nr_width, self.nr_skip, self.nr_freqs = nr_layers, nr_width, nr_skip, nr_freqs
        self.cond_dim = cond_dim
        self.cn_layers, self.cn_width, self.cn_skip, self.cn_freqs = cn_layers, cn_width, cn_skip, cn_freqs

    def params(self, nr_w, nr_b, cn_w, cn_b) -> FieldParams:
        fp = FieldParams()
        fp.nr_layers, fp.nr_width, fp.nr_skip, fp.nr_freqs = self.nr_layers, self.nr_width, self.nr_skip, self.nr_freqs
        fp.cond_dim = self.cond_dim
        fp.cn_layers, fp.cn_width, fp.cn_skip, fp.cn_freqs = self.cn_layers, self.cn_width, self.cn_skip, self.cn_freqs
        for i in range(self.nr_layers):
            fp.nr_w[i], fp.nr_b[i] = ptr(nr_w[i]), ptr(nr_b[i])
        for i in range(self.cn_layers):
            fp.cn_w[i], fp.cn_b[i] = ptr(cn_w[i]), ptr(cn_b[i])
        return fp


def pack_field_bf16(spec: FieldSpec, nr_w, nr_b, cn_w, cn_b, cond: torch.Tensor) -> torch.Tensor:
    """bf16 weight image for the tcgen05 kernel (re-pack whenever weights or the pose change)."""
    lib = _lib.load()
    ws = [_f32c(t.detach()) for t in list(nr_w) + list(nr_b) + list(cn_w) + list(cn_b)]
    nl, cl = spec.nr_layers, spec.cn_layers
    fp = spec.params(ws[:nl], ws[nl:2 * nl], ws[2 * nl:2 * nl + cl], ws[2 * nl + cl:])
    nbytes = int(lib.hn_field_packed_bytes(C.byref(fp)))
    if nbytes <= 0:
        check(-1, 'field_packed_bytes')
    cond = _f32c(cond.detach()).reshape(-1)
    packed = torch.empty(nbytes, dtype=torch.uint8, device=cond.device)
    with torch.cuda.device(cond.device):
        check(lib.hn_field_pack_bf16(C.byref(fp), ptr(cond), ptr(packed), stream_ptr()), 'field_pack_bf16')
    packed._hn_keepalive = ws   # the pack kernels read the fp32 weights asynchronously
    return packed


class PackedField:
    """Persistent bf16 weight image.  The first `refresh` (and any later one after a parameter moved)
    builds the image with its tables; every other `refresh` only re-runs the conversion kernel on the
    same buffer (`hn_field_repack_bf16`: no allocation, no host->device copy, CUDA-graph capturable)."""

    def __init__(self, spec: FieldSpec):
        self.spec, self.image, self._key = spec, None, None

    def refresh(self, nr_w, nr_b, cn_w, cn_b, cond: torch.Tensor) -> torch.Tensor:
        ws = [_f32c(t.detach()) for t in list(nr_w) + list(nr_b) + list(cn_w) + list(cn_b)]
        key = tuple(t.data_ptr() for t in ws)
        if self.image is None or key != self._key:
            self.image = pack_field_bf16(self.spec, nr_w, nr_b, cn_w, cn_b, cond)
            self._key = key
            return self.image
        lib = _lib.load()
        nl, cl = self.spec.nr_layers, self.spec.cn_layers
        fp = self.spec.params(ws[:nl], ws[nl:2 * nl], ws[2 * nl:2 * nl + cl], ws[2 * nl + cl:])
        cond = _f32c(cond.detach()).reshape(-1)
        with torch.cuda.device(cond.device):
            check(lib.hn_field_repack_bf16(C.byref(fp), ptr(cond), ptr(self.image), stream_ptr()), 'field_repack_bf16')
        return self.image


# Deferred weight gradients (bf16 path).  Between the loss and the optimizer nothing reads the MLPs' weight gradients, but
# the data gradients feed the motion field's backward and - on several GPUs - the all-reduce of the boundary gradients.
# Inside `deferred_weight_gradients()` the field's backward launches only its data phase and returns gradient tensors
# whose CONTENT arrives when `flush_weight_gradients()` launches the weight phase (on the then-current stream, which must
# be ordered after the backward - the same stream in practice).  The context manager flushes on exit, so no code outside
# it can observe unfinished gradients.
_DEFERRED = {'on': False, 'jobs': []}


class deferred_weight_gradients:
    def __enter__(self):
        if _DEFERRED['on']:
            raise RuntimeError('humannerf_b200: deferred_weight_gradients() does not nest')
        _DEFERRED['on'], _DEFERRED['jobs'] = True, []
        return self

    def __exit__(self, *exc):
        try:
            if exc[0] is None:
                flush_weight_gradients()
        finally:
            _DEFERRED['on'], _DEFERRED['jobs'] = False, []
        return False


def flush_weight_gradients() -> int:
    """Launch the weight phases queued since the context was entered; returns how many."""
    jobs, _DEFERRED['jobs'] = _DEFERRED['jobs'], []
    for job in jobs:
        job()
    return len(jobs)


class _CanonicalField(torch.autograd.Function):
    """inputs: x_skel (P,3), cond (cond_dim), hann_w (nr_freqs), then the flat parameter list
    [nr_w..., nr_b..., cn_w..., cn_b...]."""

    @staticmethod
    def forward(ctx, spec: FieldSpec, ignore_non_rigid: bool, precision: int, packed, need_grad: bool, count, x_skel,
                cond, hann_w, *flat):
        lib = _lib.load()
        nl, cl = spec.nr_layers, spec.cn_layers
        flat = [_f32c(t) for t in flat]
        nr_w, nr_b, cn_w, cn_b = flat[:nl], flat[nl:2 * nl], flat[2 * nl:2 * nl + cl], flat[2 * nl + cl:]
        fp = spec.params(nr_w, nr_b, cn_w, cn_b)
        x_skel = _f32c(x_skel)
        P = x_skel.shape[0]
        dev = x_skel.device
        raw = torch.empty(P, 4, dtype=torch.float32, device=dev)
        xyz = torch.empty(P, 3, dtype=torch.float32, device=dev)
        off = torch.empty(P, 3, dtype=torch.float32, device=dev)
        wbytes = int(lib.hn_field_workspace_bytes(C.byref(fp), P, 1 if need_grad else 0, precision))
        if wbytes < 0:
            check(-1, 'field_workspace_bytes')
        ws = torch.empty(wbytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev), _timed('field_fwd_bf16' if precision == _lib.PRECISION_BF16 else 'field_fwd_fp32'):
            check(lib.hn_field_fwd(C.byref(fp), P, ptr(x_skel), ptr(cond), ptr(hann_w), 1 if ignore_non_rigid else 0,
                                   ptr(raw), ptr(xyz), ptr(off), ptr(ws), wbytes, 1 if need_grad else 0, precision,
                                   ptr(packed), ptr(count), stream_ptr()), 'field_fwd')
        if need_grad:
            ctx.spec, ctx.ignore, ctx.precision = spec, ignore_non_rigid, precision
            ctx.ws, ctx.packed, ctx.count = ws, packed, count
            ctx.save_for_backward(x_skel, cond, hann_w, xyz, *flat)
        return raw, xyz, off

    @staticmethod
    def backward(ctx, g_raw, g_xyz, g_off):
        lib = _lib.load()
        spec = ctx.spec
        nl, cl = spec.nr_layers, spec.cn_layers
        x_skel, cond, hann_w, xyz, *flat = ctx.saved_tensors
        nr_w, nr_b, cn_w, cn_b = flat[:nl], flat[nl:2 * nl], flat[2 * nl:2 * nl + cl], flat[2 * nl + cl:]
        fp = spec.params(nr_w, nr_b, cn_w, cn_b)
        P = x_skel.shape[0]
        dev = x_skel.device
        # one zero-filled buffer carved into 16-byte aligned views (one fill launch instead of one per tensor)
        offs, total = [], 0
        for t in flat:
            offs.append(total)
            total += (t.numel() + 3) // 4 * 4
        gbuf = torch.zeros(total, dtype=torch.float32, device=dev)
        grads = [gbuf[o:o + t.numel()].view(t.shape) for o, t in zip(offs, flat)]
        fg = FieldGrads()
        gn_w, gn_b, gc_w, gc_b = grads[:nl], grads[nl:2 * nl], grads[2 * nl:2 * nl + cl], grads[2 * nl + cl:]
        for i in range(nl):
            if not ctx.ignore:
                fg.nr_w[i], fg.nr_b[i] = ptr(gn_w[i]), ptr(gn_b[i])
        for i in range(cl):
            fg.cn_w[i], fg.cn_b[i] = ptr(gc_w[i]), ptr(gc_b[i])
        g_raw = _f32c(g_raw) if g_raw is not None else torch.zeros(P, 4, device=dev)
        g_xyz = _f32c(g_xyz) if g_xyz is not None else None
        g_off = _f32c(g_off) if g_off is not None else None
        d_x = torch.empty(P, 3, dtype=torch.float32, device=dev)
        if ctx.precision == _lib.PRECISION_BF16:
            sbytes = int(lib.hn_field_bwd_bf16_scratch_bytes(C.byref(fp), P))
            scratch = torch.empty(sbytes, dtype=torch.uint8, device=dev)
            ws, packed, count, ignore = ctx.ws, ctx.packed, ctx.count, ctx.ignore

            def run(phases: int) -> None:
                with torch.cuda.device(dev), _timed('field_bwd_bf16'):
                    check(lib.hn_field_bwd_bf16_phases(C.byref(fp), C.byref(fg), P, ptr(x_skel), ptr(xyz), ptr(cond),
                                                       ptr(hann_w), 1 if ignore else 0, ptr(g_raw), ptr(g_xyz), ptr(g_off),
                                                       ptr(d_x), ptr(ws), ws.numel(), ptr(packed), ptr(scratch), sbytes,
                                                       ptr(count), phases, stream_ptr()), 'field_bwd_bf16')

            if _DEFERRED['on']:
                # data phase now; the weight-gradient kernel when `flush_weight_gradients` says so (same stream)
                run(_lib.BWD_DATA)
                shapes = [t.shape for t in flat]
                before = [(None, 0) if t.grad is None else (t.grad, t.grad._version) for t in flat]

                def weight_phase():
                    run(_lib.BWD_WEIGHTS)
                    # What autograd did with the (then still empty) gradient tensors after this backward returned:
                    # handed them on (torch.autograd.grad: .grad untouched), let AccumulateGrad take them over (.grad
                    # IS the view of gbuf: complete now), or cloned / accumulated them into .grad - which then holds
                    # a copy taken before the weight phase ran and gets the content now.
                    for t, o, shp, (g0, v0) in zip(flat, offs, shapes, before):
                        view = gbuf[o:o + t.numel()].view(shp)
                        g1 = t.grad
                        if g1 is None or g1.data_ptr() == view.data_ptr():
                            continue
                        if g1 is not g0 or g1._version != v0:
                            g1.add_(view)
                _DEFERRED['jobs'].append(weight_phase)
            else:
                run(_lib.BWD_DATA | _lib.BWD_WEIGHTS)
        else:
            sbytes = int(lib.hn_field_bwd_scratch_bytes(C.byref(fp), P))
            scratch = torch.empty(sbytes, dtype=torch.uint8, device=dev)
            with torch.cuda.device(dev), _timed('field_bwd_fp32'):
                check(lib.hn_field_bwd(C.byref(fp), C.byref(fg), P, ptr(x_skel), ptr(cond), ptr(hann_w),
                                       1 if ctx.ignore else 0, ptr(g_raw), ptr(g_xyz), ptr(g_off), ptr(d_x),
                                       ptr(ctx.ws), ctx.ws.numel(), ptr(scratch), sbytes, stream_ptr()), 'field_bwd')
        ctx.ws = ctx.packed = ctx.count = None
        return (None, None, None, None, None, None, d_x, None, None, *grads)


def canonical_field(spec: FieldSpec, x_skel, cond, hann_w, nr_w, nr_b, cn_w, cn_b, *,
                    ignore_non_rigid=False, precision=_lib.PRECISION_FP32, packed=None, count=None):
    """raw (P,4), xyz (P,3), offsets (P,3) for canonical-space samples x_skel (P,3).
    `count` (bf16 path only): int64 device tensor with the number of valid leading rows of x_skel - the rest
    of the P rows is capacity that the kernels neither read nor write (device-side empty-space skipping)."""
    if count is not None:
        if precision != _lib.PRECISION_BF16:
            raise RuntimeError('humannerf_b200: a device-side sample count needs the bf16 path')
        count = count.to(torch.int64).reshape(1).contiguous()
    flat = list(nr_w) + list(nr_b) + list(cn_w) + list(cn_b)
    cond = _f32c(cond).reshape(-1)
    hann_w = _f32c(hann_w).reshape(-1)
    # Function.forward always runs with grad mode off, so the decision to keep activations for the
    # backward pass is taken here
    need_grad = torch.is_grad_enabled() and (x_skel.requires_grad or any(t.requires_grad for t in flat))
    return _CanonicalField.apply(spec, bool(ignore_non_rigid), int(precision), packed, need_grad, count, x_skel, cond,
                                 hann_w, *flat)


# ---------------------------------------------------------------------------
# compositing
# ---------------------------------------------------------------------------
class _Composite(torch.autograd.Function):
    @staticmethod
    def forward(ctx, raw, mask, z, rays_d, xyz, bgcolor, heavy):
        lib = _lib.load()
        R, S = z.shape
        dev = raw.device
        new = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
        rgb, acc, depth = new(R, 3), new(R), new(R)
        weights = new(R, S)
        cnl_xyz, cnl_rgb, w_max = new(R, 3), new(R, 3), new(R)
        rgbs = new(R, S, 3) if heavy else None
        with torch.cuda.device(dev), _timed('composite_fwd'):
            check(lib.hn_composite_fwd(R, S, ptr(raw), ptr(mask), ptr(z), ptr(rays_d), ptr(xyz), ptr(bgcolor),
                                       ptr(rgb), ptr(acc), ptr(depth), ptr(weights), ptr(cnl_xyz), ptr(cnl_rgb),
                                       ptr(w_max), ptr(rgbs), stream_ptr()), 'composite_fwd')
        ctx.save_for_backward(raw, mask, z, rays_d, bgcolor)
        if rgbs is None:
            rgbs = torch.empty(0, device=dev)
        ctx.mark_non_differentiable(cnl_xyz, cnl_rgb, w_max, rgbs)
        return rgb, acc, depth, weights, cnl_xyz, cnl_rgb, w_max, rgbs

    @staticmethod
    def backward(ctx, g_rgb, g_acc, g_depth, g_w, *_):
        lib = _lib.load()
        raw, mask, z, rays_d, bgcolor = ctx.saved_tensors
        R, S = z.shape
        dev = raw.device
        # the contiguous copies must stay referenced until the launch: a temporary freed inside the
        # argument list hands its block straight to the next temporary (g_acc and g_depth would alias)
        g_rgb, g_acc, g_depth, g_w = [(_f32c(g) if g is not None else None) for g in (g_rgb, g_acc, g_depth, g_w)]
        d_raw = torch.empty(R, S, 4, dtype=torch.float32, device=dev)
        d_mask = torch.empty(R, S, 1, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev), _timed('composite_bwd'):
            check(lib.hn_composite_bwd(R, S, ptr(raw), ptr(mask), ptr(z), ptr(rays_d), ptr(bgcolor), ptr(g_rgb),
                                       ptr(g_acc), ptr(g_depth), ptr(g_w), ptr(d_raw), ptr(d_mask),
                                       stream_ptr()), 'composite_bwd')
        return d_raw, d_mask, None, None, None, None, None


def composite(raw, mask, z, rays_d, xyz, bgcolor, heavy=True):
    """rgb (R,3), alpha (R), depth (R), weights (R,S), cnl_xyz, cnl_rgb, cnl_weight, rgb_on_rays.
    Differentiable in raw and mask (through rgb, alpha, depth, weights)."""
    raw, mask, z = _f32c(raw), _f32c(mask), _f32c(z)
    rays_d, xyz, bgcolor = _f32c(rays_d), _f32c(xyz), _f32c(bgcolor)
    return _Composite.apply(raw, mask, z, rays_d, xyz, bgcolor, bool(heavy))


# ---------------------------------------------------------------------------
# scatter / gather of rays into images or patches
# ---------------------------------------------------------------------------
class _ScatterRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, index, base):
        lib = _lib.load()
        out = base.clone()
        n, c = src.shape
        with torch.cuda.device(src.device):
            check(lib.hn_scatter_rows(n, c, ptr(src), ptr(index), ptr(out), stream_ptr()), 'scatter_rows')
        ctx.save_for_backward(index)
        return out

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        (index,) = ctx.saved_tensors
        g = _f32c(g)
        n = index.shape[0]
        c = g.shape[-1]
        out = torch.empty(n, c, dtype=torch.float32, device=g.device)
        g_base = None
        with torch.cuda.device(g.device):
            check(lib.hn_gather_rows(n, c, ptr(g.reshape(-1, c)), ptr(index), ptr(out), stream_ptr()), 'gather_rows')
            if ctx.needs_input_grad[2]:
                # rows that were overwritten by `src` do not depend on `base`
                g_base = g.clone()
                zeros = torch.zeros(n, c, dtype=torch.float32, device=g.device)
                check(lib.hn_scatter_rows(n, c, ptr(zeros), ptr(index), ptr(g_base), stream_ptr()), 'scatter_rows')
        return out, None, g_base


class _GatherRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, index):
        lib = _lib.load()
        n, c = index.shape[0], src.shape[-1]
        out = torch.empty(n, c, dtype=torch.float32, device=src.device)
        with torch.cuda.device(src.device):
            check(lib.hn_gather_rows(n, c, ptr(src.reshape(-1, c)), ptr(index), ptr(out), stream_ptr()), 'gather_rows')
        ctx.save_for_backward(index)
        ctx.src_shape = src.shape
        return out

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        (index,) = ctx.saved_tensors
        g = _f32c(g)
        c = g.shape[-1]
        out = torch.zeros(ctx.src_shape, dtype=torch.float32, device=g.device)
        with torch.cuda.device(g.device):
            # the index holds every source row at most once (a compaction), so a plain scatter is the adjoint
            check(lib.hn_scatter_rows(index.shape[0], c, ptr(g), ptr(index), ptr(out), stream_ptr()), 'scatter_rows')
        return out, None


def gather_rows(src: torch.Tensor, index: torch.Tensor) -> torch.Tensor:
    """out[r] = src[index[r]] for a (N, C) fp32 matrix; rows with a negative index come back as zeros.
    Differentiable in `src` provided no source row appears twice in `index` (compaction)."""
    return _GatherRows.apply(_f32c(src), index.to(torch.int64).contiguous())


def scatter_rows(src: torch.Tensor, index: torch.Tensor, base: torch.Tensor) -> torch.Tensor:
    """Copy of `base` (..., C) with row index[r] of its flattened (N, C) view replaced by src[r]."""
    return _ScatterRows.apply(_f32c(src), index.to(torch.int64).contiguous(), _f32c(base))


# ---------------------------------------------------------------------------
# T16 tile images and tensor-core weight gradients (building blocks of the bf16 training path)
# ---------------------------------------------------------------------------
def tile_image_bytes(n_samples: int, f_pad: int) -> int:
    return int(_lib.load().hn_tile_image_bytes(n_samples, f_pad))


def tile_pack(x: torch.Tensor, f_pad: int) -> torch.Tensor:
    """fp32 (P, F) -> bf16 T16 tile image (uint8 tensor) with f_pad >= F features per sample."""
    lib = _lib.load()
    x = _f32c(x)
    P, F = x.shape
    img = torch.empty(tile_image_bytes(P, f_pad), dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        check(lib.hn_tile_pack(P, F, f_pad, ptr(x), F, ptr(img), stream_ptr()), 'tile_pack')
    return img


def tile_unpack(img: torch.Tensor, n_samples: int, n_features: int, f_pad: int) -> torch.Tensor:
    lib = _lib.load()
    out = torch.empty(n_samples, n_features, dtype=torch.float32, device=img.device)
    with torch.cuda.device(img.device):
        check(lib.hn_tile_unpack(n_samples, n_features, f_pad, ptr(img), ptr(out), n_features, stream_ptr()),
              'tile_unpack')
    return out


def tile_wgrad(jobs: Sequence[dict], n_samples: int) -> None:
    """Run weight-gradient jobs (dicts with the fields of `_lib.WgradJob`; tensors for dz, x, dw, db)."""
    lib = _lib.load()
    arr = (_lib.WgradJob * len(jobs))()
    dev = None
    for a, j in zip(arr, jobs):
        dev = j['dw'].device
        a.dz, a.x, a.dw = ptr(j['dz']), ptr(j['x']), ptr(j['dw'])
        a.db = ptr(j.get('db'))
        for k in ('f_dz', 'f_x', 'm', 'x_f0', 'n', 'n_valid', 'dw_ld', 'dw_col0'):
            setattr(a, k, int(j[k]))
        a.bias_col = int(j.get('bias_col', -1))
    scratch = torch.empty(len(jobs) * C.sizeof(_lib.WgradJob), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev), _timed('tile_wgrad'):
        check(lib.hn_tile_wgrad(arr, len(jobs), n_samples, ptr(scratch), stream_ptr()), 'tile_wgrad')


# ---------------------------------------------------------------------------
# LPIPS head of one trunk stage (a16)
# ---------------------------------------------------------------------------
class _LpipsHead(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, lin, out):
        lib = _lib.load()
        n2, C, H, W = feat.shape
        n = n2 // 2
        sn, sc, sh, sw = feat.stride()
        if sh != W * sw:                       # positions must be one strided axis (true for NCHW and NHWC)
            feat = feat.contiguous()
            sn, sc, sh, sw = feat.stride()
        res = out.clone() if out is not None else torch.zeros(n, dtype=torch.float32, device=feat.device)
        with torch.cuda.device(feat.device), _timed('lpips_head_fwd'):
            check(lib.hn_lpips_head_fwd(n, C, H * W, ptr_any(feat), sn, sc, sw, ptr(lin), ptr(res), stream_ptr()), 'lpips_head_fwd')
        ctx.save_for_backward(feat, lin)
        ctx.has_out = out is not None
        return res

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        feat, lin = ctx.saved_tensors
        n2, C, H, W = feat.shape
        sn, sc, sh, sw = feat.stride()
        g = _f32c(g)
        d_feat = torch.empty_like(feat)        # preserves the strides of feat
        with torch.cuda.device(feat.device), _timed('lpips_head_bwd'):
            check(lib.hn_lpips_head_bwd(n2 // 2, C, H * W, ptr_any(feat), sn, sc, sw, ptr(lin), ptr(g), ptr_any(d_feat),
                                        stream_ptr()), 'lpips_head_bwd')
        return d_feat, None, (g if ctx.has_out else None)


def ptr_any(t: torch.Tensor) -> int:
    """Device pointer of a dense CUDA fp32 tensor with arbitrary (non-overlapping) strides."""
    if not t.is_cuda or t.dtype != torch.float32:
        raise RuntimeError('humannerf_b200: expected a CUDA fp32 tensor (there is no CPU path)')
    return t.data_ptr()


def lpips_head(feat: torch.Tensor, lin: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """LPIPS head of one trunk stage (third_parties/lpips/__init__.py:40-42, lpips.py:14-15, 78-92): `feat` (2n, C, H, W) holds the
    features of n predictions followed by their n targets; returns `out` (n,) + the stage's spatially averaged,
    `lin`-weighted squared difference of the unit-normalised features (`out` None: starts from zero)."""
    return _LpipsHead.apply(feat.float(), _f32c(lin).reshape(-1), out)

