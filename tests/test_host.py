"""CPU-side tests: the C-ABI library loads and exports every symbol the header declares, the host
mirror of the reference interface (config, Network construction, parameter names, init parity,
patch sampling) behaves like the reference, and the product path fails loudly without CUDA."""
import ctypes
import json
import os
import re

import numpy as np
import pytest
import torch

from humannerf_b200 import _lib, synthetic as syn
from humannerf_b200.config import CfgNode, get_cfg, make_cfg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, 'tests', 'golden')


def _header_symbols():
    text = open(os.path.join(ROOT, 'include', 'humannerf_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(hn_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    if not os.path.isfile(_lib.lib_path()):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(_lib.lib_path())
    declared = _header_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f'{name} declared in include/humannerf_b200.h but not exported'
    assert sorted(_lib.SIGNATURES) == declared, set(_lib.SIGNATURES) ^ set(declared)
    loaded = _lib.load()
    assert loaded.hn_version() == _lib.HN_ABI_VERSION
    assert loaded.hn_has_tcgen05() == 1
    assert loaded.hn_last_error() == b'' or isinstance(loaded.hn_last_error(), bytes)


def test_argument_errors_do_not_need_a_gpu():
    lib = _lib.load()
    rc = lib.hn_sample_z(4, 0, None, None, None, None, None, None)
    assert rc < 0 and b'hn_sample_z' in lib.hn_last_error()
    rc = lib.hn_composite_fwd(1, 4096, *([None] * 15))
    assert rc < 0 and b'n_samples' in lib.hn_last_error()
    p = _lib.FieldParams()
    assert lib.hn_field_workspace_bytes(ctypes.byref(p), 10, 0, 0) == -1      # zero layers


def test_field_plan_sizes():
    """Workspace / packed-image sizes for the reference architecture (no device work)."""
    from humannerf_b200.ops import FieldSpec
    lib = _lib.load()
    fp = FieldSpec().params([None] * 7, [None] * 7, [None] * 9, [None] * 9)
    per_sample = lib.hn_field_workspace_bytes(ctypes.byref(fp), 1000, 1, 0) - 256
    # non-rigid 108+128*3+164+128*2+4, canonical 64+256*4+320+256*3+4 floats per sample
    assert per_sample == 1000 * 4 * ((108 + 128 * 3 + 164 + 128 * 2 + 4) + (64 + 256 * 4 + 320 + 256 * 3 + 4))
    # bf16 training workspace: T16 images of 8 tiles of 128 rows (encoding blocks, hidden activations H and
    # pre-activation gradients dZ per layer, 16-wide output dZ) plus the 1-bit ReLU mask images of the hidden layers
    # (one 8- / 16-byte record per row and column half)
    rows = 1024
    images = rows * 2 * ((48 + 6 * 2 * 128 + 16) + (64 + 8 * 2 * 256 + 16))
    masks = rows * 2 * (6 * 8 + 8 * 16)
    assert lib.hn_field_workspace_bytes(ctypes.byref(fp), 1000, 1, _lib.PRECISION_BF16) == images + masks + 256
    packed = lib.hn_field_packed_bytes(ctypes.byref(fp))
    # bf16 image (hidden layers carry one extra 16-row K step for the folded bias):
    nr = 48 * 128 + 3 * (128 + 16) * 128 + (128 + 48) * 128 + (128 + 16) * 128 + (128 + 16) * 16
    cn = 64 * 256 + 4 * (256 + 16) * 256 + (64 + 256) * 256 + 2 * (256 + 16) * 256 + (256 + 16) * 16
    # backward-data image: W^T windows, K = out features (16 for the output layers), N = width / pe_pad
    nr_bwd = 16 * 128 + 5 * 128 * 128 + 2 * 128 * 48
    cn_bwd = 16 * 256 + 7 * 256 * 256 + 2 * 256 * 64
    # the forward weights appear twice: the chunks of the SS kernel (mlp_tc_2sm.cu) and, with the same K rows grouped
    # per (layer, accumulator half), the chunks of the TS kernel (mlp_tc_ts.cu)
    elems = 2 * (nr + cn) + nr_bwd + cn_bwd
    assert 2 * elems < packed < 2 * elems + 64 * 1024


def test_ops_refuse_cpu_tensors():
    from humannerf_b200 import ops
    with pytest.raises(RuntimeError, match='CUDA'):
        ops.sample_z(torch.zeros(2, 1), torch.ones(2, 1), 8)


def test_config_surface(tmp_path):
    cfg = make_cfg()
    assert cfg.N_samples == 128 and cfg.chunk == 32768 and cfg.canonical_mlp.mlp_width == 256
    y = tmp_path / 'exp.yaml'
    y.write_text("task: 'zju_mocap'\nsubject: 'p387'\npatch:\n  size: 20\nbrand_new_key: 3\n")
    cfg = make_cfg([str(y)], ['N_samples', '64', 'train.lr', '0.001'])
    assert cfg.patch.size == 20 and cfg.patch.N_patches == 6 and cfg.brand_new_key == 3
    assert cfg.N_samples == 64 and cfg.train.lr == 0.001
    with pytest.raises(KeyError):
        make_cfg([], ['no.such.key', '1'])
    assert get_cfg() is make_cfg()
    assert isinstance(cfg.clone(), CfgNode)


def test_network_parameters_and_init_match_reference():
    from humannerf_b200.network import Network
    make_cfg()
    torch.manual_seed(0)
    net = Network()
    sd = net.state_dict()
    assert {k: tuple(v.shape) for k, v in sd.items()} == syn.REFERENCE_PARAM_SHAPES
    assert sum(p.numel() for p in net.parameters()) == 64417381
    g = json.load(open(os.path.join(GOLD, 'init_stats_seed0.json')))
    assert list(sd.keys()) == list(g.keys())
    for k, v in sd.items():
        assert abs(float(v.double().sum()) - g[k][0]) <= 1e-9 * max(1.0, abs(g[k][0])), k
        assert [float(x) for x in v.reshape(-1)[:4]] == g[k][2], k
    assert net.deploy_mlps_to_secondary_gpus() is net
    with pytest.raises(RuntimeError, match='CUDA'):
        fr = {k: torch.from_numpy(v) for k, v in syn.make_frame(0).items()}
        net(rays=torch.zeros(3, 2, 3), near=torch.ones(2, 1), far=torch.ones(2, 1) * 2, dst_Rs=fr['dst_Rs'],
            dst_Ts=fr['dst_Ts'], cnl_gtfms=fr['cnl_gtfms'], motion_weights_priors=fr['motion_weights_priors'],
            dst_posevec=fr['dst_posevec'], cnl_bbox_min_xyz=fr['cnl_bbox_min_xyz'],
            cnl_bbox_scale_xyz=fr['cnl_bbox_scale_xyz'])


def test_unsupported_variants_raise():
    from humannerf_b200.network import Network
    cfg = make_cfg()
    cfg.multihead.head_num = 4
    with pytest.raises(NotImplementedError, match='multihead'):
        Network()
    cfg = make_cfg()
    cfg.canonical_mlp.view_dir = True
    with pytest.raises(NotImplementedError):
        Network()
    make_cfg()


def test_host_patch_sampling_matches_reference():
    from humannerf_b200.data import sample_patches
    g = np.load(os.path.join(GOLD, 'patches.npz'))

    class Replay:                       # feeds back the two host randoms the reference drew
        def __init__(self):
            self.i = 0

        def rand(self, n):
            return np.array([g['u_subject'][self.i]])

        def choice(self, n, size=None, replace=True):
            assert n == g['choice_n'][self.i]
            v = np.array([g['choice_idx'][self.i]])
            self.i += 1
            return v
    sel, masks, div, xy = sample_patches(g['ray_mask'], g['subject_mask'], int(g['H']), int(g['W']), 6,
                                         int(g['patch_size']), Replay(), float(g['sample_subject_ratio']))
    assert np.array_equal(sel, g['select_inds']) and np.array_equal(masks, g['patch_masks'])
    assert np.array_equal(div, g['div_indices']) and np.array_equal(xy, g['xy_min'])


def test_hann_window_weights():
    from humannerf_b200.embedders.hannw_fourier import get_embedder, hann_window_weights
    from oracle import humannerf_oracle as orc
    make_cfg()
    for it in (0, 10000, 23456, 30000, 50000, 10000000, torch.full((1,), 30000)):
        assert torch.equal(hann_window_weights(it, 6), orc.hann_weights(it, 6, 10000, 50000))
    fn, dim = get_embedder(6, 30000, 0)
    x = torch.rand(5, 3)
    assert dim == 36 and torch.allclose(fn(x), orc.fourier_pe(x, 6, orc.hann_weights(30000, 6, 10000, 50000), False))
    from humannerf_b200.embedders.fourier import get_embedder as ge
    fn, dim = ge(10, 0)
    assert dim == 63 and torch.allclose(fn(x), orc.fourier_pe(x, 10))


def test_bench_reference_arm_runs_on_cpu():
    """`bench.py --impl reference` (the CPU arm the driver times next to ours: the oracle port of the reference's
    train step on all host cores) prints one JSON line with the contract's keys.  One bounded step."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, 'bench.py'), '--impl', 'reference', '--steps', '1',
                        '--warmup-ref', '0', '--ref-patch', '16'], capture_output=True, text=True, timeout=900, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith('{')][-1]
    d = json.loads(line)
    assert d['impl'] == 'reference' and d['metric'] == 'train_step_rays_per_sec' and d['value'] > 0
    assert '16x16' in d['config']['workload'] and d['steps'] == 1          # the line states what actually ran
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0


def test_first_deconv_function_matches_conv_transpose():
    """`nn_util._FirstDeconv` (the volume decoder's first ConvTranspose3d(k4, s2, p1) on its 1^3 input as a product
    with the central 2^3 taps, reference `core/utils/network_util.py:12-50`) against torch's full transposed
    convolution in float64: same output, same gradients for the embedding and every parameter - in particular exactly
    zero for the 56 taps the 1^3 input cannot reach - and the weight gradient comes back in the parameter's NDHWC
    strides (FusedAdam reads it in place)."""
    import torch
    from humannerf_b200.nn_util import ConvDecoder3D
    torch.manual_seed(0)
    dec = ConvDecoder3D(embedding_size=32, volume_size=8, voxel_channels=5).double()
    emb = torch.randn(1, 32, dtype=torch.double, requires_grad=True)
    out = dec(emb)
    g = torch.randn_like(out)
    out.backward(g)
    got = {k: v.grad.clone() for k, v in dec.named_parameters()}
    got_emb = emb.grad.clone()
    w0 = dec.block_conv[0].weight
    assert w0.grad.stride() == w0.stride() and not w0.is_contiguous()
    for p_ in dec.parameters():
        p_.grad = None
    emb.grad = None
    h = dec.block_mlp(emb)
    ref = dec.block_conv(h.view(1, -1, 1, 1, 1))
    assert float((ref - out).abs().max()) < 1e-12
    ref.backward(g)
    for k, v in dec.named_parameters():
        assert float((v.grad - got[k]).abs().max()) <= 1e-10 * max(1.0, float(v.grad.abs().max())), k
    assert float((emb.grad - got_emb).abs().max()) < 1e-10
    outer = got['block_conv.0.weight'].clone()
    outer[:, :, 1:3, 1:3, 1:3] = 0
    assert float(outer.abs().max()) == 0.0


def test_deferred_weight_gradients_context_logic():
    """ops.deferred_weight_gradients(): queued weight phases run exactly once - at flush_weight_gradients() or at the
    context's exit -, the context does not nest, and an exception inside it drops the queue instead of launching."""
    from humannerf_b200 import ops
    ran = []
    assert ops.flush_weight_gradients() == 0
    with ops.deferred_weight_gradients():
        ops._DEFERRED['jobs'].append(lambda: ran.append('a'))
        assert ops.flush_weight_gradients() == 1 and ran == ['a']
        ops._DEFERRED['jobs'].append(lambda: ran.append('b'))
    assert ran == ['a', 'b'] and not ops._DEFERRED['on'] and ops._DEFERRED['jobs'] == []
    with pytest.raises(RuntimeError):
        with ops.deferred_weight_gradients():
            with ops.deferred_weight_gradients():
                pass
    assert not ops._DEFERRED['on']
    with pytest.raises(ValueError):
        with ops.deferred_weight_gradients():
            ops._DEFERRED['jobs'].append(lambda: ran.append('never'))
            raise ValueError('step failed')
    assert ran == ['a', 'b'] and not ops._DEFERRED['on'] and ops._DEFERRED['jobs'] == []


def test_decoder_plan_sizes_and_phase_argument_errors():
    """Host-only entry points of the volume decoder and of the two-phase field backward (no device work): workspace =
    the saved activations of the reference decoder (network_util.py:25-38: 1024 @ 1^3, 512 @ 2^3, 512 @ 4^3, 256 @ 8^3,
    256 @ 16^3, 32-float logits rows @ 32^3), scratch at least the three rotating gradient buffers; bad geometries and a
    bad `phases` mask are refused with a message."""
    lib = _lib.load()
    ch = (ctypes.c_int * 6)(1024, 512, 512, 256, 256, 25)
    ws = lib.hn_vol_decoder_workspace_bytes(256, 1024, 5, ch)
    acts = [1024 * 1, 512 * 8, 512 * 64, 256 * 512, 256 * 4096, 32 * 32768]
    assert ws == 4 * sum((a + 63) // 64 * 64 for a in acts) + 256
    sc = lib.hn_vol_decoder_scratch_bytes(256, 1024, 5, ch)
    gmax = max(256 * 4096, 8 * 4096 * 32)                      # widest gradient buffer: d(logits) rows of 32 floats
    assert sc >= 4 * 3 * gmax
    small = (ctypes.c_int * 4)(1024, 512, 512, 25)
    assert 0 < lib.hn_vol_decoder_workspace_bytes(64, 1024, 3, small) < ws
    bad = (ctypes.c_int * 6)(1000, 512, 512, 256, 256, 25)
    assert lib.hn_vol_decoder_workspace_bytes(256, 1024, 5, bad) == -1 and b'hn_vol_decoder' in lib.hn_last_error()
    wide = (ctypes.c_int * 3)(1024, 1024, 25)
    assert lib.hn_vol_decoder_scratch_bytes(256, 1024, 2, wide) == -1 and b'output channels' in lib.hn_last_error()
    many = (ctypes.c_int * 6)(1024, 512, 512, 256, 256, 40)
    assert lib.hn_vol_decoder_workspace_bytes(256, 1024, 5, many) == -1
    # hn_field_bwd_bf16_phases: the phase mask is checked before anything touches the device
    from humannerf_b200.ops import FieldSpec
    fp = FieldSpec().params([None] * 7, [None] * 7, [None] * 9, [None] * 9)
    fg = _lib.FieldGrads()
    for phases in (0, 4, 7):
        rc = lib.hn_field_bwd_bf16_phases(ctypes.byref(fp), ctypes.byref(fg), 128, *([None] * 3), None, 0, *([None] * 4), None, 0, None,
                                          None, 0, None, phases, None)
        assert rc < 0 and b'phases' in lib.hn_last_error(), phases
    assert _lib.BWD_DATA == 1 and _lib.BWD_WEIGHTS == 2
