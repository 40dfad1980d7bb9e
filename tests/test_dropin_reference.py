"""Drop-in boundary: the REFERENCE's own loaders must pick this package up from a config change alone.

A scratch directory symlinks the reference's `configs/`, `core/`, `third_parties/` and this repo's `humannerf_b200/`
(INTEGRATION.md section 1); a subprocess started there imports the reference's `configs` with the dotted module strings of
INTEGRATION.md section 2 on the command line and calls the reference's `core.nets.create_network()` and
`core.train.create_optimizer()` - unmodified code that resolves `cfg.network_module` with `imp.load_source`
(core/nets/create_network.py:6-10) and groups parameters by name (optimizers/human_nerf/optimizer.py:12-43).

Runs in the build container only (the GPU box has no /root/reference): no CUDA needed, construction is host work."""
import json
import os
import subprocess
import sys
import textwrap

import pytest

REF = '/root/reference'
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = textwrap.dedent('''
    import importlib.util, json, sys, types
    import torch
    imp = types.ModuleType('imp')                      # Python 3.12 has no `imp` (SURVEY 5.9); two-line importlib shim
    def load_source(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec); sys.modules[name] = mod; spec.loader.exec_module(mod); return mod
    imp.load_source = load_source
    sys.modules['imp'] = imp
    sys.path.insert(0, '.')
    sys.argv = ['train.py', '--cfg', 'configs/human_nerf/zju_mocap/387/adventure.yaml',
                'network_module', 'humannerf_b200.network',
                'embedder.module', 'humannerf_b200.embedders.fourier',
                'non_rigid_embedder.module', 'humannerf_b200.embedders.hannw_fourier',
                'canonical_mlp.module', 'humannerf_b200.canonical_mlps.mlp_rgb_sigma',
                'non_rigid_motion_mlp.module', 'humannerf_b200.non_rigid_motion_mlps.mlp_offset',
                'mweight_volume.module', 'humannerf_b200.mweight_vol_decoders.deconv_vol_decoder',
                'pose_decoder.module', 'humannerf_b200.pose_decoders.mlp_delta_body_pose']
    real = torch.cuda.device_count
    torch.cuda.device_count = lambda: 1                 # configs/config.py wants one visible device
    from configs import cfg
    torch.cuda.device_count = real
    from core.nets import create_network
    from core.train import create_optimizer
    torch.manual_seed(0)
    net = create_network()
    opt = create_optimizer(net)
    import humannerf_b200.network, humannerf_b200.config
    sd = net.state_dict()
    out = {
        'class_module': type(net).__module__, 'class_file': sys.modules[type(net).__module__].__file__,
        'cfg_is_reference_cfg': humannerf_b200.config.get_cfg() is cfg,
        'has_b200_keys': 'b200' in cfg,
        'n_keys': len(sd), 'n_params': int(sum(v.numel() for v in sd.values())),
        'keys': sorted(sd), 'n_groups': len(opt.param_groups),
        'lrs': {g['name']: g['lr'] for g in opt.param_groups if g['name'] in ('mweight_vol_decoder', 'pose_decoder', 'non_rigid_mlp')},
        'deploy_returns_self': net.deploy_mlps_to_secondary_gpus() is net,
        'first4': {k: [float(x) for x in sd[k].reshape(-1)[:4]] for k in ('cnl_mlp.module.pts_linears.0.weight',
                                                                          'pose_decoder.block_mlps.8.weight')},
    }
    print('DROPIN ' + json.dumps(out))
''')


@pytest.mark.skipif(not os.path.isdir(REF), reason='needs the reference checkout (build container only)')
def test_reference_loaders_build_this_network(tmp_path, golden_dir):
    for name in ('configs', 'core', 'third_parties'):
        os.symlink(os.path.join(REF, name), tmp_path / name)
    os.symlink(os.path.join(REPO, 'humannerf_b200'), tmp_path / 'humannerf_b200')
    r = subprocess.run([sys.executable, '-c', SCRIPT], cwd=tmp_path, capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, PYTHONPATH=''))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith('DROPIN ')][-1]
    out = json.loads(line[7:])
    # the reference's create_network instantiated OUR class from the symlinked file
    assert out['class_module'] == 'humannerf_b200.network'
    assert os.path.realpath(out['class_file']) == os.path.realpath(os.path.join(REPO, 'humannerf_b200', 'network.py'))
    assert out['cfg_is_reference_cfg'] and out['has_b200_keys'] and out['deploy_returns_self']
    from humannerf_b200 import synthetic as syn
    assert out['keys'] == sorted(syn.REFERENCE_PARAM_SHAPES)
    assert out['n_keys'] == 55 and out['n_params'] == 64417381
    # the reference's optimizer: one group per tensor, per-name learning rates matched by substring
    assert out['n_groups'] == 55
    assert out['lrs'] == {'mweight_vol_decoder': 5e-5, 'pose_decoder': 5e-5, 'non_rigid_mlp': 5e-5}
    # same RNG consumption as the reference's constructor: identical initial weights under manual_seed(0)
    with open(os.path.join(golden_dir, 'init_stats_seed0.json')) as f:
        stats = json.load(f)
    for k, v in out['first4'].items():
        assert v == pytest.approx(stats[k][2], rel=0, abs=0), k
