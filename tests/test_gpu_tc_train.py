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
