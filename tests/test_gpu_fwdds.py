"""The kernels behind the forward-side d logit_scale (tile_kernel<MODE_FWDEU>, row_ent_split_kernel,
sum_slots_dot_kernel) against fp32/fp64 torch on the same inputs; the multi-rank comparison against the entropy path is
in tests/dist_worker.py."""
import pytest
import torch

from conftest import has_b200

pytestmark = pytest.mark.gpu

LOG2E = 1.4426950408889634


@pytest.mark.parametrize("n,N,D,ranks,scale", [(2048, 8192, 256, 4, 14.285714), (1024, 8192, 512, 8, 100.0),
                                               (4096, 4096, 768, 2, 30.0)])
def test_row_sums_split_by_owner(n, N, D, ranks, scale):
    if not has_b200():
        pytest.fail("needs a B200 (sm_100a)")
    from mrclip_b200._cabi import Shape
    from mrclip_b200.engine import default_engine
    eng = default_engine()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(N + D)
    img = torch.nn.functional.normalize(torch.randn(N, D, generator=g), dim=-1)
    txt = torch.nn.functional.normalize(0.5 * img + 0.5 * torch.randn(N, D, generator=g) / D ** 0.5, dim=-1)
    ld = eng.padded_dim(D)
    img_all = torch.zeros((N, ld), dtype=torch.bfloat16, device=dev)
    txt_all = torch.zeros((N, ld), dtype=torch.bfloat16, device=dev)
    eng.pack(img.to(dev), img_all)
    eng.pack(txt.to(dev), txt_all)
    per = N // ranks
    assert eng.fwd_row_ent_ok(n, N, per)
    off = N - n                                             # the last row block, so that label_offset is exercised
    shape = Shape(n, N, D, off)
    s = torch.tensor([scale], device=dev)
    ws = torch.empty(int(eng.workspace_bytes(n, N, D)), dtype=torch.uint8, device=dev)
    emat = torch.empty(int(eng.gmat_bytes(n, N)) // 2, dtype=torch.bfloat16, device=dev)
    lse = torch.zeros(n, device=dev)
    col_m, col_l, diag2 = torch.zeros(N, device=dev), torch.zeros(N, device=dev), torch.zeros(n, device=dev)
    out = torch.zeros((64, 2, ranks), device=dev)
    eng.clip_fwd_tiles_eu(img_all[off:], txt_all, shape, s, 0, N, ws, emat)
    eng.clip_fwd_reduce(shape, ws, lse, col_m, col_l, diag2)
    eng.row_ent_split(shape, ws, lse, per, ranks, out)
    got = out.sum(0)[0].double().cpu()
    s2 = (scale * LOG2E) * (img_all[off:, :D].double() @ txt_all[:, :D].double().t())
    lse_ref = torch.logsumexp(s2 * (1 / LOG2E), dim=1) * LOG2E
    torch.testing.assert_close(lse.double().cpu(), lse_ref.cpu(), rtol=0, atol=2e-4)
    p = torch.exp2(s2 - lse_ref[:, None])
    want = torch.stack([(p[:, q * per:(q + 1) * per] * s2[:, q * per:(q + 1) * per]).sum() for q in range(ranks)]).cpu()
    torch.testing.assert_close(got, want, rtol=2e-4, atol=2e-4 * float(want.abs().max()))
    # the E block is the one the plain forward writes
    emat2 = torch.empty_like(emat)
    eng.clip_fwd_tiles_e(img_all[off:], txt_all, shape, s, 0, N, ws, emat2)
    npad = eng.padded_cols(N)
    assert torch.equal(emat.view(-1, npad)[:n, :N], emat2.view(-1, npad)[:n, :N])


def test_slot_sum_with_dot():
    if not has_b200():
        pytest.fail("needs a B200 (sm_100a)")
    from mrclip_b200.engine import default_engine
    eng = default_engine()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(5)
    slots = torch.randn(4, 1000, 200, generator=g).to(dev)
    feat = torch.zeros((1000, 208), dtype=torch.bfloat16, device=dev)
    feat[:, :200] = torch.randn(1000, 200, generator=g).to(dev)
    for dt in (torch.float32, torch.bfloat16):
        out = torch.empty((1000, 200), dtype=dt, device=dev)
        dot = torch.full((64,), 7.0, device=dev)
        eng.sum_slots_dot(slots, out, feat, dot)
        tot = slots.double().sum(0)
        torch.testing.assert_close(out.double(), tot, rtol=1e-2 if dt == torch.bfloat16 else 1e-6, atol=1e-2 if dt == torch.bfloat16 else 1e-5)
        want = (tot * feat[:, :200].double()).sum()
        assert abs(dot.sum().item() - want.item()) <= 1e-4 * abs(want.item()) + 1e-2
