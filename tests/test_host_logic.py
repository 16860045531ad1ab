"""Host-side logic of mrclip_b200/loss.py on CPU: signatures, label cache, mode coefficients and the
multi-rank orchestration (gloo, world_size 2 and 4), with the kernels replaced by the stand-in engine.

Every expectation is the unmodified reference's output (tests/golden/)."""
import inspect
import os
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import mrclip_b200
from conftest import StandInEngine, golden_names, load_golden, rel_err
from mrclip_b200 import ClipLoss, MultiPositiveClipLoss, SigLipLoss

# stand-in math is float64 but features are packed to bf16 (exact for the fixtures) and grads are
# returned in the input dtype (fp32)
LOSS_TOL, GRAD_TOL = 5e-6, 5e-5


def test_signatures_match_reference():
    sig = inspect.signature(ClipLoss.__init__)
    assert list(sig.parameters)[1:] == ["local_loss", "gather_with_grad", "cache_labels", "rank", "world_size",
                                        "use_horovod"]
    assert [p.default for p in list(sig.parameters.values())[1:]] == [False, False, False, 0, 1, False]
    assert list(inspect.signature(ClipLoss.forward).parameters)[1:] == ["image_features", "text_features",
                                                                        "logit_scale", "output_dict"]
    assert list(inspect.signature(SigLipLoss.__init__).parameters)[1:] == ["cache_labels", "rank", "world_size",
                                                                           "dist_impl"]
    assert list(inspect.signature(SigLipLoss.forward).parameters)[1:] == ["image_features", "text_features",
                                                                          "logit_scale", "logit_bias", "output_dict"]
    assert list(inspect.signature(mrclip_b200.gather_features).parameters) == [
        "image_features", "text_features", "local_loss", "gather_with_grad", "rank", "world_size", "use_horovod"]
    with pytest.raises(AssertionError):
        SigLipLoss(dist_impl="ring")
    with pytest.raises(NotImplementedError):
        ClipLoss(use_horovod=True)
    with pytest.raises(TypeError):
        ClipLoss()(torch.zeros(2, 2), torch.zeros(2, 2), 1.0, delta=0.5)   # reference raises TypeError too


def test_label_cache_semantics():
    dev = torch.device("cpu")
    m = ClipLoss(local_loss=True, cache_labels=True, rank=3, world_size=4)
    a = m.get_ground_truth(dev, 8)
    assert a.dtype == torch.long and torch.equal(a, torch.arange(8) + 24)
    assert m.get_ground_truth(dev, 8) is a and m.prev_num_logits == 8
    b = m.get_ground_truth(dev, 4)
    assert torch.equal(b, torch.arange(4) + 12) and m.prev_num_logits == 4
    n = ClipLoss(local_loss=False, cache_labels=False, rank=3, world_size=4)
    assert torch.equal(n.get_ground_truth(dev, 8), torch.arange(8)) and n.labels == {} and n.prev_num_logits == 0
    s = SigLipLoss()
    lab = s.get_ground_truth(dev, torch.float32, 3)
    assert torch.equal(lab, 2 * torch.eye(3) - 1)
    assert torch.equal(s.get_ground_truth(dev, torch.float32, 3, negative_only=True), -torch.ones(3, 3))


def _run_rank(case, rank, world, img, txt):
    m = case["meta"]
    n = img.shape[0] // world
    i_loc = torch.from_numpy(img[rank * n:(rank + 1) * n]).clone().requires_grad_(True)
    t_loc = torch.from_numpy(txt[rank * n:(rank + 1) * n]).clone().requires_grad_(True)
    scale = torch.tensor(float(m["scale"]), requires_grad=True)
    out = {}
    if m["kind"] == "clip":
        mod = ClipLoss(local_loss=bool(m["local_loss"]), gather_with_grad=bool(m["gather_with_grad"]),
                       cache_labels=True, rank=rank, world_size=world)
        res = mod(i_loc, t_loc, scale, output_dict=True)
        assert list(res) == ["contrastive_loss"]
        loss = res["contrastive_loss"]
        num_logits = n if (world > 1 and m["local_loss"]) else img.shape[0]
        out["labels"] = mod.get_ground_truth(i_loc.device, num_logits).numpy()
    elif m["kind"] == "mpos":
        mod = MultiPositiveClipLoss(local_loss=bool(m["local_loss"]), gather_with_grad=bool(m["gather_with_grad"]),
                                    cache_labels=False, rank=rank, world_size=world)
        res = mod(i_loc, t_loc, scale, delta=float(m["delta"]),
                  tokenized_texts=torch.from_numpy(case["ranks"][rank]["labels_in"]), output_dict=True)
        assert list(res) == ["multi contrastive_loss"]
        loss = res["multi contrastive_loss"]
    else:
        bias = torch.tensor(float(m["bias"]), requires_grad=True)
        loss = SigLipLoss(rank=rank, world_size=world)(i_loc, t_loc, scale, bias)
    assert loss.dim() == 0 and loss.dtype == torch.float32
    (loss * float(m["grad_output"])).backward()
    out.update(loss=loss.item(), d_image=i_loc.grad.numpy(), d_text=t_loc.grad.numpy(), d_scale=scale.grad.item())
    if m["kind"] == "siglip":
        out["d_bias"] = bias.grad.item()
    return out


def _compare(out, ref, kind, grad_tol=GRAD_TOL, scale_tol=1e-4):
    assert abs(out["loss"] - float(ref["loss"])) <= LOSS_TOL * max(1.0, abs(float(ref["loss"])))
    assert rel_err(out["d_image"], ref["d_image"]) <= grad_tol
    assert rel_err(out["d_text"], ref["d_text"]) <= grad_tol
    assert abs(out["d_scale"] - float(ref["d_scale"])) <= scale_tol * max(abs(float(ref["d_scale"])), 1e-3)
    if kind == "clip":
        assert np.array_equal(out["labels"], ref["labels"])
    elif kind == "siglip":
        assert abs(out["d_bias"] - float(ref["d_bias"])) <= 1e-4 * max(abs(float(ref["d_bias"])), 1e-3)


@pytest.mark.parametrize("backend", ["emat", "gmat", "fused"])
@pytest.mark.parametrize("name", [n for n in golden_names() if "_w1" in n])
def test_single_rank_matches_reference(name, backend, standin_engine, monkeypatch):
    monkeypatch.setenv("MRCLIP_BWD", backend)
    case = load_golden(name)
    out = _run_rank(case, 0, 1, case["image"], case["text"])
    # G is rounded to bf16 in the gmat / emat stand-in, as in the kernels
    # (emat, one rank: d_scale = <dI, I>/scale inherits the bf16 rounding of G, which a 16-row fixture does not average out)
    bf16_g = backend != "fused" or case["meta"]["kind"] == "mpos"      # G rounded to bf16, as in the kernels
    _compare(out, case["ranks"][0], case["meta"]["kind"], grad_tol=5e-3 if bf16_g else GRAD_TOL,
             scale_tol=1e-2 if backend == "emat" else 1e-4)
    used = set(standin_engine.calls)
    if case["meta"]["kind"] == "mpos":      # always the E-block pipeline
        assert {"clip_fwd_tiles_e", "emat_to_gmat", "gmat_gemm"} <= used
        return
    assert ("gmat_gemm" in used) == (backend in ("gmat", "emat"))
    assert ("clip_bwd" in used or "siglip_bwd" in used) == (backend == "fused")
    assert ("clip_fwd_tiles_e" in used or "siglip_fwd_e" in used) == (backend == "emat")
    assert not ({"clip_gwrite", "siglip_gwrite"} & used) or backend == "gmat"


def test_default_backend_is_emat(standin_engine, monkeypatch):
    monkeypatch.delenv("MRCLIP_BWD", raising=False)
    case = load_golden("clip_w1_small")
    _run_rank(case, 0, 1, case["image"], case["text"])
    assert "clip_fwd_tiles_e" in standin_engine.calls and "emat_to_gmat" in standin_engine.calls


def test_no_grad_forward_keeps_nothing(standin_engine, monkeypatch):
    """inference: the plain forward runs (no E block is written) and the workspace returns to the pool"""
    monkeypatch.delenv("MRCLIP_BWD", raising=False)
    case = load_golden("clip_w1_small")
    mod = ClipLoss()
    with torch.no_grad():
        loss = mod(torch.from_numpy(case["image"]), torch.from_numpy(case["text"]), torch.tensor(float(case["meta"]["scale"])))
    assert abs(loss.item() - float(case["ranks"][0]["loss"])) <= LOSS_TOL * max(1.0, abs(float(case["ranks"][0]["loss"])))
    assert "clip_fwd_tiles" in standin_engine.calls and "clip_fwd_tiles_e" not in standin_engine.calls
    assert all(not w.in_use for lst in mod._pool._free.values() for w in lst)


def _dist_worker(rank, world, init_file, name, backend, ret):
    os.environ["MRCLIP_BWD"] = backend
    mrclip_b200.set_engine(StandInEngine())
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    try:
        case = load_golden(name)
        ret[rank] = _run_rank(case, rank, world, case["image"], case["text"])
        dist.barrier()
    finally:
        dist.destroy_process_group()


MULTI = [n for n in golden_names() if "_w1" not in n and "_w8" not in n]


@pytest.mark.parametrize("name", MULTI)
def test_multi_rank_gloo_matches_reference(name):
    case = load_golden(name)
    world = case["world"]
    # every case runs the default (emat: reduce-scatter of the text-gradient partials); the older two
    # orchestrations alternate over the cases
    backends = ["emat", "gmat" if (MULTI.index(name) % 2 == 0) else "fused"]
    if case["meta"]["kind"] == "mpos":
        backends = ["emat"]          # the multi-positive loss has one pipeline
    mgr = mp.Manager()
    ret = mgr.dict()
    for backend in backends:
        with tempfile.TemporaryDirectory() as td:
            mp.spawn(_dist_worker, args=(world, os.path.join(td, "init"), name, backend, ret), nprocs=world, join=True)
        for r in range(world):
            bf16_g = backend != "fused" or case["meta"]["kind"] == "mpos"
            _compare(ret[r], case["ranks"][r], case["meta"]["kind"], grad_tol=5e-3 if bf16_g else GRAD_TOL)


def _gather_worker(rank, world, init_file, ret):
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    try:
        torch.manual_seed(rank)
        res = {}
        for ll in (False, True):
            for gg in (False, True):
                a = torch.randn(3, 4, requires_grad=True)
                b = torch.randn(3, 4, requires_grad=True)
                ga, gb = mrclip_b200.gather_features(a, b, local_loss=ll, gather_with_grad=gg, rank=rank,
                                                     world_size=world)
                w = torch.arange(1, world * 3 + 1, dtype=torch.float32)[:, None]
                (ga * w).sum().backward() if ga.requires_grad else None
                res[(ll, gg)] = (ga.detach().numpy(), None if a.grad is None else a.grad.numpy())
        ret[rank] = res
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_gather_features_semantics_gloo():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    with tempfile.TemporaryDirectory() as td:
        mp.spawn(_gather_worker, args=(world, os.path.join(td, "init"), ret), nprocs=world, join=True)
    for r in range(world):
        for (ll, gg), (ga, grad) in ret[r].items():
            assert ga.shape == (6, 4)
            w_rows = np.arange(1, 7, dtype=np.float32)[r * 3:(r + 1) * 3, None]
            if gg:       # reduce-scatter of W identical weightings -> W x
                assert np.allclose(grad, world * np.broadcast_to(w_rows, (3, 4)))
            elif not ll:  # own slot re-inserted -> 1 x
                assert np.allclose(grad, np.broadcast_to(w_rows, (3, 4)))
            else:         # local loss without grad: gathered copy is detached
                assert grad is None


def _exchange_worker(rank, world, init_file, ret):
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    try:
        from mrclip_b200.loss import (neighbour_exchange, neighbour_exchange_bidir,
                                      neighbour_exchange_bidir_with_grad, neighbour_exchange_with_grad)
        left, right = (rank - 1) % world, (rank + 1) % world
        mine = torch.full((2, 3), float(rank))
        res = {"plain": neighbour_exchange(left, right, mine).numpy()}            # what the left neighbour sent
        fr, fl = neighbour_exchange_bidir(left, right, mine + 100, mine + 200)
        res["bidir"] = (fr.numpy(), fl.numpy())
        # gradients travel the hop backwards: what I sent right is weighted on the right neighbour
        x = torch.full((2, 3), 1.0, requires_grad=True)
        y = neighbour_exchange_with_grad(left, right, x)
        (y * float(rank + 1)).sum().backward()
        res["grad"] = x.grad.numpy()
        a = torch.ones(2, 3, requires_grad=True)
        b = torch.ones(2, 3, requires_grad=True)
        from_right, from_left = neighbour_exchange_bidir_with_grad(left, right, a, b)
        (from_right * float(10 * (rank + 1)) + from_left * float(rank + 1)).sum().backward()
        res["grad_bidir"] = (a.grad.numpy(), b.grad.numpy())
        ret[rank] = res
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_neighbour_exchange_semantics_gloo():
    """reference loss.py:226-311: ring hops, their return order, and the reverse hop as gradient"""
    world = 3
    mgr = mp.Manager()
    ret = mgr.dict()
    with tempfile.TemporaryDirectory() as td:
        mp.spawn(_exchange_worker, args=(world, os.path.join(td, "init"), ret), nprocs=world, join=True)
    for r in range(world):
        left, right = (r - 1) % world, (r + 1) % world
        res = ret[r]
        assert np.all(res["plain"] == left)
        fr, fl = res["bidir"]
        assert np.all(fr == right + 100) and np.all(fl == left + 200)     # right sent "to left", left sent "to right"
        assert np.all(res["grad"] == right + 1)                           # my tensor was weighted on the right neighbour
        ga, gb = res["grad_bidir"]
        assert np.all(ga == 10 * (left + 1))    # a went left, where it arrived as that rank's from_right
        assert np.all(gb == right + 1)          # b went right, where it arrived as that rank's from_left


# ------------------------------------------------------------------------------------------------
# materialised helpers kept for callers of the reference's module (not on the fused path)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["mpos_w1", "mpos_w1_delta03_ragged"])
def test_materialised_multipositive_helpers_match_reference_golden(name):
    """get_logits + multi_positive_cross_entropy_loss (loss.py:626-644, 696-747) reproduce the recorded outputs"""
    from mrclip_b200.loss import multi_positive_cross_entropy_loss
    case = load_golden(name)
    meta, ref = case["meta"], case["ranks"][0]
    img = torch.from_numpy(case["image"]).requires_grad_(True)
    txt = torch.from_numpy(case["text"]).requires_grad_(True)
    scale = torch.tensor(float(meta["scale"]), requires_grad=True)
    labels = torch.from_numpy(ref["labels_in"])
    mod = MultiPositiveClipLoss()
    li, lt, tok = mod.get_logits_custom(img, txt, labels, scale)
    assert tok is labels
    mask = torch.eq(labels[:, None], labels[None, :]).float()
    d = float(meta["delta"])
    loss = d * multi_positive_cross_entropy_loss(li, mask) + (1 - d) * multi_positive_cross_entropy_loss(lt, mask)
    (loss * float(meta["grad_output"])).backward()
    assert abs(loss.item() - float(ref["loss"])) <= 1e-5 * abs(float(ref["loss"]))
    assert rel_err(img.grad.numpy(), ref["d_image"]) <= 1e-5 and rel_err(txt.grad.numpy(), ref["d_text"]) <= 1e-5
    assert abs(scale.grad.item() - float(ref["d_scale"])) <= 1e-4 * abs(float(ref["d_scale"])) + 1e-7


def test_multi_positive_cross_entropy_quirks():
    from mrclip_b200.loss import multi_positive_cross_entropy_loss
    logits = torch.tensor([[2.0, 0.0, -1.0], [0.5, 0.5, 0.5]])
    mask = torch.tensor([[1.0, 0.0, 1.0], [0.0, 0.0, 0.0]])          # second row: no positives -> contributes 0 / 1
    lp = torch.log_softmax(logits.double(), dim=1)
    want = (-(lp[0, 0] + lp[0, 2]) / 2 + 0.0) / 2
    assert abs(multi_positive_cross_entropy_loss(logits, mask).item() - want.item()) < 1e-6


def test_siglip_materialised_chunk_loss_matches_reference_golden():
    """SigLipLoss._loss on one rank is the whole loss (loss.py:354-363, :365-368)"""
    case = load_golden("siglip_w1")
    meta, ref = case["meta"], case["ranks"][0]
    img, txt = torch.from_numpy(case["image"]), torch.from_numpy(case["text"])
    mod = SigLipLoss()
    loss = mod._loss(img, txt, torch.tensor(float(meta["scale"])), torch.tensor(float(meta["bias"])))
    assert abs(loss.item() - float(ref["loss"])) <= 1e-5 * abs(float(ref["loss"]))
    neg = mod._loss(img, txt, torch.tensor(float(meta["scale"])), torch.tensor(float(meta["bias"])), negative_only=True)
    z = float(meta["scale"]) * img.double() @ txt.double().t() + float(meta["bias"])
    assert abs(neg.item() - (torch.nn.functional.softplus(z).sum() / img.shape[0]).item()) <= 1e-5 * abs(neg.item())


def _tokens_worker(rank, world, init_file, ret):
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    try:
        from mrclip_b200.loss import gather_features_with_tokens, multi_positive_cross_entropy_loss
        case = load_golden("mpos_w2_ll1_gg1")
        meta, ref = case["meta"], case["ranks"][rank]
        n = case["image"].shape[0] // world
        img = torch.from_numpy(case["image"][rank * n:(rank + 1) * n]).requires_grad_(True)
        txt = torch.from_numpy(case["text"][rank * n:(rank + 1) * n]).requires_grad_(True)
        labels = torch.from_numpy(ref["labels_in"])
        gi, gt, tok = gather_features_with_tokens(img, txt, labels, local_loss=True, gather_with_grad=True, rank=rank,
                                                  world_size=world)
        mod = MultiPositiveClipLoss(local_loss=True, gather_with_grad=True, rank=rank, world_size=world)
        li, lt, tok2 = mod.get_logits_custom(img, txt, labels, torch.tensor(float(meta["scale"])))
        mask = torch.eq(labels[:, None], tok2[None, :]).float()
        d = float(meta["delta"])
        loss = d * multi_positive_cross_entropy_loss(li, mask) + (1 - d) * multi_positive_cross_entropy_loss(lt, mask)
        loss.backward()
        ret[rank] = dict(loss=loss.item(), tok=tok.numpy(), tok2=tok2.numpy(), gi=gi.detach().numpy(),
                         d_image=img.grad.numpy(), d_text=txt.grad.numpy())
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_gather_with_tokens_and_materialised_multipositive_gloo():
    world = 2
    case = load_golden("mpos_w2_ll1_gg1")
    mgr = mp.Manager()
    ret = mgr.dict()
    with tempfile.TemporaryDirectory() as td:
        mp.spawn(_tokens_worker, args=(world, os.path.join(td, "init"), ret), nprocs=world, join=True)
    all_labels = np.concatenate([case["ranks"][r]["labels_in"] for r in range(world)])
    for r in range(world):
        out, ref = ret[r], case["ranks"][r]
        assert np.array_equal(out["tok"], all_labels) and np.array_equal(out["tok2"], all_labels)   # rank order, bit-exact
        assert np.array_equal(out["gi"], case["image"])
        assert abs(out["loss"] - float(ref["loss"])) <= 1e-5 * abs(float(ref["loss"]))
        assert rel_err(out["d_image"], ref["d_image"]) <= 1e-5 and rel_err(out["d_text"], ref["d_text"]) <= 1e-5


# ------------------------------------------------------------------------------------------------
# MRCLIP_DS=fwd (opt-in): d logit_scale of the multi-rank local loss from forward-side row sums and <dT_r, T_r>
# ------------------------------------------------------------------------------------------------
def _fwd_ds_worker(rank, world, init_file, name, exact, ret):
    os.environ["MRCLIP_BWD"] = "emat"
    os.environ["MRCLIP_DS"] = "fwd"
    import mrclip_b200.loss as L
    L._FWD_DS_MIN_PAIRS = 0          # the fixtures are far below the production threshold
    eng = StandInEngine()
    eng.exact_g = exact
    mrclip_b200.set_engine(eng)
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    try:
        case = load_golden(name)
        out = _run_rank(case, rank, world, case["image"], case["text"])
        out["calls"] = list(eng.calls)
        ret[rank] = out
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name,exact", [("clip_w2_ll1_gg1", True), ("clip_w4_ll1_gg1", True), ("clip_w4_ll1_gg1", False),
                                        ("clip_w2_ll0_gg1", True)])
def test_forward_side_d_scale_matches_reference(name, exact):
    """s dL_r/ds = <dT_r, T_r> + ln2/(2n) (R2(r,*) - R2(*,r)) reproduces the reference's per-rank d logit_scale; with
    G rounded to bf16 (exact=False) only up to the noise a 16..32-row fixture cannot average out."""
    case = load_golden(name)
    world = case["world"]
    mgr = mp.Manager()
    ret = mgr.dict()
    with tempfile.TemporaryDirectory() as td:
        mp.spawn(_fwd_ds_worker, args=(world, os.path.join(td, "init"), name, exact, ret), nprocs=world, join=True)
    local = bool(case["meta"]["local_loss"])
    for r in range(world):
        _compare(ret[r], case["ranks"][r], "clip", grad_tol=GRAD_TOL if exact else 5e-3,
                 scale_tol=1e-4 if (exact or not local) else 5e-2)
        used = ret[r]["calls"]
        assert ("clip_fwd_tiles_eu" in used and "row_ent_split" in used) == local      # global-loss modes keep the entropy path
        assert ("clip_fwd_tiles_e" in used) == (not local)


# ---- workspace leases: no leak without backward, no silent second backward, bounded pool --------
def _feat(n=16, d=8, seed=0):
    g = torch.Generator().manual_seed(seed)
    return (torch.nn.functional.normalize(torch.randn(n, d, generator=g), dim=-1).requires_grad_(True),
            torch.nn.functional.normalize(torch.randn(n, d, generator=g), dim=-1).requires_grad_(True))


def test_workspace_returns_when_graph_is_dropped(standin_engine):
    mod = ClipLoss()
    i, t = _feat()
    s = torch.tensor(10.0, requires_grad=True)
    for _ in range(6):              # forward only, graph dropped every time: the same set must be reused
        loss = mod(i, t, s)
        del loss
    sets = [w for lst in mod._pool._free.values() for w in lst]
    assert len(sets) == 1 and not sets[0].in_use
    loss = mod(i, t, s)             # a live graph holds its set ...
    assert sets[0].in_use
    loss.backward()                 # ... and backward returns it
    assert not sets[0].in_use


def test_second_backward_raises(standin_engine):
    mod = ClipLoss()
    i, t = _feat()
    s = torch.tensor(10.0, requires_grad=True)
    loss = mod(i, t, s)
    loss.backward(retain_graph=True)
    with pytest.raises(RuntimeError, match="second time"):
        loss.backward()


def test_returned_loss_may_be_modified_in_place(standin_engine):
    i, t = _feat(seed=3)
    s = torch.tensor(10.0, requires_grad=True)
    ref = ClipLoss()(i, t, s)
    ref.backward()
    want = s.grad.clone()
    s.grad = None
    loss = ClipLoss()(i, t, s)
    with torch.no_grad():
        loss.mul_(0.0)              # e.g. loss /= accum_freq on the returned tensor
    loss.backward()
    torch.testing.assert_close(s.grad, want)


def test_pool_evicts_idle_keys(standin_engine, monkeypatch):
    monkeypatch.setenv("MRCLIP_POOL_KEYS", "3")
    mod = ClipLoss()
    s = torch.tensor(10.0)
    for n in (8, 9, 10, 11, 12, 13):          # ragged batch sizes
        i, t = _feat(n=n)
        with torch.no_grad():
            mod(i, t, s)
    assert len(mod._pool._free) <= 3
    assert list(mod._pool._free)[-1][1] == 13


def test_cpu_tensors_raise_without_engine_override():
    mrclip_b200.set_engine(None)
    i, t = _feat()
    with pytest.raises(RuntimeError, match="no CPU path"):
        ClipLoss()(i, t, torch.tensor(10.0))


# ---- feature hand-off (forward_raw) on the CPU path: exactly the reference composition -------------------------------
def test_forward_raw_falls_back_to_the_reference_composition(standin_engine):
    """Without the whole-step path (stand-in engine) forward_raw is forward(F.normalize(i), F.normalize(t), exp(log_scale))
    -- model.py:282-301, :324 in front of loss.py:128-139 -- with gradients reaching the raw rows and the log-scale."""
    import torch_ref
    g = torch.Generator().manual_seed(5)
    raw_i = (torch.randn(24, 16, generator=g) * 3).requires_grad_(True)
    raw_t = (torch.randn(24, 16, generator=g) * 0.2).requires_grad_(True)
    ls = torch.tensor(2.0, requires_grad=True)
    out = ClipLoss().forward_raw(raw_i, raw_t, ls, output_dict=True)
    assert list(out) == ["contrastive_loss"]
    out["contrastive_loss"].backward()
    ref = torch_ref.clip_reference(raw_i, raw_t, 2.0, raw=True)
    assert abs(float(out["contrastive_loss"]) - float(ref["loss"])) <= 2e-3 * abs(float(ref["loss"]))
    assert rel_err(raw_i.grad.numpy(), ref["d_image"].numpy()) <= 1e-2        # (features rounded to bf16 by the engine)
    assert rel_err(raw_t.grad.numpy(), ref["d_text"].numpy()) <= 1e-2
    assert abs(float(ls.grad) - float(ref["d_scale"])) <= 1e-2 * abs(float(ref["d_scale"]))
    b = torch.tensor(-3.0, requires_grad=True)
    loss = SigLipLoss().forward_raw(raw_i.detach(), raw_t.detach(), ls.detach(), b)
    ref = torch_ref.siglip_reference(raw_i, raw_t, 2.0, -3.0, raw=True)
    assert abs(float(loss) - float(ref["loss"])) <= 2e-3 * abs(float(ref["loss"]))


# ---- the peer-memory / NCCL choice is made by all ranks together ------------------------------------
def _agree_worker(rank, world, init_file, ret):
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    try:
        from mrclip_b200.loss import _all_ranks_ok, _symmetric_alloc
        dev = torch.device("cpu")
        res = [_all_ranks_ok(True, dev), _all_ranks_ok(rank != 1, dev)]          # one rank fails -> nobody proceeds
        # a module whose world_size is not the default group's must not build peer tables on that group
        res.append(_symmetric_alloc([((4,), torch.float32)], dev, world + 1, "test") is None)
        ret[rank] = res
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_symmetric_memory_decision_is_collective():
    import warnings
    mgr = mp.Manager()
    ret = mgr.dict()
    with tempfile.TemporaryDirectory() as td, warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mp.spawn(_agree_worker, args=(3, os.path.join(td, "init"), ret), nprocs=3, join=True)
    for r in range(3):
        assert ret[r] == [True, False, True]
