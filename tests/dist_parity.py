"""Production-size parity worker: the public modules against the reference's fp32 torch graph, per rank, on real GPUs.

    python tests/dist_parity.py [case ...]                                   # one GPU
    torchrun --nproc-per-node W tests/dist_parity.py [case ...]             # W GPUs, NCCL + NVLink peer memory

Every rank feeds the same bf16-rounded synthetic features to (a) ``mrclip_b200`` and (b)
``tests/torch_ref.py`` -- the reference's operators in fp32 with ``torch.distributed.nn.all_gather`` -- and compares
loss (1e-3), feature gradients (1e-2, norm-wise), d logit_scale / d logit_bias (1e-2) and, for ClipLoss, the labels
(bit-exact).  Cases are BASELINE.json's configs at full size:

    c3       ClipLoss (T,T)            N=32768 D=768      (configs[2], the headline)
    c2       ClipLoss all four modes   N=4096  D=512      (configs[1]; (F,T) and (T,T) are the ones it names)
    c2raw    ClipLoss.forward_raw      N=4096  D=512      (un-normalised fp32 features, log-scale)
    c4raw    SigLipLoss.forward_raw    N=4096  D=768
    c4       SigLipLoss + logit_bias   N=16384 D=768      (configs[3])
    mpos     MultiPositiveClipLoss     N=8192  D=512      (512 label classes)
    ragged   ClipLoss (T,T)            n=1000 per rank, D=200 (no tile / vector alignment anywhere)
    exchange neighbour_exchange*       n=512 per rank: hop results and gradients exact; the reference's "shift" SigLIP
                                       loop (loss.py:404-421) built on them equals the gathered form (several ranks only)

and each is run over NVLink peer memory (default) and over NCCL collectives (``-nccl``).  Exit code 1 on any mismatch;
one line per case / backend with the worst error over ranks.
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch_ref  # noqa: E402
from mrclip_b200 import ClipLoss, MultiPositiveClipLoss, SigLipLoss  # noqa: E402

CASES = {
    "c3": dict(kind="clip", N=32768, D=768, scale=14.285714, modes=[(True, True)]),
    "c2": dict(kind="clip", N=4096, D=512, scale=14.285714, modes=[(False, True), (True, True), (True, False), (False, False)]),
    # (weakly correlated pairs: at scale 100 the usual 0.5 mix saturates the loss to ~1e-30 and every gradient to 0)
    "c2s100": dict(kind="clip", N=4096, D=512, scale=100.0, modes=[(True, True)], grad_output=65536.0, corr=0.06),
    "c4": dict(kind="siglip", N=16384, D=768, scale=10.0, bias=-10.0),
    # feature hand-off fusion (forward_raw): un-normalised fp32 tower outputs + log-scale parameter
    "c2raw": dict(kind="clip", N=4096, D=512, scale=2.659, modes=[(True, True), (False, True)], raw=True),
    "c4raw": dict(kind="siglip", N=4096, D=768, scale=2.3026, bias=-10.0, raw=True),
    "mpos": dict(kind="mpos", N=8192, D=512, scale=14.285714, classes=512, delta=0.3),
    # the ring hops (neighbour_exchange*, loss.py:226-311) on NCCL + the reference's "shift" SigLIP built on them
    "exchange": dict(kind="exchange", n=512, D=256, scale=10.0, bias=-10.0),
    "ragged": dict(kind="clip", n=1000, D=200, scale=30.0, modes=[(True, True), (False, True)], corr=0.2),
}


def features(N, D, seed, device, corr=0.5):
    g = torch.Generator().manual_seed(seed)
    img = torch.nn.functional.normalize(torch.randn(N, D, generator=g), dim=-1)
    txt = torch.nn.functional.normalize(corr * img + (1.0 - corr) * torch.randn(N, D, generator=g) / D ** 0.5, dim=-1)
    return img.bfloat16().to(device), txt.bfloat16().to(device)


def exchange_case(c, rank, world, dev):
    """The ring hops on the real process group: values and gradients are exact; the "shift" SigLIP loop of the reference
    (loss.py:404-421) written with them gives the loss / gradients of the gathered form (torch_ref.siglip_reference)."""
    import torch.nn.functional as F
    from mrclip_b200 import (neighbour_exchange, neighbour_exchange_bidir, neighbour_exchange_bidir_with_grad,
                             neighbour_exchange_with_grad)
    left, right = (rank - 1) % world, (rank + 1) % world
    bad = []
    x = torch.full((4, 8), float(rank), device=dev)
    if not torch.equal(neighbour_exchange(left, right, x), torch.full_like(x, float(left))):
        bad.append("neighbour_exchange")
    # (two ranks: left == right and the reference never calls the two-sided hop -- (W - 1) // 2 == 0, loss.py:372)
    fr, fl = neighbour_exchange_bidir(left, right, x + 0.25, x + 0.5) if world > 2 else (x + 0.25 - rank + right, x + 0.5 - rank + left)
    if not (torch.equal(fr, torch.full_like(x, right + 0.25)) and torch.equal(fl, torch.full_like(x, left + 0.5))):
        bad.append("neighbour_exchange_bidir")
    a = x.clone().requires_grad_(True)
    (neighbour_exchange_with_grad(left, right, a) * (rank + 1)).sum().backward()
    if not torch.equal(a.grad, torch.full_like(x, float(right + 1))):       # my rows were weighted by the rank to my right
        bad.append("neighbour_exchange_with_grad")
    if world > 2:
        a, b = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
        fr, fl = neighbour_exchange_bidir_with_grad(left, right, a, b)
        (fr * (rank + 1) + fl * 10 * (rank + 1)).sum().backward()
        # a went left: the left rank received it as its "from right" (weight left+1); b went right: weight 10*(right+1)
        if not (torch.equal(a.grad, torch.full_like(x, float(left + 1))) and torch.equal(b.grad, torch.full_like(x, 10.0 * (right + 1)))):
            bad.append("neighbour_exchange_bidir_with_grad")
    # the reference's shift loop on top of the hops
    n, D = c["n"], c["D"]
    img, txt = features(n * world, D, 4321, dev)
    rows = slice(rank * n, (rank + 1) * n)
    i = img[rows].float().clone().requires_grad_(True)
    t = txt[rows].float().clone().requires_grad_(True)
    s = torch.tensor(c["scale"], device=dev, requires_grad=True)
    bz = torch.tensor(c["bias"], device=dev, requires_grad=True)

    def chunk_loss(tf, negative_only):
        logits = s * i @ tf.T + bz
        labels = -torch.ones_like(logits)
        if not negative_only:
            labels = labels + 2 * torch.eye(n, device=dev)
        return -F.logsigmoid(labels * logits).sum() / n
    loss = chunk_loss(t, False)
    to_right = t
    for _ in range(world - 1):
        from_left = neighbour_exchange_with_grad(left, right, to_right)
        loss = loss + chunk_loss(from_left, True)
        to_right = from_left
    loss.backward()
    ref = torch_ref.siglip_reference(img[rows], txt[rows], c["scale"], c["bias"], rank, world)
    errs, bad2 = torch_ref.compare(dict(loss=loss.detach(), d_image=i.grad, d_text=t.grad, d_scale=s.grad, d_bias=bz.grad), ref,
                                   tol_loss=1e-5, tol_grad=1e-4)
    return bad + bad2, errs


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cuda.matmul.allow_tf32 = False
    names = [a for a in sys.argv[1:] if not a.startswith("-")] or list(CASES)
    # push: peer memory, copy-engine all-gather for shares >= 4 MB; push-ce: the same with the copy engines forced on
    # every shape (small cases only); nccl: NCCL collectives through the Python orchestration
    transports = ["push", "push-ce", "nccl"] if world > 1 else ["local"]
    failures = 0
    for name in names:
        c = CASES[name]
        if c["kind"] == "exchange":
            if world == 1:
                continue
            bad, errs = exchange_case(c, rank, world, dev)
            flag = torch.tensor([float(len(bad))], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
            failures += int(flag.item() > 0)
            if bad:
                print(f"MISMATCH rank {rank} exchange: {bad} {errs}", flush=True)
            if rank == 0:
                print(f"exchange W={world}: ring hops exact, shift-SigLIP on the hops vs gathered form: " +
                      " ".join(f"{k}={v:.2e}" for k, v in errs.items()) + f"  [{'ok' if flag.item() == 0 else 'FAIL'}]", flush=True)
            continue
        N = c["N"] if "N" in c else c["n"] * world
        if N % world:
            continue
        n, D = N // world, c["D"]
        img, txt = features(N, D, 1234 + len(name) + N, dev, c.get("corr", 0.5))
        raw = bool(c.get("raw", False))
        if raw:     # un-normalised rows of varying length, fp32 (what a tower hands over under AMP)
            g = torch.Generator().manual_seed(99)
            img = img.float() * (0.5 + 4.0 * torch.rand(N, 1, generator=g).to(dev))
            txt = txt.float() * (0.5 + 4.0 * torch.rand(N, 1, generator=g).to(dev))
        rows = slice(rank * n, (rank + 1) * n)
        go = float(c.get("grad_output", 1.0))
        modes = c.get("modes", [(True, True)])
        for mode in modes:
            ref = None
            for tr in transports:
                if tr == "push-ce" and N > 8192:
                    continue
                os.environ["MRCLIP_RS"] = os.environ["MRCLIP_AG"] = ("nccl" if tr == "nccl" else "push")
                os.environ["MRCLIP_AG_OVERLAP_MIN_BYTES"] = "0" if tr == "push-ce" else str(4 << 20)
                i = img[rows].clone().requires_grad_(True)
                t = txt[rows].clone().requires_grad_(True)
                s = torch.tensor(c["scale"], device=dev, requires_grad=True)
                if c["kind"] == "clip":
                    ll, gg = mode
                    mod = ClipLoss(local_loss=ll, gather_with_grad=gg, cache_labels=True, rank=rank, world_size=world)
                    loss = mod.forward_raw(i, t, s) if raw else mod(i, t, s)
                    if ref is None:
                        ref = torch_ref.clip_reference(img[rows], txt[rows], c["scale"], ll, gg, rank, world, go, raw=raw)
                    nl = n if (world > 1 and ll) else N
                    labels_ok = torch.equal(mod.get_ground_truth(dev, nl), ref["labels"])
                elif c["kind"] == "siglip":
                    b = torch.tensor(c["bias"], device=dev, requires_grad=True)
                    sl = SigLipLoss(rank=rank, world_size=world)
                    loss = sl.forward_raw(i, t, s, b) if raw else sl(i, t, s, b)
                    if ref is None:
                        ref = torch_ref.siglip_reference(img[rows], txt[rows], c["scale"], c["bias"], rank, world, go, raw=raw)
                    labels_ok = True
                else:
                    lab = torch.randint(0, c["classes"], (N,), generator=torch.Generator().manual_seed(7)).to(dev)
                    mod = MultiPositiveClipLoss(local_loss=True, gather_with_grad=True, rank=rank, world_size=world)
                    loss = mod(i, t, s, delta=c["delta"], tokenized_texts=lab[rows])
                    if ref is None:
                        ref = torch_ref.mpos_reference(img[rows], txt[rows], c["scale"], lab[rows], c["delta"], rank, world, go)
                    labels_ok = True
                (loss * go).backward()
                if c["kind"] == "clip" and not raw:       # evaluation: the no-grad forward (no E block) gives the same loss
                    with torch.no_grad():
                        if abs(float(mod(i, t, s)) - float(loss)) > 1e-5 * abs(float(loss)):
                            labels_ok = False
                ours = dict(loss=loss.detach(), d_image=i.grad, d_text=t.grad, d_scale=s.grad)
                if c["kind"] == "siglip":
                    ours["d_bias"] = b.grad
                errs, bad = torch_ref.compare(ours, ref)
                torch.cuda.synchronize()
                if not labels_ok:
                    bad.append("labels")
                vals = torch.tensor([errs.get(k, 0.0) for k in ("loss", "d_image", "d_text", "d_scale", "d_bias")] +
                                    [float(len(bad))], device=dev, dtype=torch.float64)
                if world > 1:
                    dist.all_reduce(vals, op=dist.ReduceOp.MAX)
                failures += int(vals[-1].item() > 0)
                if bad:
                    print(f"MISMATCH rank {rank} {name} {mode} {tr}: {bad} {errs}", flush=True)
                if rank == 0:
                    tag = f"{name} W={world} N={N} D={D}" + (f" (ll={int(mode[0])},gg={int(mode[1])})" if c["kind"] == "clip" else "")
                    print(f"{tag:44s} {tr:7s} loss={float(ours['loss']):.6f} worst over ranks: loss={vals[0]:.2e} dI={vals[1]:.2e} "
                          f"dT={vals[2]:.2e} ds={vals[3]:.2e}" + (f" db={vals[4]:.2e}" if c["kind"] == "siglip" else "") +
                          f" labels={'exact' if labels_ok else 'WRONG'}  [{'ok' if vals[-1].item() == 0 else 'FAIL'}]", flush=True)
            del ref
            torch.cuda.empty_cache()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print("dist_parity:", "FAILED" if failures else "all green", flush=True)
    sys.exit(1 if failures else 0)


if __name__ == "__main__":
    main()
