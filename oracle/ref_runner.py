"""Times the reference's own ``ClipLoss`` / ``SigLipLoss`` on the host cores  --  TEST / BASELINE INFRASTRUCTURE ONLY.

``oracle/_ref/src/open_clip/loss.py`` is the unmodified reference file (copied by ``oracle/make_ref.py``; the sha256
manifest beside it says so).  It is loaded by path -- importing the ``open_clip`` package would pull ``ftfy`` through
the tokenizer (tokenizer.py:14), which this image does not have -- and driven through its public API:

* ``time_single(N, D)``: ``ClipLoss(world_size=1)`` forward + ``.backward()`` on N pairs (BASELINE config c1 is
  N=256, D=512), every host thread, median over the timed steps.
* ``time_rank_of_job(N, D, world)``: one rank of the reference's distributed job, for problem sizes whose full CPU
  step would take minutes (N=32768: ~10 TFLOP of fp32).  A real gloo group of ``world`` processes is started; rank 0
  runs ``ClipLoss(local_loss=True, gather_with_grad=True, rank=0, world_size=world)`` forward + backward on its
  n = N/world pairs against all N candidates with every host thread, while the other ranks only take part in the
  collectives (the reference's ``gather_features`` and its backward) on one thread each.  Per-pair work equals the full
  job's, so pairs/s = n / (rank 0's step time) is the whole-job rate the host cores sustain.

When ``oracle/_ref`` is absent the callers fall back to the port (``oracle/clip_port.py``, kind "port").
"""
from __future__ import annotations

import importlib.util
import os
import statistics
import tempfile
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_LOSS = os.path.join(HERE, "_ref", "src", "open_clip", "loss.py")


def available():
    return os.path.exists(REF_LOSS)


def load_ref_loss():
    spec = importlib.util.spec_from_file_location("mrclip_reference_loss", REF_LOSS)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def features(n_total, d, seed=1237):
    g = torch.Generator().manual_seed(seed)
    img = torch.nn.functional.normalize(torch.randn(n_total, d, generator=g), dim=-1).bfloat16().float()
    txt = torch.nn.functional.normalize(0.5 * img + 0.5 * torch.randn(n_total, d, generator=g) / d ** 0.5,
                                        dim=-1).bfloat16().float()
    return img, txt


def _summary(times, pairs, cores, sample):
    sec = statistics.median(times)
    return dict(pairs_per_s=pairs / sec, seconds_per_step=sec, seconds_total=sum(times), steps=len(times), cores=cores,
                sample=sample, kind="reference")


def time_single(n_total, d, steps=10, warmup=3, kind="clip", scale=14.285714, bias=-10.0, threads=None):
    """The reference module with world_size=1 on all n_total pairs; every step is a full forward + backward."""
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    ref = load_ref_loss()
    img, txt = features(n_total, d)
    mod = ref.ClipLoss() if kind == "clip" else ref.SigLipLoss()
    times = []
    for k in range(warmup + steps):
        i = img.clone().requires_grad_(True)
        t = txt.clone().requires_grad_(True)
        s = torch.tensor(scale, requires_grad=True)
        t0 = time.perf_counter()
        if kind == "clip":
            loss = mod(i, t, s)
        else:
            loss = mod(i, t, s, torch.tensor(bias, requires_grad=True))
        loss.backward()
        if k >= warmup:
            times.append(time.perf_counter() - t0)
    name = "ClipLoss" if kind == "clip" else "SigLipLoss"
    return _summary(times, n_total, threads,
                    f"reference {name}(world_size=1) (oracle/_ref, unmodified), all {n_total} pairs, D={d}, fp32, fwd+bwd, "
                    f"median of {steps} after {warmup} warm-ups")


def _job_worker(rank, world, init_file, n_total, d, steps, warmup, threads, kind, scale, bias, ret):
    import torch.distributed as dist
    torch.set_num_threads(threads if rank == 0 else 1)
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    try:
        ref = load_ref_loss()
        n = n_total // world
        img, txt = features(n_total, d)
        img, txt = img[rank * n:(rank + 1) * n].clone(), txt[rank * n:(rank + 1) * n].clone()
        if kind == "clip":
            mod = ref.ClipLoss(local_loss=True, gather_with_grad=True, rank=rank, world_size=world)
        else:
            mod = ref.SigLipLoss(rank=rank, world_size=world, dist_impl="gather")
        times = []
        for k in range(warmup + steps):
            i = img.clone().requires_grad_(True)
            t = txt.clone().requires_grad_(True)
            s = torch.tensor(scale, requires_grad=True)
            dist.barrier()
            t0 = time.perf_counter()
            if rank == 0:
                loss = mod(i, t, s) if kind == "clip" else mod(i, t, s, torch.tensor(bias, requires_grad=True))
                loss.backward()
            elif kind == "clip":
                # serve the collectives of rank 0's step: the reference's own gather and its backward, no logits
                ai, at = ref.gather_features(i, t, local_loss=True, gather_with_grad=True, rank=rank, world_size=world)
                (ai.sum() * 0.0 + at.sum() * 0.0).backward()
            else:
                at = torch.cat(torch.distributed.nn.all_gather(t))     # loss.py:436 (dist_impl="gather")
                (at.sum() * 0.0).backward()
            if k >= warmup:
                times.append(time.perf_counter() - t0)
        if rank == 0:
            ret["times"] = times
        dist.barrier()
    finally:
        dist.destroy_process_group()


def time_rank_of_job(n_total, d, world=8, steps=5, warmup=2, kind="clip", scale=14.285714, bias=-10.0, threads=None):
    import torch.distributed.nn  # noqa: F401
    import torch.multiprocessing as mp
    threads = threads or os.cpu_count() or 1
    ret = mp.Manager().dict()
    with tempfile.TemporaryDirectory() as td:
        mp.spawn(_job_worker, args=(world, os.path.join(td, "init"), n_total, d, steps, warmup, threads, kind, scale, bias, ret),
                 nprocs=world, join=True)
    n = n_total // world
    name = "ClipLoss(local_loss=True, gather_with_grad=True" if kind == "clip" else "SigLipLoss(dist_impl='gather'"
    return _summary(list(ret["times"]), n, threads,
                    f"rank 0 of a {world}-rank gloo job running the reference {name}, world_size={world}) (oracle/_ref, "
                    f"unmodified): {n} of {n_total} pairs vs all {n_total} candidates, D={d}, fp32, fwd+bwd, {threads} "
                    f"threads on rank 0 (the other ranks only serve the collectives), median of {steps} after {warmup} warm-ups")


def time_reference(n_total, d, steps, warmup, kind="clip", scale=14.285714, bias=-10.0, budget_flop=3.0e12):
    """Picks the bounded sample: the whole problem on one process while a step stays under ``budget_flop`` of fp32 work
    (12 n N D executed by the reference per rank), one rank of an 8-rank job (or 16, 32 ...) otherwise."""
    world = 1
    while 12.0 * (n_total / world) * n_total * d > budget_flop and world < 64 and n_total % (world * 2) == 0:
        world *= 2
    if world == 1:
        return time_single(n_total, d, steps, warmup, kind, scale, bias)
    return time_rank_of_job(n_total, d, world, steps, warmup, kind, scale, bias)
