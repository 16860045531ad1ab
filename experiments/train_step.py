#!/usr/bin/env python
"""SURVEY.md 8f N2 / BASELINE config c5: the full MR-CLIP train step with the loss dropped in.

    python experiments/train_step.py [--batch-per-gpu 1024] [--steps 6] [--loss ours,reference,none]
    torchrun --nproc-per-node W experiments/train_step.py ...        # DDP, global batch = W * batch-per-gpu

Everything around the loss is the UNMODIFIED reference, loaded from ``oracle/_ref/src`` (the byte-for-byte copy made by
``oracle/make_ref.py``; /root/reference does not exist on the GPU box): ``open_clip.create_model("ViT-B-16")`` (ViT-B/16
on 224x224 slices + the 12-layer, 98-token text tower of ``model_configs/ViT-B-16.json``, random init), DDP,
``train_one_epoch`` (train.py:70-269) with bf16 autocast, AdamW and the ``logit_scale`` clamp.  Only ``ftfy`` (absent in
this image, used by the tokenizer alone) is stubbed and the data are synthetic (``data.py:506-553`` does the same).
The loss module is swapped between

    reference   open_clip.loss.ClipLoss / MultiPositiveClipLoss   (loss.py:68-139, :671-747)
    ours        mrclip_b200.ClipLoss / MultiPositiveClipLoss      (same constructor arguments)
    none        a stand-in that only touches the features          (isolates the towers + optimizer time)

from identical initial weights.  Reported: ms per step for each, the share of the step the loss takes
((t - t_none) / t), and -- the drop-in check -- that ``ours`` and ``reference`` produce the same losses and parameter
updates (bf16 tolerance).  One JSON line on stdout, rank 0.
"""
import argparse
import copy
import json
import math
import os
import sys
import time
import types

import torch
import torch.distributed as dist
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF_SRC = os.path.join(ROOT, "oracle", "_ref", "src")


def load_reference():
    if not os.path.isdir(os.path.join(REF_SRC, "open_clip_train")):
        raise SystemExit("oracle/_ref is missing: run `python oracle/make_ref.py` in the build container")
    if "ftfy" not in sys.modules:
        stub = types.ModuleType("ftfy")
        stub.fix_text = lambda s: s
        sys.modules["ftfy"] = stub
    sys.path.insert(0, REF_SRC)
    import open_clip
    import open_clip.loss as ref_loss
    import open_clip_train.train as ref_train
    return open_clip, ref_loss, ref_train


class NullLoss(nn.Module):
    """Keeps the graph alive through both towers and logit_scale at (almost) no cost."""

    def forward(self, image_features, text_features, logit_scale, output_dict=False, **_):
        loss = (image_features.float().mean() + text_features.float().mean()) * 0.0 + logit_scale * 0.0
        return {"contrastive_loss": loss} if output_dict else loss


class Loader(list):
    pass


class TrainData:
    def __init__(self, batches, num_samples):
        self.dataloader = Loader(batches)
        self.dataloader.num_batches = len(batches)
        self.dataloader.num_samples = num_samples

    def set_epoch(self, epoch):
        pass


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="ViT-B-16")
    ap.add_argument("--batch-per-gpu", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--loss", default="ours,reference,none")
    ap.add_argument("--precision", default="amp_bf16")
    ap.add_argument("--multipositive", action="store_true", help="MultiPositiveClipLoss with 64 label classes (train.py:123)")
    ap.add_argument("--grad-checkpointing", action="store_true")
    ap.add_argument("--classes", type=int, default=64)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    use_cuda = torch.cuda.is_available()
    dev = torch.device("cuda", local) if use_cuda else torch.device("cpu")
    if use_cuda:
        torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl" if use_cuda else "gloo", **({"device_id": dev} if use_cuda else {}))
    open_clip, ref_loss, ref_train = load_reference()
    import mrclip_b200

    torch.manual_seed(0)
    init = open_clip.create_model(args.model, pretrained=None, precision=args.precision, device=dev, output_dict=True)
    if args.grad_checkpointing:
        init.set_grad_checkpointing()
    B = args.batch_per_gpu
    ctx_len = init.context_length
    vocab = init.vocab_size
    size = init.visual.image_size if isinstance(init.visual.image_size, int) else init.visual.image_size[0]
    g = torch.Generator().manual_seed(1000 + rank)
    n_batches = args.warmup + args.steps
    batches = [(torch.randn(B, 3, size, size, generator=g), torch.randint(1, vocab, (B, ctx_len), generator=g),
                torch.randint(0, args.classes, (B,), generator=g)) for _ in range(min(n_batches, 3))]
    batches = [batches[k % len(batches)] for k in range(n_batches)]
    targs = argparse.Namespace(device=str(dev), precision=args.precision, distill=False, accum_freq=1, freeze=False,
                               skip_scheduler=True, distance=False, multipositiveloss=args.multipositive, delta=0.5,
                               horovod=False, grad_clip_norm=None, log_every_n_steps=10 ** 9, world_size=world,
                               batch_size=B, wandb=False, rank=rank, local_rank=local)

    def make_loss(kind):
        kw = dict(local_loss=True, gather_with_grad=True, cache_labels=True, rank=rank, world_size=world)
        if kind == "none":
            return NullLoss()
        mod = ref_loss if kind == "reference" else mrclip_b200
        return (mod.MultiPositiveClipLoss if args.multipositive else mod.ClipLoss)(**kw)

    results, finals, seen_losses = {}, {}, {}
    for kind in [k for k in args.loss.split(",") if k]:
        model = copy.deepcopy(init)
        if world > 1:
            model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local] if use_cuda else None,
                                                              static_graph=False)
        opt = torch.optim.AdamW(model.parameters(), lr=1e-4, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.2)
        loss_mod = make_loss(kind)
        seen = []
        hook = loss_mod.register_forward_hook(lambda _m, _i, out: seen.append(sum(v.detach().float() for v in out.values())))
        loss_ms = []
        if use_cuda and kind != "none":      # device time of the loss forward (events around the module call)
            def pre(_m, _i):
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                loss_ms.append([e])

            def post(_m, _i, _o):
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                loss_ms[-1].append(e)
            loss_mod.register_forward_pre_hook(pre)
            loss_mod.register_forward_hook(post)

        def run(bs):
            ref_train.train_one_epoch(model, {"train": TrainData(bs, len(bs) * B * world)}, loss_mod, 0, opt, None, None,
                                      None, targs)
        run(batches[:args.warmup])
        if use_cuda:
            torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        if use_cuda:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        run(batches[args.warmup:])
        if use_cuda:
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
        else:
            ms = (time.perf_counter() - t0) * 1e3 / args.steps
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        results[kind] = float(t.item())
        hook.remove()
        seen_losses[kind] = [float(v) for v in seen]
        if loss_ms:
            fw = [a.elapsed_time(b) for a, b in (p for p in loss_ms if len(p) == 2)][args.warmup:]
            results[kind + "_loss_forward_ms"] = sum(fw) / max(len(fw), 1)
        m = model.module if world > 1 else model
        finals[kind] = {k: v.detach().float().clone() for k, v in m.named_parameters()}
        if use_cuda:
            results[kind + "_peak_mem_gb"] = torch.cuda.max_memory_allocated() / 2 ** 30
            torch.cuda.reset_peak_memory_stats()
        del model, opt, loss_mod
        if use_cuda:
            torch.cuda.empty_cache()

    out = {"workload": f"MR-CLIP train step (reference train_one_epoch + {args.model} towers, {args.precision}, AdamW), "
                       f"{'MultiPositiveClipLoss' if args.multipositive else 'ClipLoss'} local_loss gather_with_grad, "
                       f"batch {B} per GPU x {world} GPUs = global batch {B * world}",
           "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": results,
           "samples_per_s": {k: B * world / (v * 1e-3) for k, v in results.items() if k in ("ours", "reference", "none")}}
    if "none" in results:
        for k in ("ours", "reference"):
            if k in results:
                out[f"loss_share_of_step_{k}"] = max(results[k] - results["none"], 0.0) / results[k]
    if "ours" in finals and "reference" in finals:
        start = {k: v.detach().float() for k, v in init.named_parameters()}
        num = den = 0.0
        worst = ("", 0.0)
        for k in finals["ours"]:
            du, dr = finals["ours"][k] - start[k], finals["reference"][k] - start[k]
            num += float((du - dr).double().pow(2).sum())
            den += float(dr.double().pow(2).sum())
        out["param_update_rel_err_ours_vs_reference"] = math.sqrt(num / max(den, 1e-30))
        lo, lr_ = seen_losses["ours"], seen_losses["reference"]
        out["loss_first_step"] = {"ours": lo[0], "reference": lr_[0]}
        out["loss_last_step"] = {"ours": lo[-1], "reference": lr_[-1]}
        out["loss_rel_err_first_step"] = abs(lo[0] - lr_[0]) / max(abs(lr_[0]), 1e-30)
    if rank == 0:
        line = json.dumps(out)
        print(line, flush=True)
        if args.out:
            with open(args.out, "w") as f:
                f.write(line + "\n")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
