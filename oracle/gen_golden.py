"""Generate golden vectors by running the UNMODIFIED reference loss module.

Run in the build container only (it reads /root/reference, which does not exist on the GPU box):

    python oracle/gen_golden.py            # writes tests/golden/*.npz

``/root/reference/src/open_clip/loss.py`` imports only torch (loss.py:1-18), so it is loaded by
path with importlib; multi-rank cases run under torch.multiprocessing + gloo on CPU.  For each case
the script stores the per-rank inputs and what the reference returned on every rank: loss, labels,
d(image_features), d(text_features), d(logit_scale) [, d(logit_bias)].  Inputs are bf16-representable
fp32 values so that the GPU path (bf16 operands) and the reference see identical numbers.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REF_LOSS = "/root/reference/src/open_clip/loss.py"
OUT_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def load_reference():
    spec = importlib.util.spec_from_file_location("mrclip_reference_loss", REF_LOSS)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make_features(n_total: int, d: int, seed: int, corr: float):
    """SURVEY.md §8d recipe: unit-norm rows, text correlated with image, rounded to bf16."""
    g = torch.Generator().manual_seed(seed)
    img = torch.nn.functional.normalize(torch.randn(n_total, d, generator=g), dim=-1)
    txt = torch.nn.functional.normalize(corr * img + (1 - corr) * torch.randn(n_total, d, generator=g) / d ** 0.5,
                                        dim=-1)
    return img.bfloat16().float(), txt.bfloat16().float()


def make_labels(n_total: int, classes: int, seed: int):
    """class labels with repeats (the reference's binned TE/TR/TI classes, preprocessing.py:442-491)"""
    g = torch.Generator().manual_seed(seed + 17)
    return torch.randint(0, classes, (n_total,), generator=g, dtype=torch.int64)


def _worker(rank, world, init_file, case, img, txt, ret):
    ref = load_reference()
    if world > 1:
        dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    n = img.shape[0] // world
    i_loc = img[rank * n:(rank + 1) * n].clone().requires_grad_(True)
    t_loc = txt[rank * n:(rank + 1) * n].clone().requires_grad_(True)
    scale = torch.tensor(case["scale"], dtype=torch.float32, requires_grad=True)
    res = {}
    if case["kind"] == "mpos":
        # MR-CLIP's MultiPositiveClipLoss (loss.py:671-747): samples that share the binned-metadata label are positives
        lab_all = make_labels(case["n_total"], case["classes"], case["seed"])
        lab = lab_all[rank * n:(rank + 1) * n].clone()
        mod = ref.MultiPositiveClipLoss(local_loss=case["local_loss"], gather_with_grad=case["gather_with_grad"],
                                        cache_labels=False, rank=rank, world_size=world)
        out = mod(i_loc, t_loc, scale, delta=case["delta"], tokenized_texts=lab, output_dict=True)
        assert list(out) == ["multi contrastive_loss"]
        loss = out["multi contrastive_loss"]
        res["labels_in"] = lab.numpy()
        (loss * case["grad_output"]).backward()
    elif case["kind"] == "clip":
        mod = ref.ClipLoss(local_loss=case["local_loss"], gather_with_grad=case["gather_with_grad"],
                           cache_labels=True, rank=rank, world_size=world)
        loss = mod(i_loc, t_loc, scale)
        num_logits = n if (world > 1 and case["local_loss"]) else img.shape[0]
        res["labels"] = mod.get_ground_truth(i_loc.device, num_logits).numpy()
        (loss * case["grad_output"]).backward()
    else:
        bias = torch.tensor(case["bias"], dtype=torch.float32, requires_grad=True)
        outs = {}
        for impl in (("bidir", "shift", "reduce", "gather") if world > 1 else ("bidir",)):
            for t in (i_loc, t_loc, scale, bias):
                t.grad = None
            mod = ref.SigLipLoss(rank=rank, world_size=world, dist_impl=impl)
            loss = mod(i_loc, t_loc, scale, bias)
            (loss * case["grad_output"]).backward()
            outs[impl] = (loss.item(), i_loc.grad.clone(), t_loc.grad.clone(), scale.grad.item(), bias.grad.item())
        base = outs["bidir"]
        for impl, o in outs.items():  # the four exchange schemes are the same loss (SURVEY §8c iii)
            assert abs(o[0] - base[0]) <= 1e-5 * max(1.0, abs(base[0])), (impl, o[0], base[0])
            assert torch.allclose(o[1], base[1], rtol=1e-4, atol=1e-6), impl
            assert torch.allclose(o[2], base[2], rtol=1e-4, atol=1e-6), impl
        res["d_bias"] = np.float32(bias.grad.item())
    res.update(loss=np.float32(loss.item()), d_image=i_loc.grad.numpy().copy(), d_text=t_loc.grad.numpy().copy(),
               d_scale=np.float32(scale.grad.item()))
    ret[rank] = res
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_case(case):
    img, txt = make_features(case["n_total"], case["d"], case["seed"], case["corr"])
    world = case["world"]
    if world == 1:
        ret = {}
        _worker(0, 1, None, case, img, txt, ret)
    else:
        mgr = mp.Manager()
        ret = mgr.dict()
        with tempfile.TemporaryDirectory() as td:
            mp.spawn(_worker, args=(world, os.path.join(td, "init"), case, img, txt, ret), nprocs=world, join=True)
        ret = dict(ret)
    out = dict(image=img.numpy(), text=txt.numpy())
    meta = {k: v for k, v in case.items()}
    for k, v in meta.items():
        out["meta_" + k] = np.array(v)
    for r in range(world):
        for k, v in ret[r].items():
            out[f"r{r}_{k}"] = np.asarray(v)
    return out


CASES = []
for (ll, gg) in ((False, False), (False, True), (True, False), (True, True)):
    for world in (2, 4):
        CASES.append(dict(name=f"clip_w{world}_ll{int(ll)}_gg{int(gg)}", kind="clip", world=world, n_total=64, d=32,
                          seed=100 + world, corr=0.3, scale=14.285714, local_loss=ll, gather_with_grad=gg,
                          grad_output=1.0))
CASES += [
    dict(name="clip_w1_small", kind="clip", world=1, n_total=64, d=32, seed=1, corr=0.3, scale=14.285714,
         local_loss=False, gather_with_grad=False, grad_output=1.0),
    dict(name="clip_w1_c1", kind="clip", world=1, n_total=256, d=512, seed=1235, corr=0.5, scale=14.285714,
         local_loss=False, gather_with_grad=False, grad_output=1.0),
    dict(name="clip_w1_scale100_gradscaler", kind="clip", world=1, n_total=200, d=72, seed=7, corr=0.15, scale=100.0,
         local_loss=False, gather_with_grad=False, grad_output=65536.0),
    dict(name="clip_w8_ll1_gg1", kind="clip", world=8, n_total=128, d=64, seed=8, corr=0.3, scale=14.285714,
         local_loss=True, gather_with_grad=True, grad_output=1.0),
    dict(name="clip_w2_ragged", kind="clip", world=2, n_total=300, d=72, seed=9, corr=0.2, scale=30.0,
         local_loss=True, gather_with_grad=True, grad_output=1.0),
    dict(name="mpos_w1", kind="mpos", world=1, n_total=64, d=32, seed=21, corr=0.3, scale=14.285714, classes=12,
         delta=0.5, local_loss=False, gather_with_grad=False, grad_output=1.0),
    dict(name="mpos_w1_delta03_ragged", kind="mpos", world=1, n_total=200, d=72, seed=22, corr=0.15, scale=30.0,
         classes=23, delta=0.3, local_loss=False, gather_with_grad=False, grad_output=3.0),
    dict(name="mpos_w2_ll1_gg1", kind="mpos", world=2, n_total=64, d=32, seed=23, corr=0.3, scale=14.285714,
         classes=10, delta=0.5, local_loss=True, gather_with_grad=True, grad_output=1.0),
    dict(name="mpos_w4_ll1_gg1", kind="mpos", world=4, n_total=128, d=64, seed=24, corr=0.3, scale=14.285714,
         classes=9, delta=0.7, local_loss=True, gather_with_grad=True, grad_output=1.0),
    dict(name="siglip_w1", kind="siglip", world=1, n_total=64, d=32, seed=11, corr=0.3, scale=10.0, bias=-10.0,
         grad_output=1.0),
    dict(name="siglip_w4", kind="siglip", world=4, n_total=64, d=32, seed=12, corr=0.3, scale=10.0, bias=-10.0,
         grad_output=1.0),
    dict(name="siglip_w3_ragged", kind="siglip", world=3, n_total=150, d=40, seed=13, corr=0.3, scale=10.0, bias=-10.0,
         grad_output=2.0),
]


def main(only=None):
    if not os.path.exists(REF_LOSS):
        sys.exit(f"{REF_LOSS} not found: golden vectors can only be generated in the build container")
    os.makedirs(OUT_DIR, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(4)
    for case in CASES:
        if only and not case["name"].startswith(only):
            continue
        out = run_case(case)
        path = os.path.join(OUT_DIR, case["name"] + ".npz")
        np.savez_compressed(path, **out)
        print(f"{case['name']:36s} world={case['world']} loss[r0]={float(out['r0_loss']):.6f} -> {os.path.relpath(path)}")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else None)     # optional name prefix, e.g. "mpos"
