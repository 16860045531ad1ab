"""torchrun worker for the multi-GPU parity test: every rank runs the public modules on its share of a
golden case (NCCL), compares with what the reference produced on that rank and exits non-zero on mismatch.

    torchrun --nproc-per-node W tests/dist_worker.py <golden-name> [<golden-name> ...]
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import load_golden, rel_err  # noqa: E402
from mrclip_b200 import ClipLoss, MultiPositiveClipLoss, SigLipLoss  # noqa: E402


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    failures = []
    for name in sys.argv[1:]:
        case = load_golden(name)
        m = case["meta"]
        assert case["world"] == world, (name, case["world"], world)
        # emat: collectives over NVLink peer memory (push all-gathers, GEMM push epilogue); emat-nccl: the same through NCCL
        # emat: whole-step C entries, every exchange over NVLink peer memory (bf16 gradient tiles); emat-f32: the same with
        # the fp32 payload; emat-py: the per-kernel Python orchestration over peer memory; emat-nccl: over NCCL
        backends = ("emat", "emat-f32", "emat-py", "emat-nccl", "gmat", "fused")
        for backend in backends:
            os.environ["MRCLIP_BWD"] = backend.split("-")[0]
            os.environ["MRCLIP_RS"] = "nccl" if backend.endswith("-nccl") else "push"
            os.environ["MRCLIP_PUSH_DTYPE"] = "fp32" if backend.endswith("-f32") else "bf16"
            os.environ["MRCLIP_STEP"] = "py" if backend.endswith("-py") else "c"
            os.environ["MRCLIP_AG"] = os.environ["MRCLIP_RS"]      # all-gathers: NCCL or peer stores, likewise
            n = case["image"].shape[0] // world
            rows = slice(rank * n, (rank + 1) * n)
            i = torch.from_numpy(case["image"][rows]).to(dev).requires_grad_(True)
            t = torch.from_numpy(case["text"][rows]).to(dev).requires_grad_(True)
            s = torch.tensor(float(m["scale"]), device=dev, requires_grad=True)
            ref = case["ranks"][rank]
            if m["kind"] == "mpos":
                if not backend.startswith("emat"):
                    continue        # the multi-positive loss always runs the E-block pipeline
                mod = MultiPositiveClipLoss(local_loss=bool(m["local_loss"]), gather_with_grad=bool(m["gather_with_grad"]),
                                            rank=rank, world_size=world)
                lab = torch.from_numpy(ref["labels_in"]).to(dev)
                loss = mod(i, t, s, delta=float(m["delta"]), tokenized_texts=lab)
                ok_labels = True
            elif m["kind"] == "clip":
                mod = ClipLoss(local_loss=bool(m["local_loss"]), gather_with_grad=bool(m["gather_with_grad"]),
                               cache_labels=True, rank=rank, world_size=world)
                loss = mod(i, t, s)
                nl = n if (world > 1 and m["local_loss"]) else case["image"].shape[0]
                ok_labels = np.array_equal(mod.get_ground_truth(dev, nl).cpu().numpy(), ref["labels"])
            else:
                b = torch.tensor(float(m["bias"]), device=dev, requires_grad=True)
                loss = SigLipLoss(rank=rank, world_size=world)(i, t, s, b)
                ok_labels = True
            (loss * float(m["grad_output"])).backward()
            errs = dict(loss=abs(loss.item() - float(ref["loss"])) / abs(float(ref["loss"])),
                        d_image=rel_err(i.grad.cpu().numpy(), ref["d_image"]),
                        d_text=rel_err(t.grad.cpu().numpy(), ref["d_text"]),
                        d_scale=abs(s.grad.item() - float(ref["d_scale"])) / max(abs(float(ref["d_scale"])), 1e-6))
            if m["kind"] == "siglip":
                errs["d_bias"] = abs(b.grad.item() - float(ref["d_bias"])) / max(abs(float(ref["d_bias"])), 1e-6)
            bad = [k for k, v in errs.items() if v > (1e-3 if k == "loss" else 1e-2)] + ([] if ok_labels else ["labels"])
            if bad:
                failures.append((name, backend, rank, bad, errs))
            if rank == 0:
                print(f"{name:32s} {backend:9s} " + " ".join(f"{k}={v:.2e}" for k, v in errs.items()), flush=True)
    for step in ("c", "py"):
        # forward-side d logit_scale against the entropy path on the same inputs, at a size where <dT_r, T_r> averages
        # the bf16 rounding of G out (n*N >= 2^22); whole-step C entries and Python orchestration
        os.environ["MRCLIP_BWD"], os.environ["MRCLIP_RS"], os.environ["MRCLIP_AG"] = "emat", "push", "push"
        os.environ["MRCLIP_STEP"] = step
        N, D = 16384, 384
        n = N // world
        g = torch.Generator().manual_seed(77)
        img = torch.nn.functional.normalize(torch.randn(N, D, generator=g), dim=-1)
        txt = torch.nn.functional.normalize(0.4 * img + 0.6 * torch.randn(N, D, generator=g) / D ** 0.5, dim=-1)
        res = {}
        for mode in ("entropy", "fwd"):
            os.environ["MRCLIP_DS"] = mode
            i = img[rank * n:(rank + 1) * n].bfloat16().to(dev).requires_grad_(True)
            t = txt[rank * n:(rank + 1) * n].bfloat16().to(dev).requires_grad_(True)
            s = torch.tensor(30.0, device=dev, requires_grad=True)
            mod = ClipLoss(local_loss=True, gather_with_grad=True, rank=rank, world_size=world)
            (mod(i, t, s) * 3.0).backward()
            res[mode] = (s.grad.item(), i.grad.float().cpu().numpy(), t.grad.float().cpu().numpy())
        os.environ["MRCLIP_DS"] = "fwd"
        errs = dict(d_scale=abs(res["fwd"][0] - res["entropy"][0]) / abs(res["entropy"][0]),
                    d_image=rel_err(res["fwd"][1], res["entropy"][1]), d_text=rel_err(res["fwd"][2], res["entropy"][2]))
        print(f"rank {rank} MRCLIP_STEP={step} MRCLIP_DS=fwd vs entropy: " + " ".join(f"{k}={v:.2e}" for k, v in errs.items()), flush=True)
        if errs["d_scale"] > 2e-3 or errs["d_image"] > 1e-6 or errs["d_text"] > 1e-6:
            failures.append(("fwd_ds", "emat", rank, list(errs), errs))
    os.environ["MRCLIP_STEP"] = "c"
    flag = torch.tensor([len(failures)], device=dev)
    dist.all_reduce(flag)
    for f in failures:
        print("MISMATCH", f, flush=True)
    dist.destroy_process_group()
    sys.exit(1 if flag.item() else 0)


if __name__ == "__main__":
    main()
