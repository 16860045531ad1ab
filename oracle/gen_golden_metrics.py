"""Golden vectors for the retrieval metrics: runs the UNMODIFIED reference ``get_clip_metrics``
(/root/reference/src/open_clip_train/train.py:465-534) and stores inputs and outputs in tests/golden/metrics/*.npz.

Build container only (reads /root/reference).  ``ftfy`` -- imported by the reference's tokenizer, absent in this image
and irrelevant to the function -- is stubbed.  Features are bf16-representable so that the GPU path (bf16 operands) sees
identical numbers; the duplicate-caption case gives all samples of a "unique" label the same text feature, as MR-CLIP's
validation captions do (ties among positives only).
"""
import os
import sys
import types

import numpy as np
import torch

REF_SRC = "/root/reference/src"
OUT_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "metrics")

CASES = [
    dict(name="metrics_small", n=96, d=32, general=12, unique=30, seed=1, corr=0.4, scale=14.285714, dup=False),
    dict(name="metrics_ragged", n=300, d=72, general=7, unique=50, seed=2, corr=0.15, scale=30.0, dup=False),
    dict(name="metrics_singletons", n=130, d=40, general=130, unique=None, seed=3, corr=0.3, scale=100.0, dup=False),
    dict(name="metrics_duplicate_captions", n=257, d=64, general=9, unique=40, seed=4, corr=0.5, scale=14.285714, dup=True),
]


def make_case(c):
    g = torch.Generator().manual_seed(c["seed"])
    n, d = c["n"], c["d"]
    if c["unique"] is None:
        uniq = torch.arange(n)
        gen = uniq.clone()
    else:
        uniq = torch.randint(0, c["unique"], (n,), generator=g)
        gen = uniq % c["general"]
    img = torch.nn.functional.normalize(torch.randn(n, d, generator=g), dim=-1)
    if c["dup"]:
        proto = torch.nn.functional.normalize(torch.randn(c["unique"], d, generator=g), dim=-1)
        txt = proto[uniq]
        img = torch.nn.functional.normalize(c["corr"] * txt + (1 - c["corr"]) * img, dim=-1)
    else:
        txt = torch.nn.functional.normalize(c["corr"] * img + (1 - c["corr"]) * torch.randn(n, d, generator=g) / d ** 0.5, dim=-1)
    return img.bfloat16().float(), txt.bfloat16().float(), [int(x) for x in gen], None if c["unique"] is None else [int(x) for x in uniq]


def main():
    if not os.path.isdir(REF_SRC):
        sys.exit("needs /root/reference (build container only)")
    stub = types.ModuleType("ftfy")
    stub.fix_text = lambda s: s
    sys.modules.setdefault("ftfy", stub)
    sys.path.insert(0, REF_SRC)
    import open_clip_train.train as ref_train
    os.makedirs(OUT_DIR, exist_ok=True)
    for c in CASES:
        img, txt, gen, uniq = make_case(c)
        m = ref_train.get_clip_metrics(img, txt, torch.tensor(c["scale"]), gen, uniq)
        out = dict(image=img.numpy(), text=txt.numpy(), scale=np.float32(c["scale"]), general=np.asarray(gen),
                   has_unique=np.array(uniq is not None))
        if uniq is not None:
            out["unique"] = np.asarray(uniq)
        for k, v in m.items():
            out["m_" + k] = np.float64(v)
        np.savez_compressed(os.path.join(OUT_DIR, c["name"] + ".npz"), **out)
        print(c["name"], {k: round(float(v), 4) for k, v in list(m.items())[:4]})


if __name__ == "__main__":
    main()
