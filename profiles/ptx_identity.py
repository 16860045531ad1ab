"""Compare the PTX of every kernel entry between two builds, ignoring label / register / call-sequence numbering:
    nvcc -gencode arch=compute_100a,code=compute_100a -O3 -std=c++17 -ptx -o old.ptx <old tree>/mrclip_b200/csrc/mrclip_cabi.cu
    nvcc ... -o new.ptx mrclip_b200/csrc/mrclip_cabi.cu
    python profiles/ptx_identity.py old.ptx new.ptx
Used to show that default-off additions leave the kernels validated on the GPU untouched."""
import hashlib
import re
import sys


def entries(path):
    out = {}
    for m in re.finditer(r"\.visible \.entry (\w+)\((.*?)\n\}\n", open(path).read(), re.S):
        name, body = m.group(1), m.group(0)
        key = re.sub(r"gemm2_kernelILb(\d)EL[bi](\d)E", r"gemm2_kernel<\1,\2>", name)   # bool -> int template parameter
        body = body.replace(name, "NAME")
        for pat, rep in ((r"\$L__BB\d+_", "$L__BB_"), (r"__local_depot\d+", "__local_depot"), (r"_param_3\[\d+\]", "_param_3[]"),
                         (r"// callseq \d+(, \d+)?", ""), (r"%(r|rd|p|f|rs|fd)\d+", r"%\1")):
            body = re.sub(pat, rep, body)
        out[key] = hashlib.sha1(body.encode()).hexdigest()
    return out


old, new = entries(sys.argv[1]), entries(sys.argv[2])
changed = [k for k in old if old[k] != new.get(k)]
print(f"{len(old) - len(changed)} of {len(old)} entries identical; {len(new) - len(old)} new entries")
for k in changed:
    print("CHANGED" if k in new else "REMOVED", k)
sys.exit(1 if changed else 0)
