"""SASS evidence for profiles/: mnemonic counts per kernel family from `cuobjdump -sass` of the built library
(one representative instantiation each).  Runs without a GPU.

    python tools/sass_summary.py > profiles/r02_sass_mnemonics.txt
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "easykv_b200", "libeasykv_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
want = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UBLKCP", "SYNCS", "HMMA", "LDSM", "LDGSTS", "FHFMA", "UCGABAR", "MUFU.EX2"]
pick = ["chunk_umma_kernelI6__halfLi4ELb1ELi4E", "decode_umma_kernelI6__halfLi8ELb1E", "decode_umma_kernelI6__halfLi4ELb1E",
        "umma_probe_kernelI6__half", "umma_probe16_kernelI6__half", "decode_kernelI6__halfLi1ELi2E", "decode_cluster_kernelI6__halfLi8ELb1E",
        "chunk_tc_kernelI6__halfLi4ELi2ELb1E", "chunk_tail_kernelI6__half", "chunk_out_kernelI6__halfLi4E"]
print("SASS mnemonic counts per kernel (cuobjdump -sass easykv_b200/libeasykv_b200.so, sm_100a; one instantiation per family).")
print("UTCHMMA = tcgen05.mma | UTCBAR = tcgen05.commit | LDTM / STTM = tcgen05.ld / st | UTMALDG = cp.async.bulk.tensor (tensor-map TMA) |")
print("UBLKCP = cp.async.bulk | SYNCS = mbarrier | UCGABAR = barrier.cluster | HMMA / LDSM = mma.sync / ldmatrix | LDGSTS = cp.async | FHFMA = fma.rn.f32.f16\n")
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n", 1)[0]
    if not any(p in name for p in pick):
        continue
    c = collections.Counter()
    n = 0
    for line in f.split("\n"):
        m = re.search(r"^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            n += 1
            for w in want:
                if m.group(1).startswith(w):
                    c[w] += 1
    d = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    d = re.sub(r"\(ekv::KernelArgs.*", "", d)
    print(f"{d}   [{n} instructions]\n    " + ", ".join(f"{k} {v}" for k, v in sorted(c.items()) if v))
tot = collections.Counter()
for w in want:
    tot[w] = len(re.findall(r"\s" + re.escape(w), txt))
print("\nwhole library: " + ", ".join(f"{k} {v}" for k, v in sorted(tot.items()) if v))
