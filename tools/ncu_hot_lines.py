"""Per-source-line view of an ncu capture (development aid): warp-stall samples and executed instructions of one
kernel, mapped from SASS addresses to file:line through `nvdisasm -g` of the matching cubin in the built library.

    python tools/ncu_hot_lines.py gpurun_out/x.ncu-rep ekv_decode.sm_100a.cubin _ZN3ekv13decode_kernelI6__halfLi1ELi2E [top]

The library must be the one the capture ran (same SASS); compile with -lineinfo (build.py does)."""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, cubin, mangled = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "easykv_b200", "libeasykv_b200.so")], cwd=tmp, capture_output=True)
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
line_of, cur, inside = {}, None, False
for ln in dis:
    if ln.startswith(".text."):
        inside = mangled in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*);", ln)
    if m:
        line_of[int(m.group(1), 16)] = (cur, m.group(2).strip())
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
blocks = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
per_line = collections.defaultdict(lambda: [0, 0])
total = 0
for bi, b in enumerate(blocks[:1]):
    hdr = rows[b + 1]
    ci, cx = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
    body = rows[b + 2:(blocks[bi + 1] if bi + 1 < len(blocks) else len(rows))]
    base = int(body[0][0], 16)
    for r in body:
        if len(r) <= cx or not r[ci].isdigit():
            continue
        off = int(r[0], 16) - base
        src = line_of.get(off, (None, r[1]))[0]
        per_line[src][0] += int(r[ci]); per_line[src][1] += int(r[cx]); total += int(r[ci])
print(f"{rows[blocks[0]][1][:100]}: {total} samples")
files = {}
for src, (smp, ins) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
    text = ""
    if src:
        if src[0] not in files:
            for d in ("easykv_b200/csrc", "include"):
                p = os.path.join(ROOT, d, src[0])
                if os.path.exists(p):
                    files[src[0]] = open(p).read().splitlines()
        f = files.get(src[0])
        text = f[src[1] - 1].strip()[:110] if f and src[1] <= len(f) else ""
    print(f"{100.0 * smp / max(total, 1):5.1f}%  {smp:6d} smp {ins:9d} inst  {src[0] + ':' + str(src[1]) if src else '?':28s} {text}")
