"""Copy one gpu_round.sh visit (gpurun_out/<tag>_*) into profiles/ under the round's names and regenerate the ncu
summaries bench.py and the README refer to.      python tools/refresh_profiles.py <tag> [round, default r01]"""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1]
rnd = sys.argv[2] if len(sys.argv) > 2 else "r01"
COPIES = {"timeline_bench.txt": "timeline_persistent_bench_state.txt", "decode_umma_timeline.txt": "decode_umma_timeline_bench_state.txt", "smoke.txt": "smoke.txt",
          "bench.json": "bench_b64.json", "bench_b32.json": "bench_b32.json", "bench_ref.json": "bench_reference_arm.json",
          "pytest.txt": "pytest_gpu.txt", "sweep_decode.jsonl": "sweep_decode.jsonl", "sweep_cluster.jsonl": "sweep_decode_cluster.jsonl",
          "sweep_chunk.jsonl": "sweep_chunk.jsonl", "chunk_launches.csv": "chunk_launches.csv", "launches.csv": "launches.csv",
          "e2e_llama7b.json": "e2e_generate_llama7b.json", "e2e_mistral7b.json": "e2e_generate_mistral7b_c3_literal.json"}
for src, dst in COPIES.items():
    s = os.path.join(G, f"{tag}_{src}")
    if os.path.exists(s) and os.path.getsize(s):
        shutil.copy(s, os.path.join(P, f"{rnd}_{dst}"))
KEEP = ("Kernel Name", "Grid Size", "Block Size", "gpu__time_duration", "dram__bytes", "dram__throughput", "gpu__dram_throughput",
        "dram__cycles_active", "sm__warps_active", "launch__", "sm__throughput", "lts__t_sector_hit_rate", "lts__t_bytes",
        "smsp__issue_active", "sm__inst_executed_pipe", "l1tex__data_bank_conflicts", "smsp__cycles_active", "sm__cycles_elapsed",
        "sm__pipe_tensor", "smsp__average_warp", "smsp__inst_executed.sum")


def raw(rep, dst):
    if not os.path.exists(rep):
        return None
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(rows) - 2)])
        for h in hdr:
            if any(k in h for k in KEEP):
                w.writerow([h, units[idx[h]]] + [r[idx[h]] for r in rows[2:]])
    return rows, idx, units


r = raw(os.path.join(G, f"{tag}_decode.ncu-rep"), os.path.join(P, f"{rnd}_decode_ncu_raw.csv"))
if r:
    rows, idx, units = r
    val = lambda row, h: float(row[idx[h]].replace(",", ""))
    tb = lambda row, h: val(row, h) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[units[idx[h]]]
    tr = [tb(x, "dram__bytes_read.sum") + tb(x, "dram__bytes_write.sum") for x in rows[2:]]
    traffic = {"c2": {"round": int(rnd[1:]), "kernel": "ekv::decode_kernel<__half,1,2,false>", "seqs_per_gpu": 64,
                      "dram_bytes_per_launch": sum(tr) / len(tr),
                      "dram_bytes_read": [tb(x, "dram__bytes_read.sum") for x in rows[2:]],
                      "dram_bytes_write": [tb(x, "dram__bytes_write.sum") for x in rows[2:]],
                      "gpu_time_us": [val(x, "gpu__time_duration.sum") for x in rows[2:]], "bytes_alg_per_launch": 1197531136,
                      "source": "ncu --set full --clock-control none -k regex:decode_kernel -s 60 -c 1 python bench.py --steps 2 --warmup 3 "
                                "--no-sweep --no-cpu-baseline --no-gpu-reference --min-seconds 0.05 --layers 4 (gpurun, 1x B200, tools/gpu_ncu.sh)"}}
    for wl, rep, seqs, src in (("c5", "decode_umma", 8, "ncu --set full -k regex:decode_umma_kernel -c 1 python bench.py --workload c5 ... (tools/gpu_ncu.sh)"),
                               ("c3_chunk", "chunk_umma", 8, "ncu --set full -k regex:chunk_umma_kernel -c 1 python tools/chunk_profile.py 8 32 8 8208 16 h2o_head (tools/gpu_ncu.sh); "
                                                              "the tcgen05 kernel only (chunk_out / chunk_tail add their own passes over the state)")):
        rr = raw(os.path.join(G, f"{tag}_{rep}.ncu-rep"), os.path.join(P, f"{rnd}_{rep}_ncu_raw.csv"))
        if rr:
            rows2, idx2, units2 = rr
            tb2 = lambda row, h: float(row[idx2[h]].replace(",", "")) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[units2[idx2[h]]]
            x = rows2[2]
            traffic[wl] = {"round": int(rnd[1:]), "kernel": x[idx2["Kernel Name"]][:60], "seqs_per_gpu": seqs,
                           "dram_bytes_per_launch": tb2(x, "dram__bytes_read.sum") + tb2(x, "dram__bytes_write.sum"),
                           "gpu_time_us": [float(x[idx2["gpu__time_duration.sum"]].replace(",", ""))], "source": src}
    json.dump(traffic, open(os.path.join(P, f"traffic_{rnd}.json"), "w"), indent=1)
    print("decode: traffic", sum(tr) / len(tr), "us", [val(x, "gpu__time_duration.sum") for x in rows[2:]])
for rep, dst in (("chunk_umma", "chunk_umma_ncu_raw.csv"), ("decode_umma", "decode_umma_ncu_raw.csv"), ("chunk_tc", "chunk_tc_ncu_raw.csv")):
    r = raw(os.path.join(G, f"{tag}_{rep}.ncu-rep"), os.path.join(P, f"{rnd}_{dst}"))
    if r:
        rows, idx, units = r
        g = lambda x, h: x[idx[h]] if h in idx else "n/a"
        for x in rows[2:]:
            print(x[idx["Kernel Name"]][:70], x[idx["gpu__time_duration.sum"]], "us; dram read", g(x, "dram__bytes_read.sum"), g(x, "dram__bytes_write.sum"),
                  "; tensor pipe", g(x, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"), "% issue",
                  g(x, "smsp__issue_active.avg.pct_of_peak_sustained_active"), "% regs", g(x, "launch__registers_per_thread"))
for f in ("bench_b64.json", "bench_b32.json", "bench_reference_arm.json"):
    p = os.path.join(P, f"{rnd}_{f}")
    if os.path.exists(p):
        d = json.loads(open(p).read().strip().splitlines()[-1])
        print(f, "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "frac", (d.get("roofline") or {}).get("frac"),
              d.get("clocks"), (d.get("cpu_baseline") or {}).get("value"))
for f in ("sweep_decode.jsonl", "sweep_chunk.jsonl"):
    p = os.path.join(P, f"{rnd}_{f}")
    if os.path.exists(p):
        for l in open(p):
            if l.startswith("{"):
                d = json.loads(l)
                print(f"  {d['case'][:62]:62s} {d['us_per_launch']:9.1f} us {d['frac_of_measured']:.3f}")
