"""Turn the scratch artefacts of tools/gpu_round.sh (gpurun_out/) into the tracked summaries under profiles/.
usage: python tools/make_profiles.py r01"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

# 1. launch list (ncu --metrics gpu__time_duration.sum): per-kernel counts, totals and shares
rows = list(csv.reader(open(os.path.join(G, "launches.csv"))))
hdr = None
agg = collections.OrderedDict()
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    if d.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(d["Metric Value"].replace(",", ""))
    v = v / 1e3 if d["Metric Unit"] == "ns" else (v * 1e3 if d["Metric Unit"] == "ms" else v)
    a = agg.setdefault(d["Kernel Name"], [0, 0.0, d["Grid Size"], d["Block Size"], []])
    a[0] += 1
    a[1] += v
    a[4].append(v)
tot = sum(a[1] for a in agg.values())
with open(os.path.join(P, f"{tag}_launches.md"), "w") as f:
    f.write(f"# ncu launch list, {tag}\n\n`ncu --metrics gpu__time_duration.sum --clock-control none -c 400 python bench.py --steps 2 --warmup 1 "
            "--iters 10 --no-tts --no-cpu-baseline` (config B, 256^3, one B200).  Per-launch times under ncu are cold-cache and "
            "serialised: read the SHARES.  Launches of `k_fpcg` with a duration of a few microseconds are iterations of a CUDA-graph batch "
            "after the solve has finished (they exit on the `done` flag); the median is the figure of a working launch.\n\n")
    f.write("| kernel | launches | total us | share | median us | grid | block |\n|---|---:|---:|---:|---:|---|---|\n")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        med = sorted(a[4])[len(a[4]) // 2]
        f.write(f"| `{k[:100]}` | {a[0]} | {a[1]:.1f} | {100 * a[1] / tot:.1f} % | {med:.1f} | {a[2]} | {a[3]} |\n")

# 2. full-set capture of the production kernel
rep = os.path.join(G, "prof_fused.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
h, units = rr[0], rr[1]
idx = {k: i for i, k in enumerate(h)}
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.avg", "lts__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
best = None
for r in rr[2:]:
    t = float(r[idx["gpu__time_duration.sum"]].replace(",", ""))
    if best is None or t > best[0]:
        best = (t, r)
t, r = best
with open(os.path.join(P, f"{tag}_k_fpcg_ncu.md"), "w") as f:
    f.write(f"# `ncu --set full --clock-control none --import-source on -k regex:k_fpcg`, {tag}\n\nconfig B 256^3, one B200; the working launch "
            "of the capture (a number taken under the profiler is not a bench value).\n\n")
    f.write(f"kernel: `{r[idx['Kernel Name']][:160]}`\n\n| metric | value | unit |\n|---|---:|---|\n")
    for k in keys:
        if k in idx:
            f.write(f"| {k} | {r[idx[k]]} | {units[idx[k]]} |\n")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:k_fpcg"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = None
    data = []
    blocks = 0
    for q in rows:
        if q and q[0] == "Kernel Name":
            blocks += 1
            if blocks > 1:
                break
            continue
        if q and q[0] == "Address":
            hdr = q
            continue
        if hdr and len(q) == len(hdr):
            data.append(dict(zip(hdr, q)))
    if data:
        stalls = [s for s in hdr if s.startswith("stall_") and "Not Issued" not in s]
        totals = {s: sum(int(d[s]) for d in data) for s in stalls}
        T = sum(totals.values())
        f.write(f"\n## warp stall samples ({T})\n\n| reason | share |\n|---|---:|\n")
        for s, v in sorted(totals.items(), key=lambda kv: -kv[1])[:10]:
            f.write(f"| {s} | {100 * v / T:.1f} % |\n")
        mix = collections.Counter()
        for d in data:
            tt = d["Source"].split()
            op = tt[1] if tt[0].startswith("@") else tt[0]
            mix[op.split(".")[0]] += int(d["Instructions Executed"])
        TI = sum(mix.values())
        f.write(f"\n## executed warp instructions ({TI}; {TI * 32 / 16777216:.0f} thread instructions per node)\n\n| opcode | share |\n|---|---:|\n")
        for op, v in mix.most_common(16):
            f.write(f"| {op} | {100 * v / TI:.1f} % |\n")
        f.write("\nSASS evidence of the Blackwell data path: `UTMALDG.3D` (TMA box loads), `SYNCS.ARRIVE.TRANS64` / "
                "`SYNCS.PHASECHK.TRANS64.TRYWAIT` (mbarrier expect_tx / try_wait); no tensor-core instructions (FP64 vector path).\n")
dram = float(r[idx["dram__bytes_read.sum"]].replace(",", "")) + float(r[idx["dram__bytes_write.sum"]].replace(",", ""))
scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.}
dram = float(r[idx["dram__bytes_read.sum"]].replace(",", "")) * scale[units[idx["dram__bytes_read.sum"]]] + \
    float(r[idx["dram__bytes_write.sum"]].replace(",", "")) * scale[units[idx["dram__bytes_write.sum"]]]
json.dump({"k_fpcg_dram_bytes_per_launch": dram, "source": f"profiles/{tag}_k_fpcg_ncu.md", "workload": "config B 256^3"},
          open(os.path.join(P, "traffic.json"), "w"), indent=1)

# 2b. the line-Jacobi iteration: line solve kernel and operator step (k_fpcg MODE 2), headline metrics per kernel
rep2 = os.path.join(G, "prof_line.ncu-rep")
if os.path.exists(rep2):
    raw = subprocess.run(["ncu", "-i", rep2, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(raw)))
    h, units = rr[0], rr[1]
    idx = {k: i for i, k in enumerate(h)}
    per = {}
    for r in rr[2:]:
        name = r[idx["Kernel Name"]]
        t = float(r[idx["gpu__time_duration.sum"]].replace(",", ""))
        if name not in per or t > per[name][0]:
            per[name] = (t, r)
    with open(os.path.join(P, f"{tag}_line_ncu.md"), "w") as f:
        f.write(f"# `ncu --set full --clock-control none -k regex:'k_line_I|k_fpcg' python tools/time_line.py 256 012`, {tag}\n\n"
                "The two kernels of one line-Jacobi PCG iteration (DESIGN.md §4.4), config B 256^3, lines along I; the longest launch of each "
                "kernel in the capture (numbers under the profiler are not bench values; tools/time_line.py prints the event timings).\n")
        for name, (t, r) in per.items():
            f.write(f"\n## `{name[:150]}`\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in keys:
                if k in idx:
                    f.write(f"| {k} | {r[idx[k]]} | {units[idx[k]]} |\n")

# 3. the bench line and the test summary of the same session
for name in ("bench.log", "summary.txt", "smoke.log", "gpu.txt"):
    src = os.path.join(G, name)
    if os.path.exists(src):
        open(os.path.join(P, f"{tag}_{name}"), "w").write(open(src).read())
print("profiles written:", sorted(os.listdir(P)))
