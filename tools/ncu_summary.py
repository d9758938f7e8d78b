"""Summarise an .ncu-rep: headline metrics per kernel, stall reasons and instruction mix of the first kernel.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-regex]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
rx = sys.argv[2] if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__shared_mem_per_block_dynamic', 'launch__grid_size', 'launch__block_size',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__cycles_elapsed.avg', 'lts__t_bytes.sum',
        'lts__t_sector_hit_rate.pct', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__cycles_active.avg', 'l1tex__m_xbar2l1tex_read_bytes.sum', 'sm__inst_executed_pipe_lsu.sum',
        'smsp__inst_executed_op_shared_ld.sum', 'smsp__inst_executed_op_shared_st.sum', 'smsp__inst_executed_op_global_ld.sum',
        'smsp__inst_executed_op_global_st.sum', 'local_load_requests', 'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum']
for r in rows[2:]:
    name = r[idx['Kernel Name']]
    print('=====', name[:110])
    for k in keys:
        if k in idx:
            print(f"  {k:82s} {r[idx[k]]:>18s} {units[idx[k]]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + (["--kernel-name", "regex:" + rx] if rx else []),
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = None
data = []
blocks = 0
for r in rows:
    if r and r[0] == "Kernel Name":
        blocks += 1
        if blocks > 1:
            break
        continue
    if r and r[0] == "Address":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        data.append(dict(zip(hdr, r)))
if data:
    stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    tot = {s: sum(int(d[s]) for d in data) for s in stalls}
    T = sum(tot.values())
    print("---- stall samples (all):", T)
    for s, v in sorted(tot.items(), key=lambda kv: -kv[1])[:10]:
        print(f"  {s:28s} {v:8d} {100 * v / T:5.1f}%")
    mix = collections.Counter()
    for d in data:
        t = d['Source'].split()
        op = t[1] if t[0].startswith('@') else t[0]
        mix[op.split('.')[0]] += int(d['Instructions Executed'])
    TI = sum(mix.values())
    print("---- warp instructions:", TI)
    for op, v in mix.most_common(28):
        print(f"  {op:12s} {v:10d} {100 * v / TI:5.1f}%")
    # hottest instructions by stall samples
    print("---- top instructions by samples")
    for d in sorted(data, key=lambda d: -int(d['# Samples']))[:25]:
        print(f"  {int(d['# Samples']):6d}  {d['Source'][:100]}")
