"""Print the bench lines of the round-2 scaling study (gpurun_out/r02_scale_*.log) as a table."""
import glob, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = []
for f in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "r02_scale_*_n*.log"))):
    for l in open(f):
        if l.startswith('{"metric"'):
            d = json.loads(l)
            t = d.get("time_to_solution") or {}
            rows.append((d["scaling"] if d["n_gpus"] > 1 else os.path.basename(f).split("_")[2], d["n_gpus"], d["value"] / 1e9, d["ms_per_step"] / d["config"]["iters_per_step"],
                         d["e2e"]["value"] / 1e9, d["clocks"]["sm_mhz"], t.get("seconds"), t.get("pcg_iterations"), (t.get("line_jacobi") or {}).get("seconds"),
                         (t.get("multilevel") or {}).get("seconds"), (d.get("parity") or {}).get("max_abs_dT_K")))
print("| scaling | GPUs | G DOF·iter/s | ms / iteration | e2e G DOF·iter/s | SM MHz | TTS jac s | iterations | TTS ljac s | TTS mlj s | parity max|dT| K |")
print("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---|")
for r in rows:
    print("| " + " | ".join("" if v is None else (f"{v:.4g}" if isinstance(v, float) else str(v)) for v in r) + " |")
