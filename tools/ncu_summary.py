"""profiles/rNN_ncu_summary.json from an `ncu -i rep --page raw --csv` dump of tools/perf_mlp_tc.py <rows> 1 --profile --infer.
usage: ncu -i x.ncu-rep --page raw --csv | python tools/ncu_summary.py <rows> > profiles/r01_ncu_summary.json"""
import csv, json, sys
rows_n = int(sys.argv[1])
rows = list(csv.reader(sys.stdin))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
def val(r, name, unit_scale=None):
    v = float(r[col[name]].replace(",", ""))
    u = units[col[name]]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3}.get(u, 1.0)
    return v * scale
out = {"source": "ncu --set full --clock-control none --import-source on -k regex:mlp_tc_(fwd|bwd|wgrad)_k ; "
                 f"python tools/perf_mlp_tc.py {rows_n} 1 --profile --infer  ({rows_n // 128} tiles of 128 MLP evaluations, one 8x256 "
                 "network); last launch of each kernel", "rows": rows_n, "kernels": {}}
for r in data:   # later launches overwrite earlier ones (warm-up first)
    name = r[col["Kernel Name"]].replace("void ", "").split("(")[0]
    dur = val(r, "gpu__time_duration.sum")
    rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
    out["kernels"][name] = {
        "duration_us": round(dur, 1), "dram_read_GB": round(rd / 1e9, 4), "dram_write_GB": round(wr / 1e9, 4),
        "tensor_pipe_active_pct": round(float(r[col["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"]]), 2),
        "issue_active_pct": round(float(r[col["sm__issue_active.avg.pct_of_peak_sustained_elapsed"]]), 2),
        "registers": int(r[col["launch__registers_per_thread"]]),
        "dram_bytes_per_eval": round((rd + wr) / rows_n, 1), "dram_GBps": round((rd + wr) / dur / 1e3, 1)}
json.dump(out, sys.stdout, indent=1)
