"""Developer script: summarise an `ncu --set full ... --page raw --csv` capture of the tensor-core kernels of one C2 step.

usage: python tools/ncu_summary.py gpurun_out/r2_tc_kernels_ncu_raw.csv profiles/r2_tc_kernels_ncu.csv profiles/r2_conv_ncu_summary.json
Writes a compact per-launch CSV (duration, tensor-pipe activity, DRAM bytes, L2 / shared-memory throughput, registers) and
a JSON with the conv-kernel totals that bench.py reports as ``roofline.traffic`` (cold-cache, serialised replay: an upper
bound of the warm traffic; compare SHARES, not absolutes, with the CUDA-event times of bench.py).
"""
import csv
import json
import sys

src, out_csv, out_json = sys.argv[1:4]
rows = list(csv.reader(open(src)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
want = [("duration_us", "gpu__time_duration.sum"),
        ("tensor_pipe_pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        ("dram_read_MB", "dram__bytes_read.sum"), ("dram_write_MB", "dram__bytes_write.sum"),
        ("l2_throughput_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("l1_smem_throughput_pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("sm_throughput_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("smem_bank_conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
        ("regs_per_thread", "launch__registers_per_thread"), ("grid", "launch__grid_size"),
        ("dyn_smem_KB", "launch__shared_mem_per_block_dynamic")]


def val(r, name):
    i = col.get(name)
    if i is None or i >= len(r):
        return None
    try:
        v = float(r[i].replace(",", ""))
    except ValueError:
        return None
    u = units[i]
    if name.startswith("dram__bytes"):
        scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
        return v * scale
    if name == "gpu__time_duration.sum":
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}.get(u, 1.0)
        return v * scale
    if name == "launch__shared_mem_per_block_dynamic":
        scale = {"byte": 1 / 1024, "Kbyte": 1.0}.get(u, 1 / 1024)
        return v * scale
    return v


out = []
for k, r in enumerate(data):
    name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("(anonymous namespace)::", "")
    row = {"launch": k, "kernel": name}
    for key, metric in want:
        row[key] = val(r, metric)
    out.append(row)
with open(out_csv, "w", newline="") as f:
    w = csv.DictWriter(f, fieldnames=list(out[0].keys()))
    w.writeheader()
    for row in out:
        w.writerow({k: (round(v, 3) if isinstance(v, float) else v) for k, v in row.items()})
conv = [r for r in out if "k_conv" in r["kernel"]]
tot_t = sum(r["duration_us"] for r in conv)
dram = sum((r["dram_read_MB"] or 0) + (r["dram_write_MB"] or 0) for r in conv)
summary = {
    "workload": "C2", "source": f"ncu --set full --clock-control none, one C2 step ({out_csv}); cold-cache serialised replay",
    "conv_launches": len(conv), "conv_duration_us_sum": tot_t, "dram_MB_sum": dram, "dram_bytes_per_launch": dram * 1e6 / max(len(conv), 1),
    "tensor_pipe_pct_time_weighted": sum(r["duration_us"] * (r["tensor_pipe_pct"] or 0) for r in conv) / max(tot_t, 1e-9),
    "tensor_pipe_pct_by_kernel": {},
}
for name in sorted(set(r["kernel"] for r in out)):
    sel = [r for r in out if r["kernel"] == name]
    t = sum(r["duration_us"] for r in sel)
    summary["tensor_pipe_pct_by_kernel"][name] = {"launches": len(sel), "duration_us": round(t, 1),
                                                   "tensor_pipe_pct": round(sum(r["duration_us"] * (r["tensor_pipe_pct"] or 0) for r in sel) / max(t, 1e-9), 1)}
json.dump(summary, open(out_json, "w"), indent=1)
print(json.dumps(summary, indent=1))
