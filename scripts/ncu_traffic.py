"""Condense an `ncu --set full` report of one forward into profiles/<name>.json: per launch duration, DRAM
bytes read+written, tensor-pipe and L1 data-pipe utilisation.  bench.py reads the DRAM bytes of the dominant
kernel from that file for `roofline.traffic`.
usage: ncu_traffic.py report.ncu-rep out.json"""
import csv, io, json, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}


def val(d, name, scale_units=True):
    if name not in ix:
        return None
    try:
        v = float(d[ix[name]].replace(",", ""))
    except ValueError:
        return None
    u = units[ix[name]]
    if scale_units:
        v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1e3, "ms": 1e6, "s": 1e9}.get(u, 1.0)   # -> bytes / ns
    return v


launches = []
for d in data:
    name = d[ix["Kernel Name"]].replace("void ", "").replace("ievm::", "").split("(")[0]
    rd, wr = val(d, "dram__bytes_read.sum"), val(d, "dram__bytes_write.sum")
    launches.append({
        "kernel": name,
        "duration_us": round(val(d, "gpu__time_duration.sum") / 1e3, 2),
        "dram_read_bytes": rd, "dram_write_bytes": wr,
        "tensor_pipe_pct": val(d, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", False),
        "l1_tc_operand_wavefronts_pct": val(d, "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", False),
        "l1_lsu_wavefronts_pct": val(d, "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", False),
        "issue_active_pct": val(d, "smsp__issue_active.avg.pct_of_peak_sustained_active", False),
        "registers": val(d, "launch__registers_per_thread", False),
        "grid": val(d, "launch__grid_size", False),
    })
conv = [l for l in launches if l["kernel"].startswith(("conv_tc_kernel", "conv_dual_kernel", "conv_s2_kernel"))]
summary = {
    "source": rep.split("/")[-1],
    "how": "ncu --set full --clock-control none, one forward at batch 256 (cold cache, serialised launches)",
    "conv_tc_launches": len(conv),
    "conv_tc_dram_bytes_per_launch": (sum(l["dram_read_bytes"] + l["dram_write_bytes"] for l in conv) / len(conv)) if conv else None,
    "total_dram_bytes": sum((l["dram_read_bytes"] or 0) + (l["dram_write_bytes"] or 0) for l in launches),
    "total_duration_us": round(sum(l["duration_us"] for l in launches), 1),
    "launches": launches,
}
json.dump(summary, open(out, "w"), indent=1)
print(json.dumps({k: v for k, v in summary.items() if k != "launches"}))
