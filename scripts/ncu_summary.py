"""Summarise an .ncu-rep (read on the CPU box): per-launch duration, DRAM bytes, tensor/LSU utilisation,
and (optionally) the top stall instructions of one kernel instance.
usage: ncu_summary.py report.ncu-rep [--stalls KERNEL_INDEX]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
cols = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "rdMB"), ("dram__bytes_write.sum", "wrMB"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1%"),
        ("smsp__inst_executed.sum", "winst"), ("sm__cycles_elapsed.max", "cycles"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid")]
def conv(v, u):
    try: v = float(v.replace(",", ""))
    except ValueError: return v
    if u in ("byte",): return v / 1e6
    if u in ("Kbyte",): return v / 1e3
    if u in ("Gbyte",): return v * 1e3
    if u in ("ns",): return v / 1e3
    if u in ("ms",): return v * 1e3
    return v
print("idx kernel".ljust(44) + " ".join(n.rjust(9) for _, n in cols))
tot = 0.0
for k, d in enumerate(data):
    name = d[ix["Kernel Name"]].replace("void ", "").replace("ievm::", "")[:38]
    vals = []
    for c, n in cols:
        v = conv(d[ix[c]], units[ix[c]]) if c in ix else ""
        vals.append(("%.2f" % v if isinstance(v, float) and v < 1e5 else ("%d" % v if isinstance(v, float) else str(v))).rjust(9))
    tot += conv(d[ix["gpu__time_duration.sum"]], units[ix["gpu__time_duration.sum"]])
    print(("%2d %s" % (k, name)).ljust(44) + " ".join(vals))
print("total us %.1f" % tot)
if "--stalls" in sys.argv:
    which = int(sys.argv[sys.argv.index("--stalls") + 1])
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    secs, cur = [], None
    for r in csv.reader(io.StringIO(src)):
        if r and r[0] == "Kernel Name": cur = {"name": r[1], "rows": []}; secs.append(cur)
        elif r and r[0] == "Address": cur["hdr"] = r
        elif cur is not None and r: cur["rows"].append(r)
    s = secs[which]; h = s["hdr"]; jx = {k: i for i, k in enumerate(h)}
    stall = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
    tot = sum(int(r[jx["# Samples"]]) for r in s["rows"])
    print(s["name"], "samples", tot)
    agg = {k: sum(int(r[jx[k]]) for r in s["rows"]) for k in stall}
    print(sorted(agg.items(), key=lambda kv: -kv[1])[:8])
    order = sorted(range(len(s["rows"])), key=lambda i: -int(s["rows"][i][jx["# Samples"]]))[:int(sys.argv[sys.argv.index("--stalls") + 2]) if len(sys.argv) > sys.argv.index("--stalls") + 2 else 16]
    for i in order:
        r = s["rows"][i]
        st = {k: int(r[jx[k]]) for k in stall if int(r[jx[k]]) > 0}
        print(str(i).rjust(5), r[jx["# Samples"]].rjust(6), r[jx["Instructions Executed"]].rjust(9), r[jx["Source"]].strip()[:76].ljust(76), sorted(st.items(), key=lambda kv: -kv[1])[:3])
