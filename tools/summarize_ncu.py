"""Dev tool: turn ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.

  python tools/summarize_ncu.py launches gpurun_out/launches.csv profiles/rN_launches_bench_step.md "<command>" "<note>"
  python tools/summarize_ncu.py full gpurun_out/top.ncu-rep profiles/rN_ncu_full_top_kernels.md profiles/traffic.json "<command>"
"""
import csv
import io
import json
import subprocess
import sys
from collections import OrderedDict

FULL_METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
                "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
                "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
                "launch__registers_per_thread", "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers",
                "launch__occupancy_limit_shared_mem", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sector_hit_rate.pct",
                "lts__t_sector_hit_rate.pct", "launch__grid_size", "launch__block_size"]


def short(name):
    name = name.replace("dtb::", "")
    return name.split("(")[0].strip()


def launches(path, out, command, note, marker="energies_fwd_kernel"):
    rows = [r for r in csv.reader(open(path)) if r]
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    data = [r for r in rows[start + 1:] if len(r) > vi and r[mi] == "gpu__time_duration.sum"]
    marks = [i for i, r in enumerate(data) if marker and marker in r[ki]]
    if len(marks) >= 5:                       # one complete step: from the 4th to the 5th launch of the marker kernel
        data = data[marks[3]:marks[4]]
    for r in data:
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] in ("ns", "nsecond") else v
        k = short(r[ki])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    n = sum(a[0] for a in agg.values())
    with open(out, "w") as f:
        f.write("# ncu launch list (per-launch times are cold-cache and serialised: compare SHARES)\n\nCommand (under gpurun, 1 GPU): `%s`\n\n%s\n\n" % (command, note))
        f.write("%d launches, %.1f us in total.\n\n| kernel | launches | total us | share |\n|---|---|---|---|\n" % (n, tot))
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.1f | %.3f |\n" % (k[-90:], a[0], a[1], a[1] / tot))
    print("wrote", out, n, "launches", tot, "us")


def full(rep, out, traffic_out, command):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    seen, traffic = OrderedDict(), {}
    for r in rows[2:]:
        seen.setdefault(short(r[ki]), r)           # first captured launch of each kernel
    with open(out, "w") as f:
        f.write("# `ncu --set full` of the top kernels of the bench step\n\nCommand: `%s`\n\n" % command)
        for k, r in seen.items():
            f.write("## `%s`\n\n| metric | value | unit |\n|---|---|---|\n" % k[-100:])
            for m in FULL_METRICS:
                if m in hdr:
                    f.write("| %s | %s | %s |\n" % (m, r[hdr.index(m)], units[hdr.index(m)]))
            f.write("\n")
            try:
                def val(m):
                    i = hdr.index(m)
                    s = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units[i], 1.0)
                    return float(r[i].replace(",", "")) * s
                key = k.split("<")[0].split(" ")[-1]
                traffic[key] = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
                traffic[key + ".sm_throughput_pct"] = float(r[hdr.index("sm__throughput.avg.pct_of_peak_sustained_elapsed")])
            except Exception as e:  # pragma: no cover
                print("traffic", k, e)
    json.dump(traffic, open(traffic_out, "w"), indent=1)
    print("wrote", out, traffic_out, list(seen))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4], sys.argv[5] if len(sys.argv) > 5 else "")
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4], sys.argv[5])
