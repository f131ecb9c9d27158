"""Dev tool: per-source-line instruction counts / stall samples of one kernel from an ncu report (needs -lineinfo + --import-source on).
usage: python tools/ncu_hot_lines.py report.ncu-rep kernel_regex [top_n]"""
import csv, subprocess, sys, collections
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = None
agg = collections.OrderedDict()
tot_inst = tot_samp = 0
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) < 10 or r[0] in ("Line No", ""):
        continue
    try:
        line = int(r[0]); inst = int(r[7]); samp = int(r[4]); tinst = int(r[8])
    except ValueError:
        continue
    key = (cur_file, line)
    a = agg.setdefault(key, [r[1].strip()[:110], 0, 0, 0])
    a[1] += inst; a[2] += samp; a[3] += tinst
    tot_inst += inst; tot_samp += samp
print("total warp instructions %d, samples %d" % (tot_inst, tot_samp))
for (f, l), (src, inst, samp, tinst) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%5.1f%% inst %5.1f%% stall  lanes %4.1f  %s:%d  %s" % (100.0 * inst / max(tot_inst, 1), 100.0 * samp / max(tot_samp, 1), tinst / max(inst, 1), f, l, src))
