"""Top source lines of an ncu capture (--import-source on): python tools/ncu_source_hot.py rep.ncu-rep [n_lines]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
kern, file_name, hdr, first_file = -1, None, None, None
names = []
data = collections.defaultdict(lambda: collections.defaultdict(lambda: [0, 0, 0, ""]))
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        file_name = r[1].split("/")[-1]
        if first_file is None:
            first_file = r[1]
        if r[1] == first_file:          # the file list restarts for every captured launch
            kern += 1
    elif r[0] == "Function Name":
        if len(names) <= kern:
            names.append(r[1][:80])
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0].isdigit() and len(r) == len(hdr):
        i_s, i_i, i_t = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
        try:
            vals = int(r[i_s] or 0), int(r[i_i] or 0), int(r[i_t] or 0)
        except ValueError:
            continue
        d = data[kern][(file_name, int(r[0]))]
        d[0] += vals[0]; d[1] += vals[1]; d[2] += vals[2]; d[3] = r[1].strip()[:110]
for k in sorted(data):
    tot_s = sum(v[0] for v in data[k].values()); tot_i = sum(v[1] for v in data[k].values()); tot_t = sum(v[2] for v in data[k].values())
    print("=== launch %d %s: warp inst %d, samples %d, avg active lanes %.1f" % (k, names[k], tot_i, tot_s, tot_t / max(tot_i, 1)))
    for (f, l), v in sorted(data[k].items(), key=lambda kv: -kv[1][0])[:top]:
        print("%-12s %4d  smp %5.1f%%  inst %5.1f%%  lanes %4.1f  %s" % (f, l, 100 * v[0] / max(tot_s, 1), 100 * v[1] / max(tot_i, 1), v[2] / max(v[1], 1), v[3]))
