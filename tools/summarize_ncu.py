"""Turn ncu exports into the text summaries committed under profiles/.

  python tools/summarize_ncu.py launches gpurun_out/launches.csv            > profiles/r01_launches.txt
  python tools/summarize_ncu.py full gpurun_out/prof.ncu-rep                > profiles/r01_full.txt
"""
import collections
import csv
import json
import subprocess
import sys


def short(name):
    return name.split("(")[0].replace("void ", "").replace("arvc::", "").replace("<unnamed>::", "")


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = short(row["Kernel Name"])
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit.startswith("n") else (v * 1e3 if unit.startswith("m") else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised launches: compare SHARES)")
    print("%-44s %8s %14s %8s %12s" % ("kernel", "launches", "total_us", "share", "avg_us"))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-44s %8d %14.1f %7.1f%% %12.1f" % (k[:44], v[0], v[1], 100 * v[1] / tot, v[1] / v[0]))
    print("%-44s %8d %14.1f" % ("TOTAL", sum(v[0] for v in agg.values()), tot))


WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_fp64.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "sm__cycles_elapsed.max", "smsp__cycles_active.avg"]


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    traffic = {}

    def num(r, k):
        v = float(r[idx[k]].replace(",", ""))
        u = units[idx[k]].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)

    for r in rows[2:]:
        name = short(r[idx["Kernel Name"]])
        print("== %s" % name)
        for w in WANT:
            if w in idx:
                print("  %-70s %s %s" % (w, r[idx[w]], units[idx[w]]))
        if "dram__bytes_read.sum" in idx:
            traffic.setdefault(name, []).append(num(r, "dram__bytes_read.sum") + num(r, "dram__bytes_write.sum"))
    print("# dram traffic per launch (bytes):", json.dumps(traffic))


def traffic(paths, n_scans=9):
    """profiles/traffic.json from --set full captures of tools/prof_target.py <n_scans> (k_normals + k_icp_search passes);
    several reports (comma separated) are merged."""
    rows, hdr, units = [None, None], None, None
    for path in paths.split(","):
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rr = list(csv.reader(out.splitlines()))
        if hdr is None:
            hdr, units = rr[0], rr[1]
            rows += rr[2:]
        else:                                   # same metric set: align the columns by name
            pos = {h: i for i, h in enumerate(rr[0])}
            rows += [[r[pos[h]] if h in pos else "0" for h in hdr] for r in rr[2:]]
    path = paths
    idx = {h: i for i, h in enumerate(hdr)}

    def num(r, k):
        v = float(r[idx[k]].replace(",", ""))
        u = units[idx[k]].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)

    def pcts(r):
        return {"issue_active_pct": num(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                "fma_pipe_pct": num(r, "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
                "fp64_pipe_pct": num(r, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
                "warps_active_pct": num(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                "l2_hit_pct": num(r, "lts__t_sector_hit_rate.pct"),
                "dram_pct_of_peak": num(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")}

    res = {"source": "ncu --set full of tools/prof_target.py %d (%s)" % (n_scans, path.split("/")[-1])}
    nrm = [r for r in rows[2:] if "k_normals_blk" in r[idx["Kernel Name"]]] or [r for r in rows[2:] if "k_normals<" in r[idx["Kernel Name"]]]
    srch = [r for r in rows[2:] if "k_icp_search" in r[idx["Kernel Name"]]]
    if nrm:
        r = nrm[0]
        d = pcts(r)
        d["dram_bytes_per_scan"] = (num(r, "dram__bytes_read.sum") + num(r, "dram__bytes_write.sum")) / n_scans
        d["warp_instructions"] = num(r, "smsp__inst_executed.sum")
        d["registers"] = num(r, "launch__registers_per_thread")
        res["normals"] = d
    if srch:
        per = [(num(r, "dram__bytes_read.sum") + num(r, "dram__bytes_write.sum")) / (n_scans - 1) for r in srch]
        d = {"dram_bytes_per_pair_pass": sum(per) / len(per), "by_pass": per, "registers": num(srch[0], "launch__registers_per_thread"),
             "warp_instructions_by_pass": [num(r, "smsp__inst_executed.sum") for r in srch],
             "active_lanes_by_pass": [num(r, "smsp__thread_inst_executed_per_inst_executed.ratio") for r in srch]}
        for k in pcts(srch[0]):
            d[k] = [pcts(r)[k] for r in srch]
        res["icp_pass"] = d
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](sys.argv[2])
