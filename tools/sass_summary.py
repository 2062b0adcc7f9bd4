"""Static SASS opcode counts of the hot kernels: python tools/sass_summary.py [lib.so] > profiles/rNN_sass_summary.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "lidar_slam_arvc_b200/libarvc_icp.so"
OPS = ["LDGSTS", "UBLKCP", "SYNCS", "LDG", "LDS", "STS", "ATOMS", "ATOMG", "SHFL", "DFMA", "DADD", "DMUL", "FFMA", "F2F", "BAR", "UTMALDG", "HMMA"]
WANT = ["k_normals<false, false>", "k_normals_blk<false>", "k_icp_finish<1>", "k_icp_accum<1, false, false>", "k_icp_far<false, false",
        "k_icp_search<false, false", "k_icp_select<false, false>", "k_icp_cond"]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = name.split("(")[0].replace("void ", "").replace("arvc::", "")
        name = re.sub(r"\(bool\)0", "false", re.sub(r"\(bool\)1", "true", re.sub(r"\(int\)", "", name)))
        cur = funcs.setdefault(name, collections.Counter())
        continue
    m = re.search(r"/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur is not None:
        op = m.group(1)
        cur["instr"] += 1
        for o in OPS:
            if op == o or (o in ("LDG", "LDS", "STS", "BAR") and op == o):
                cur[o] += 1
        if op.startswith("HMMA") or op.startswith("UTC"):
            cur["HMMA"] += 1
print("# cuobjdump -sass %s (sm_100a): static instruction counts of the hot kernels (tools/sass_summary.py)" % lib)
print("# (LDGSTS = cp.async global->shared; UBLKCP = cp.async.bulk, SYNCS = mbarrier ops: only in a -DARVC_STAGE_BULK=1 build;")
print("#  no HMMA / UTC*MMA / UTMALDG: the path is not a contraction and moves 16-byte records, not tensor tiles)")
print("%-44s %7s " % ("kernel", "instr") + " ".join("%7s" % o for o in OPS))
for name, c in funcs.items():
    if any(name.startswith(w) or name == w for w in WANT):
        print("%-44s %7d " % (name[:44], c["instr"]) + " ".join("%7d" % c[o] for o in OPS))
