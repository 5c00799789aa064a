"""Warp-stall samples per source line of one kernel from an ncu report (needs -lineinfo + --import-source on).
usage: python tools/ncu_stalls.py REPORT.ncu-rep KERNEL_REGEX [TOP]"""
import collections, csv, os, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:" + kern],
                     capture_output=True, text=True).stdout
cur = hdr = None
KEYS = ["stall_long_sb", "stall_wait", "stall_short_sb", "stall_lg", "stall_barrier", "stall_no_inst", "stall_math", "stall_mio",
        "stall_branch_resolving", "stall_not_selected", "stall_selected", "stall_dispatch"]
agg = collections.defaultdict(lambda: collections.Counter())
tot = collections.Counter()
for r in csv.reader(out.splitlines()):
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; ix = {h: i for i, h in enumerate(hdr)}; continue
    if hdr is None or len(r) < 40 or r[2] != "-": continue
    try: line = int(r[0])
    except Exception: continue
    for k in KEYS:
        v = int(r[ix[k]] or 0)
        agg[(cur, line)][k] += v; tot[k] += v
    agg[(cur, line)]["all"] += int(r[ix["# Samples"]] or 0); tot["all"] += int(r[ix["# Samples"]] or 0)
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ltransv.2b_b200", "csrc")
src = {f: open(os.path.join(root, f)).read().split("\n") for f in os.listdir(root) if f.endswith((".cuh", ".cu", ".h"))}
print("samples:", tot["all"], {k: "%.1f%%" % (100.0 * tot[k] / max(1, tot["all"])) for k in KEYS if tot[k]})
for (f, l), c in sorted(agg.items(), key=lambda kv: -kv[1]["all"])[:top]:
    t = src[f][l - 1].strip()[:90] if f in src and l - 1 < len(src[f]) else ""
    print("%5.1f%% (lsb %4.1f wait %4.1f ssb %4.1f bar %4.1f) %s:%d %s" % (100.0 * c["all"] / tot["all"], 100.0 * c["stall_long_sb"] / tot["all"],
          100.0 * c["stall_wait"] / tot["all"], 100.0 * c["stall_short_sb"] / tot["all"], 100.0 * c["stall_barrier"] / tot["all"], f, l, t))
