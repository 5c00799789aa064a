"""Per-source-line instruction counts of one kernel from an ncu report (needs -lineinfo + --import-source on).
usage: python tools/ncu_lines.py REPORT.ncu-rep KERNEL_REGEX UNITS [TOP]   (UNITS = divisor, e.g. particles)"""
import collections, csv, os, subprocess, sys
rep, kern, units = sys.argv[1], sys.argv[2], float(sys.argv[3])
top = int(sys.argv[4]) if len(sys.argv) > 4 else 50
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:" + kern],
                     capture_output=True, text=True).stdout
cur = hdr = None
agg = collections.defaultdict(lambda: [0, 0])
for r in csv.reader(out.splitlines()):
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; ix = {h: i for i, h in enumerate(hdr)}; continue
    if hdr is None or len(r) < 12 or r[2] != "-": continue
    try:
        line = int(r[0]); ie = int(r[ix["Instructions Executed"]] or 0); te = int(r[ix["Thread Instructions Executed"]] or 0)
    except Exception:
        continue
    agg[(cur, line)][0] += ie; agg[(cur, line)][1] += te
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ltransv.2b_b200", "csrc")
src = {f: open(os.path.join(root, f)).read().split("\n") for f in os.listdir(root) if f.endswith((".cuh", ".cu", ".h"))}
tot = sum(v[0] for v in agg.values())
print("total warp-instructions per unit: %.1f   (avg active lanes %.1f)" % (tot / units, sum(v[1] for v in agg.values()) / max(1, tot)))
for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    t = src[f][l - 1].strip()[:105] if f in src and l - 1 < len(src[f]) else ""
    print("%8.1f act %4.1f %s:%d %s" % (v[0] / units, v[1] / max(1, v[0]), f, l, t))
