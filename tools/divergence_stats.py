"""Distribution of CUDA-vs-oracle differences for BASELINE configs[0] (5,000 passive particles)
after 1 internal step, 1 external step and `nx` external steps."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import World, LtransLib, make_params, setup
from oracle.oracle import Oracle
n = 5000; nx = int(sys.argv[1]) if len(sys.argv) > 1 else 48
w = World()
prm = make_params(w, n, HTurbOn=0, VTurbOn=0, Behavior=0, settlementon=0, mortality=0)
g, o = LtransLib(), Oracle()
setup(g, w, prm, n); setup(o, w, prm, n); o.set_threads(os.cpu_count() or 1)
L = float(max(np.ptp(w.x_r), np.ptp(w.y_r))); H = float(w.h.max())
stepIT = prm.dt // prm.idt
def report(tag):
    fg, fo = g.fetch(("x", "y", "z", "status", "r_ele")), o.fetch(("x", "y", "z", "status", "r_ele"))
    dxy = np.maximum(np.abs(fg["x"] - fo["x"]), np.abs(fg["y"] - fo["y"])) / L; dz = np.abs(fg["z"] - fo["z"]) / H
    q = lambda a: " ".join("%.1e" % v for v in np.quantile(a, [0.5, 0.9, 0.99, 0.999, 1.0]))
    sg, so = g.fetch_sigerr(), o.fetch_sigerr()
    same = sg == so                                  # identical SigErr fall-back counts (necessary for identical histories)
    d = np.maximum(dxy, dz)
    print("%-18s xy q50/90/99/99.9/max %s | z %s | >1e-9: xy %d z %d | status diff %d r_ele diff %d" % (
        tag, q(dxy), q(dz), (dxy > 1e-9).sum(), (dz > 1e-9).sum(), (fg["status"] != fo["status"]).sum(), (fg["r_ele"] != fo["r_ele"]).sum()), flush=True)
    print("%-18s SigErr fall-backs gpu %d oracle %d | particles with equal counts %d: max diff %.2e | with unequal counts %d: max diff %.2e" % (
        "", sg.sum(), so.sum(), same.sum(), d[same].max() if same.any() else 0.0, (~same).sum(), d[~same].max() if (~same).any() else 0.0), flush=True)
for p in range(1, nx + 1):
    if p > 2:
        rec = w.record(p)
        for e in (g, o): e.push_hydro(rec); e.rotate_hydro()
    for it in range(1, stepIT + 1):
        g.step(p, it); o.step(p, it)
        if p == 1 and it == 1: report("1 internal step")
    if p in (1, 4, 12, 24, nx): report("%d external steps" % p)
