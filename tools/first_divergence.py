"""Step the CUDA library and the CPU oracle side by side (BASELINE configs[0]: 5,000 passive
particles, advection only) and report the first internal step at which any particle differs by
more than 1e-9 (relative to the domain / depth scale), with that particle's state on both sides."""
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
prev = None
for p in range(1, nx + 1):
    if p > 2:
        rec = w.record(p)
        for e in (g, o): e.push_hydro(rec); e.rotate_hydro()
    for it in range(1, stepIT + 1):
        g.step(p, it); o.step(p, it)
        fg, fo = g.fetch(("x", "y", "z", "status", "r_ele", "hitBottom", "hitLand")), o.fetch(("x", "y", "z", "status", "r_ele", "hitBottom", "hitLand"))
        d = np.maximum(np.maximum(np.abs(fg["x"] - fo["x"]) / L, np.abs(fg["y"] - fo["y"]) / L), np.abs(fg["z"] - fo["z"]) / H)
        k = int(np.argmax(d))
        if d[k] > 1e-9:
            print("first divergence at p=%d it=%d particle %d (1-based %d): rel diff %.3e" % (p, it, k, k + 1, d[k]))
            for name, f in (("gpu", fg), ("ora", fo)):
                print("  %s x %.9f y %.9f z %.12f status %d r_ele %d hitB %d hitL %d" % (name, f["x"][k], f["y"][k], f["z"][k], f["status"][k], f["r_ele"][k], f["hitBottom"][k], f["hitLand"][k]))
            if prev is not None:
                for name, f in (("gpu prev", prev[0]), ("ora prev", prev[1])):
                    print("  %s x %.9f y %.9f z %.12f status %d r_ele %d" % (name, f["x"][k], f["y"][k], f["z"][k], f["status"][k], f["r_ele"][k]))
                print("  prev rel diff %.3e ; depth at r_ele nodes:" % (max(abs(prev[0]["x"][k] - prev[1]["x"][k]) / L, abs(prev[0]["z"][k] - prev[1]["z"][k]) / H)),
                      w.h.ravel()[w.grid()["RE"][fg["r_ele"][k] - 1] - 1])
            print("  max diff over the others: %.3e" % np.sort(d)[-2])
            sys.exit(0)
        prev = (fg, fo)
print("no divergence > 1e-9 in %d external steps; max %.3e" % (nx, d.max()))
