"""Longer run of the benchmark configuration (1 M turbulent particles, N external steps) checking
that the state stays finite, statistics are plausible and per-step time is flat (no stragglers)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import World, LtransLib, make_params
n = 1_000_000; nx = int(sys.argv[1]) if len(sys.argv) > 1 else 24
w = World(); prm = make_params(w, n, Behavior=0, settlementon=0, mortality=0, TrackCollisions=1, ErrorFlag=3)
g = LtransLib().create(prm); g.set_grid(w.grid()); g.set_bounds(w.bounds())
x, y, z, dob, r, u, v = w.seed_particles(n); g.set_particles(x, y, z, dob, None, r, u, v)
for k in range(3): g.push_hydro(w.record(k))
ms = []
for p in range(1, nx + 1):
    if p > 2: g.push_hydro(w.record(p)); g.rotate_hydro()
    g.timer_start(); g.run_external(p); ms.append(g.timer_stop())
rc = g.sync(); f = g.fetch(("x", "y", "z", "status")); st = g.stats(); sg = g.fetch_sigerr()
act = f["status"] == 0
print("rc", rc, "ms per external step: min %.1f median %.1f max %.1f" % (min(ms[2:]), float(np.median(ms[2:])), max(ms[2:])))
print("stats", st.tolist(), "finite", bool(np.isfinite(f["x"]).all() and np.isfinite(f["z"]).all()),
      "z range of active [%.2f, %.2f]" % (f["z"][act].min(), f["z"][act].max()), "SigErr fall-backs per particle-step %.2e" % (sg.sum() / (n * 30.0 * nx)),
      "events", len(g.drain_events(1 << 20)))
