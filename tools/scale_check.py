"""Gulf-scale shape check (BASELINE configs[3]/[4] per-GPU share): 1024x768 rho grid, us 36 /
ws 37, 12.5 M buoyant particles (Behavior 6, sink > 0), HTurb + VTurb, open boundary, one GPU.
Not a bench line: it verifies that the per-GPU shard of the 100 M-particle configuration fits
and runs, and prints its device time and memory."""
import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import World, LtransLib, make_params
import torch

# python tools/scale_check.py [particles] [gulf|oyster]
#   gulf   (default) BASELINE configs[3]/[4] per-GPU share: 1024x768x36, buoyant Behavior 6, open boundary
#   oyster BASELINE configs[2]: Chesapeake-scale 120x80x20, Behavior 4, 64 settlement polygons with holes, mortality
which = sys.argv[2] if len(sys.argv) > 2 else "gulf"
n = int(sys.argv[1]) if len(sys.argv) > 1 else (12_500_000 if which == "gulf" else 10_000_000)
t0 = time.time()
if which == "gulf":
    w = World(ni=1024, nj=768, us=36, hmin=50.0, hmax=3000.0, dlon=0.02, dlat=0.018, speed=0.9)
    prm = make_params(w, n, Behavior=6, sink=0.002, settlementon=0, mortality=0, TrackCollisions=0, ErrorFlag=3)
else:
    w = World(ni=120, nj=80, us=20, dlon=0.02, dlat=0.018)
    prm = make_params(w, n, Behavior=4, settlementon=1, holesExist=1, mortality=1, TrackCollisions=0, ErrorFlag=3,
                      pediage=3600.0, deadage=3 * 3600.0)
g = LtransLib().create(prm)
g.set_grid(w.grid()); g.set_bounds(w.bounds())
if prm.settlementon:
    g.set_habitat(w.habitat(npoly=64))
x, y, z, dob, r, u, v = w.seed_particles(n)
t1 = time.time()
g.set_particles(x, y, z, dob, None, None, None, None)             # located on the device (element buckets)
t2 = time.time()
rc, counts, bad = g.screen_initial()
t3 = time.time()
e = g.fetch(("r_ele", "u_ele", "v_ele"))
print("device locate + upload %.2f s, start-up screen %.3f s, elements == host locate: %s, screened %s" % (
    t2 - t1, t3 - t2, bool(np.array_equal(e["r_ele"], r) and np.array_equal(e["u_ele"], u) and np.array_equal(e["v_ele"], v)),
    counts.tolist()), flush=True)
recs = [w.record(k) for k in range(4)]
for k in range(3):
    g.push_hydro(recs[k])
g.sync()
print("setup %.1f s, device memory used %.2f GB" % (time.time() - t0, (torch.cuda.mem_get_info()[1] - torch.cuda.mem_get_info()[0]) / 1e9), flush=True)
stepIT = prm.dt // prm.idt
g.run_external(1); g.sync()
g.timer_start(); g.run_external(2); ms = g.timer_stop()
print("external step 2: %.1f ms -> %.3e particle-steps/s" % (ms, n * stepIT / (ms * 1e-3)))
g.push_hydro(recs[3]); g.rotate_hydro()
g.kernel_times(True); g.run_external(3); km, ks = g.kernel_times(False)
print("kernels ms per internal step:", [round(v / ks, 2) for v in km[:3]])
st = g.stats(); f = g.fetch(("z", "status"))
print("stats", st.tolist(), "finite", bool(np.isfinite(f["z"]).all()), "events", g.drain_events(5)[:5])
