"""Wall-clock breakdown of one end-to-end bench step (host-side costs around run_external)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import World, LtransLib, make_params
n = 1_000_000
w = World(); prm = make_params(w, n, Behavior=0, settlementon=0, mortality=0, TrackCollisions=0)
g = LtransLib().create(prm); g.set_grid(w.grid()); g.set_bounds(w.bounds())
x, y, z, dob, r, u, v = w.seed_particles(n); g.set_particles(x, y, z, dob, None, r, u, v)
recs = [w.record(k) for k in range(8)]
for k in range(3): g.push_hydro(recs[k])
g.run_external(1); g.run_external(2); g.sync()
def t(f, *a):
    t0 = time.perf_counter(); r = f(*a); g.sync(); return (time.perf_counter() - t0) * 1e3, r
for p in (3, 4, 5):
    a, _ = t(g.push_hydro, recs[p]); b, _ = t(g.rotate_hydro)
    t0 = time.perf_counter(); g.run_external(p); enq = (time.perf_counter() - t0) * 1e3; g.sync(); run = (time.perf_counter() - t0) * 1e3
    c, _ = t(g.fetch, ("x", "y", "z", "status"))
    print("p=%d push %.2f ms, rotate %.2f ms, run_external enqueue %.2f ms / total %.2f ms, fetch %.2f ms" % (p, a, b, enq, run, c))

import threading, subprocess
def loop(k):
    global p
    g.sync(); t0 = time.perf_counter()
    for _ in range(k):
        p += 1
        g.rotate_hydro(); g.push_hydro(recs[(p + 1) % 8]); g.run_external(p); g.fetch(("x", "y", "z", "status"))
    g.sync(); return (time.perf_counter() - t0) * 1e3 / k
p = 5
g.push_hydro(recs[6])
print("e2e loop, no sampler: %.2f ms/step" % loop(4))
stop = threading.Event()
def sampler():
    while not stop.is_set():
        subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm", "--format=csv,noheader"], capture_output=True); stop.wait(0.2)
th = threading.Thread(target=sampler); th.start()
print("e2e loop, nvidia-smi every 0.2 s: %.2f ms/step" % loop(4))
stop.set(); th.join()
