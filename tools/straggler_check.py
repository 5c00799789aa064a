"""Per-external-step kernel times (and, with a -DLT_DEBUG_TRACE build named by LTRANS_B200_LIB,
the solver-cap census) for the particle set of bench rank R."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import World, LtransLib, make_params
n = 1_000_000; rank = int(sys.argv[1]) if len(sys.argv) > 1 else 1
w = World(); prm = make_params(w, n, Behavior=0, settlementon=0, mortality=0, TrackCollisions=0, vturb_window_sigs=int(os.environ.get('WINDOW_SIGS', '0')))
g = LtransLib().create(prm); g.set_grid(w.grid()); g.set_bounds(w.bounds())
x, y, z, dob, r, u, v = w.seed_particles(n, seed=1234 + rank); g.set_particles(x, y, z, dob, None, r, u, v, first_id=1 + rank * n)
recs = [w.record(k) for k in range(12)]
for k in range(3): g.push_hydro(recs[k])
dbg = getattr(g.lib, "ltgpu_debug_counters", None) if hasattr(g, "lib") else None
g.kernel_times(True)
for p in range(1, 5):
    if p > 2: g.push_hydro(recs[p]); g.rotate_hydro()
    g.run_external(p); ms, st = g.kernel_times(True)
    extra = ""
    if dbg:
        c = (ctypes.c_ulonglong * 88)(); dbg(g.ctx, c, 1); print("   up", list(c)[24:56]); print("   dn", list(c)[56:88]); extra = "  census " + str(list(c)[:7]) + " build " + str(list(c)[16:20])
        if c[3]: extra += "\n   case " + " ".join(np.array(list(c)[8:], np.uint64).view(np.float64).astype(float).__iter__().__class__ and [float.hex(float(v)) for v in np.array(list(c)[8:], np.uint64).view(np.float64)])
    print("p=%d  advect %.1f  vturb %.1f (walk %.1f)  finish %.1f  (ms per external step)%s" % (p, ms[0], ms[1], ms[3], ms[2], extra))
