"""Index-assertion sweep with a -DLT_DEBUG_TRACE build (compute-sanitizer is closed on the pool):
runs a spread of configurations and prints the highest source line whose LT_ASSERT failed (0 = none).
  nvcc ... -DLT_DEBUG_TRACE -o dbg/libltrans_dbg.so ltrans_b200.cu
  LTRANS_B200_LIB=$PWD/dbg/libltrans_dbg.so python tools/assert_run.py"""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import SMALL, World, LtransLib, make_params, setup, run

GULF = dict(ni=48, nj=40, us=36, hmin=40.0, hmax=600.0, dlon=0.02, dlat=0.018, speed=0.9)
CASES = [
    ("passive, turbulence", SMALL, 4000, 3, dict(Behavior=0, settlementon=0, mortality=0, HTurbOn=1, VTurbOn=1)),
    ("oyster, settlement, holes", SMALL, 4000, 4, dict(Behavior=4, HTurbOn=1, VTurbOn=1, pediage=3600.0, deadage=9000.0)),
    ("behaviour 7 + SaltTemp + FreeSlip", SMALL, 3000, 3, dict(Behavior=7, HTurbOn=1, VTurbOn=1, SaltTempOn=1, FreeSlip=1, settlementon=0)),
    ("gulf-like buoyant, ws 37", GULF, 3000, 2, dict(Behavior=6, sink=0.002, HTurbOn=1, VTurbOn=1, settlementon=0, OpenOceanBoundary=1)),
    ("shallow, us 5 (p2 < window)", dict(ni=40, nj=36, us=5, hmin=1.0, hmax=3.0), 3000, 2, dict(Behavior=0, HTurbOn=1, VTurbOn=1, settlementon=0)),
    ("strong currents, ErrorFlag 3", dict(SMALL, speed=3.0), 4000, 3, dict(Behavior=0, HTurbOn=1, VTurbOn=1, settlementon=0, ErrorFlag=3, ConstantHTurb=50.0)),
]
worst = 0
for name, wk, n, nx, kw in CASES:
    w = World(**wk)
    g = LtransLib()
    dbg = getattr(g.lib, "ltgpu_debug_counters", None)
    if dbg is None:
        sys.exit("not a debug build: set LTRANS_B200_LIB to a -DLT_DEBUG_TRACE library")
    prm = make_params(w, n, **kw)
    setup(g, w, prm, n, locate=False)
    rc = run(g, w, nx)
    c = (ctypes.c_ulonglong * 88)(); dbg(g.ctx, c, 1)
    f = g.fetch(("z", "status"))
    print("%-36s rc %s  failed-assert line %d  finite %s  stats %s" % (name, rc, c[7], bool(np.isfinite(f["z"]).all()), g.stats().tolist()), flush=True)
    worst = max(worst, int(c[7]))
sys.exit(1 if worst else 0)
