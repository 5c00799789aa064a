"""How often a spline interval's T = max(D1/D2, D2/D1) falls in the band where the reference's
convexity Newton loop can cycle (SigErr).  Needs a -DLT_DEBUG_TRACE library (LTRANS_B200_LIB)."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import World, LtransLib, make_params, setup, run
for name, kw in (("advection only", dict(HTurbOn=0, VTurbOn=0)), ("HTurb + VTurb", dict(HTurbOn=1, VTurbOn=1))):
    w = World(); n = 20000
    g = LtransLib(); prm = make_params(w, n, Behavior=0, settlementon=0, mortality=0, **kw)
    setup(g, w, prm, n); run(g, w, 2)
    c = (ctypes.c_ulonglong * 88)(); g.lib.ltgpu_debug_counters(g.ctx, c, 1)
    ps = n * 60.0
    print("%-16s per particle-step: intervals classified %.1f, convexity solves %.2f, T in (2.0245,2.047) %.4f, T in (2.02,2.10) %.4f, Newton cycles %.2e" % (
        name, c[16 + 7] / ps, c[16 + 4] / ps, c[16 + 5] / ps, c[16 + 6] / ps, c[1] / ps))
