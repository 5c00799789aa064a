"""SigErr (linint) fall-back rates, CUDA vs oracle, with and without turbulence (same Philox stream)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import World, LtransLib, make_params, setup, run
from oracle.oracle import Oracle
w = World(); n = 20000
for name, kw in (("advection only", dict(HTurbOn=0, VTurbOn=0)), ("HTurb only", dict(HTurbOn=1, VTurbOn=0)), ("HTurb + VTurb", dict(HTurbOn=1, VTurbOn=1)),
                 ("HTurb + VTurb, window-only SIGS (opt-in)", dict(HTurbOn=1, VTurbOn=1, vturb_window_sigs=1))):
    prm = make_params(w, n, Behavior=0, settlementon=0, mortality=0, **kw)
    g, o = LtransLib(), Oracle()
    setup(g, w, prm, n); setup(o, w, prm, n); o.set_threads(os.cpu_count() or 1)
    run(g, w, 1); run(o, w, 1)
    sg, so = g.fetch_sigerr(), o.fetch_sigerr()
    fg, fo = g.fetch(("z",)), o.fetch(("z",))
    dz = np.abs(fg["z"] - fo["z"]) / float(w.h.max())
    clean = (sg == 0) & (so == 0)
    print("%-32s fall-backs per particle-step: cuda %.2e oracle %.2e | particles clean on both sides %.4f, their max dz %.1e, within 1e-9 overall %.4f" % (
        name, sg.sum() / (n * 30.0), so.sum() / (n * 30.0), clean.mean(), dz[clean].max(), (dz <= 1e-9).mean()), flush=True)
