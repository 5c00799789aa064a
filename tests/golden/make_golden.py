"""Generates tests/golden/*.npz: final particle state of small seeded runs of the CPU
oracle.  The reference itself has no golden vectors and cannot run here (no Fortran
compiler / NetCDF), so these pin the ORACLE (regression) and give the GPU tests a
fixture that does not need the oracle library at run time.  Usage:
    python tests/golden/make_golden.py
Only turbulence-free cases are stored: with VTurb on, 1-ulp libm differences are
amplified by the random-displacement model (parity section)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from common import SMALL, World, make_params, setup, run  # noqa: E402

PASSIVE = dict(HTurbOn=0, VTurbOn=0, Behavior=0, settlementon=0, mortality=0)
CASES = {
    "passive": dict(n=300, next=3, kw=PASSIVE),
    "hturb_salttemp": dict(n=300, next=3, kw=dict(PASSIVE, HTurbOn=1, SaltTempOn=1)),
    "oyster4_settle": dict(n=300, next=3, kw=dict(Behavior=4, HTurbOn=1, VTurbOn=0, pediage=3600.0, deadage=9000.0)),
    "tidal7": dict(n=300, next=2, kw=dict(Behavior=7, HTurbOn=0, VTurbOn=0, settlementon=0, mortality=0)),
}
FIELDS = ("x", "y", "z", "age", "status", "salt", "temp", "hitBottom", "hitLand", "endpoly", "lifespan",
          "r_ele", "u_ele", "v_ele")


def run_case(name, factory):
    c = CASES[name]
    w = World(**SMALL)
    prm = make_params(w, c["n"], **c["kw"])
    lib = factory()
    setup(lib, w, prm, c["n"], dob_max=3600.0)
    run(lib, w, c["next"])
    f = lib.fetch()
    lib.destroy()
    return {k: f[k] for k in FIELDS}


if __name__ == "__main__":
    from oracle.oracle import Oracle
    for name in CASES:
        out = run_case(name, Oracle)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, {k: (v.dtype.str, float(np.abs(v).max())) for k, v in out.items()})
