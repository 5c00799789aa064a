"""Shared harness of the parity tests: build a synthetic world, feed the SAME arrays to
the CUDA library (through its C ABI) and to the CPU oracle, run the reference's
external / internal loop (LTRANS.f90:156-161, 548-614) on both."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
import ltrans_b200  # noqa: E402,F401
from ltrans_b200.host.binding import LtransLib, Params  # noqa: E402
from ltrans_b200.host.world import World  # noqa: E402

SMALL = dict(ni=40, nj=36, us=10)


def make_params(world, n, **kw):
    base = dict(numpar=n, us=world.us, ws=world.ws, ErrorFlag=1, TrackCollisions=1)
    base.update(kw)
    return Params.shipped(**base)


def setup(lib, world, prm, n, seed=1234, dob_max=0.0, habitat=None, locate=True, first_id=1, sl=None,
          dtype=np.float32):
    lib.create(prm)
    lib.set_grid(world.grid())
    lib.set_bounds(world.bounds())
    if prm.settlementon:
        lib.set_habitat(habitat if habitat is not None else world.habitat())
    x, y, z, dob, r, u, v = world.seed_particles(n, seed=seed, dob_max=dob_max)
    if sl is not None:
        x, y, z, dob, r, u, v = (a[sl] for a in (x, y, z, dob, r, u, v))
    if locate:
        lib.set_particles(x, y, z, dob, None, r, u, v, first_id=first_id)
    else:
        lib.set_particles(x, y, z, dob, None, None, None, None, first_id=first_id)
    for k in range(3):
        lib.push_hydro(world.record(k, dtype))
    return x, y, z


def run(lib, world, nexternal, dtype=np.float32, nint=None):
    """run_LTRANS loop: external steps p = 1..nexternal, updateHydro from p = 3 on."""
    rc = 0
    for p in range(1, nexternal + 1):
        if p > 2:
            lib.push_hydro(world.record(p, dtype))
            lib.rotate_hydro()
        if nint is None:
            rc = lib.run_external(p)
        else:
            for it in range(1, nint + 1):
                rc = lib.step(p, it)
                if rc:
                    break
        if rc:
            break
    return lib.sync()


def rel_err(a, b, scale):
    return float(np.max(np.abs(a - b)) / scale)


def compare(fg, fo, world, tol=1e-9, ztol=None):
    """Trajectory parity: positions within tol relative to the domain extent (x, y) /
    water depth (z); identical element ids, status and event counters."""
    L = float(world.x_r.max() - world.x_r.min())
    H = float(world.h.max())
    out = dict(ex=rel_err(fg["x"], fo["x"], L), ey=rel_err(fg["y"], fo["y"], L),
               ez=rel_err(fg["z"], fo["z"], H))
    for k in ("r_ele", "u_ele", "v_ele", "status", "hitBottom", "hitLand", "endpoly"):
        out["n_" + k] = int(np.sum(fg[k] != fo[k]))
    out["eage"] = float(np.max(np.abs(fg["age"] - fo["age"])))
    return out


def assert_parity(res, tol=1e-9):
    assert res["ex"] <= tol and res["ey"] <= tol and res["ez"] <= tol, res
    for k, v in res.items():
        if k.startswith("n_"):
            assert v == 0, res
    assert res["eage"] == 0.0, res
