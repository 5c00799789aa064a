"""ctypes wrapper of the CPU oracle (oracle/libltrans_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  Same call surface as
ltrans_b200.host.binding.LtransLib so the parity tests feed both with identical arrays,
but its own marshalling (see class Oracle).  PARITY UNPINNED BY THE REFERENCE (see ltrans_oracle.h).
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
import ltrans_b200  # noqa: E402,F401

LIB = os.path.join(_HERE, "libltrans_oracle.so")
ORA_RNG_PHILOX, ORA_RNG_MT = 1, 2


def build(force=False):
    if force or not os.path.exists(LIB):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return LIB


def _nd(dtype):
    return np.ctypeslib.ndpointer(dtype=dtype, flags="C_CONTIGUOUS")


class _Opt:
    """ndpointer that also accepts None (NULL)"""

    def __init__(self, dtype):
        self.base = _nd(dtype)

    def from_param(self, obj):
        return None if obj is None else self.base.from_param(obj)


F64, I32 = _nd(np.float64), _nd(np.int32)
OF64, OI32 = _Opt(np.float64), _Opt(np.int32)
_FETCH = (("x", np.float64), ("y", np.float64), ("z", np.float64), ("age", np.float64), ("status", np.int32),
          ("salt", np.float64), ("temp", np.float64), ("hitBottom", np.int32), ("hitLand", np.int32),
          ("endpoly", np.int32), ("lifespan", np.float64), ("r_ele", np.int32), ("u_ele", np.int32), ("v_ele", np.int32))


class Oracle:
    """The checker's own marshalling: argument types are declared once per entry point
    (numpy.ctypeslib.ndpointer enforces dtype and contiguity) instead of reusing the product's
    binding, so a column-order or dtype slip in ltrans_b200/host/binding.py is not shared by the
    two sides of a parity test.  Only the parameter struct (checked field by field against the C
    header in tests/test_host.py) and the world tables are common."""

    def __init__(self):
        build()
        L = self.lib = C.CDLL(LIB)
        self.ctx = C.c_void_p()
        self.n = 0
        self.events_lost = 0
        self._keep = None
        i32, i64, d, vp = C.c_int32, C.c_int64, C.c_double, C.c_void_p
        L.ora_set_grid.argtypes = [vp, i32, i32, i32, i32] + [F64] * 8 + [I32] * 3 + [F64] * 4 + [I32] * 3 + [i32] * 3 + [I32] * 3
        L.ora_set_bounds.argtypes = [vp, i32, F64, F64, I32, i32, F64, F64, i32, F64, F64, I32]
        L.ora_set_habitat.argtypes = [vp, i32, F64, i32, F64, i32, I32, I32, I32, F64, i32, I32, I32, I32, F64, I32, I32, I32, I32]
        L.ora_set_particles.argtypes = [vp, i32, i64, F64, F64, F64, F64, OI32, OI32, OI32, OI32]
        L.ora_push_hydro.argtypes = [vp, i32] + [vp] * 7
        L.ora_fetch.argtypes = [vp] + [OF64 if t == np.float64 else OI32 for _, t in _FETCH]
        L.ora_fetch_lonlat.argtypes = [vp, i32, d, d, d, F64, F64]
        L.ora_fetch_sigerr.argtypes = [vp, I32]
        L.ora_step.argtypes = [vp, i32, i32]; L.ora_run_external.argtypes = [vp, i32]
        L.ora_set_threads.argtypes = [vp, i32]; L.ora_set_rng.argtypes = [vp, i32]

    def _ok(self, rc, what, allow=()):
        if rc != 0 and rc not in allow:
            raise RuntimeError("ora_%s failed: status %d" % (what, rc))
        return rc

    def create(self, prm, device=0):
        self.prm = prm
        self._ok(self.lib.ora_create(C.byref(prm), C.byref(self.ctx)), "create")
        return self

    def destroy(self):
        if self.ctx:
            self.lib.ora_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    def set_threads(self, n):
        self.lib.ora_set_threads(self.ctx, n)

    def set_rng(self, mode):
        self.lib.ora_set_rng(self.ctx, mode)

    def set_grid(self, g):
        f = lambda k: np.ascontiguousarray(g[k], np.float64)
        i = lambda k: np.ascontiguousarray(g[k], np.int32)
        self._ok(self.lib.ora_set_grid(
            self.ctx, g["vi"], g["uj"], g["ui"], g["vj"], f("rx"), f("ry"), f("ux"), f("uy"), f("vx"), f("vy"), f("depth"), f("angle"),
            i("rho_mask"), i("u_mask"), i("v_mask"), f("SC"), f("CS"), f("SCW"), f("CSW"), i("RE"), i("UE"), i("VE"),
            g["nRE"], g["nUE"], g["nVE"], i("rAdj"), i("uAdj"), i("vAdj")), "set_grid")

    def set_bounds(self, b):
        f = lambda k: np.ascontiguousarray(b[k], np.float64)
        i = lambda k: np.ascontiguousarray(b[k], np.int32)
        self._ok(self.lib.ora_set_bounds(self.ctx, len(b["land"]), f("bnd_x"), f("bnd_y"), i("land"), len(b["bx"]), f("bx"), f("by"),
                                         len(b["hx"]), f("hx"), f("hy"), i("hid")), "set_bounds")

    def set_habitat(self, h):
        f = lambda k: np.ascontiguousarray(h[k], np.float64)
        i = lambda k: np.ascontiguousarray(h[k], np.int32)
        self._ok(self.lib.ora_set_habitat(
            self.ctx, h["pedges"], f("polys"), h["hedges"], f("holes"),
            len(h["poly_id"]), i("poly_id"), i("poly_start"), i("poly_size"), f("poly_maxdis"),
            len(h["hole_id"]), i("hole_id"), i("hole_start"), i("hole_size"), f("hole_maxdis"),
            i("elepoly_ptr"), i("elepoly_idx"), i("polyhole_ptr"), i("polyhole_idx")), "set_habitat")

    def set_particles(self, x, y, z, dob, startpoly=None, r_ele=None, u_ele=None, v_ele=None, first_id=1):
        f = lambda a: np.ascontiguousarray(a, np.float64)
        i = lambda a: None if a is None else np.ascontiguousarray(a, np.int32)
        self.n = len(x)
        self._ok(self.lib.ora_set_particles(self.ctx, self.n, first_id, f(x), f(y), f(z), f(dob), i(startpoly), i(r_ele), i(u_ele), i(v_ele)),
                 "set_particles")

    def push_hydro(self, rec):
        dt = rec["zeta"].dtype
        arrs = [np.ascontiguousarray(rec[k], dtype=dt) if rec.get(k) is not None else None for k in ("zeta", "u", "v", "w", "aks", "salt", "temp")]
        self._keep = arrs
        self._ok(self.lib.ora_push_hydro(self.ctx, 4 if dt == np.float32 else 8, *[None if a is None else a.ctypes.data for a in arrs]), "push_hydro")

    def rotate_hydro(self):
        self._ok(self.lib.ora_rotate_hydro(self.ctx), "rotate_hydro")

    def step(self, p, it):
        return self._ok(self.lib.ora_step(self.ctx, p, it), "step", allow=(4,))

    def run_external(self, p):
        return self._ok(self.lib.ora_run_external(self.ctx, p), "run_external", allow=(4,))

    def screen_initial(self):
        c = (C.c_int64 * 5)(); bad = C.c_int64(0)
        rc = self._ok(self.lib.ora_screen_initial(self.ctx, c, C.byref(bad)), "screen_initial", allow=(4,))
        return rc, np.array(list(c), dtype=np.int64), bad.value

    def sync(self):
        bad = C.c_int32(0)
        rc = self._ok(self.lib.ora_sync(self.ctx, C.byref(bad)), "sync", allow=(4,))
        return rc, bad.value

    def fetch(self, fields=tuple(k for k, _ in _FETCH), out=None):
        out = {k: (np.empty(self.n, dtype=t) if k in fields else None) for k, t in _FETCH}
        self._ok(self.lib.ora_fetch(self.ctx, *[out[k] for k, _ in _FETCH]), "fetch")
        return {k: v for k, v in out.items() if v is not None}

    def fetch_lonlat(self, proj, spherical=True):
        lon, lat = np.zeros(self.n), np.zeros(self.n)
        self._ok(self.lib.ora_fetch_lonlat(self.ctx, 1 if spherical else 0, proj.lonmin, proj.latmin, proj.R, lon, lat), "fetch_lonlat")
        return lon, lat

    def fetch_sigerr(self):
        c = np.zeros(self.n, dtype=np.int32)
        self._ok(self.lib.ora_fetch_sigerr(self.ctx, c), "fetch_sigerr")
        return c

    def reset_hits(self):
        self._ok(self.lib.ora_reset_hits(self.ctx), "reset_hits")

    def stats(self):
        c = (C.c_int64 * 8)()
        self._ok(self.lib.ora_stats(self.ctx, c), "stats")
        return np.array(list(c), dtype=np.int64)

    def drain_events(self, cap=4096, everything=False):
        from ltrans_b200.host.binding import Event
        out = []
        while True:
            buf = (Event * cap)(); n = C.c_int32(0)
            self._ok(self.lib.ora_drain_events(self.ctx, buf, cap, C.byref(n)), "drain_events")
            out += [(buf[k].particle, buf[k].code, buf[k].time) for k in range(n.value)]
            if not everything or n.value < cap:
                return out

    def lost_events(self):
        return 0

    def launch_count(self):
        return 0


def leaf():
    """Raw library handle with restypes set for the exported leaf functions."""
    build()
    lib = C.CDLL(LIB)
    d, i32, pd, pi = C.c_double, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_int32)
    lib.ora_polintd.restype = d; lib.ora_polintd.argtypes = [pd, pd, d]
    lib.ora_linint.argtypes = [pd, pd, i32, d, pd, pd]
    lib.ora_gridcell.restype = i32; lib.ora_gridcell.argtypes = [pd, pd, d, d]
    lib.ora_inpoly.restype = i32; lib.ora_inpoly.argtypes = [d, d, i32, pd, pd, i32]
    lib.ora_tspsi.argtypes = [i32, pd, pd, pd, pd, pi, pi]
    lib.ora_hval.restype = d; lib.ora_hval.argtypes = [d, i32, pd, pd, pd, pd, pi]
    lib.ora_hpval.restype = d; lib.ora_hpval.argtypes = [d, i32, pd, pd, pd, pd, pi]
    lib.ora_snhcsh.argtypes = [d, pd, pd, pd]
    lib.ora_slevel.restype = d; lib.ora_slevel.argtypes = [d, d, d, d, C.c_float, i32]
    lib.ora_mt_init_genrand.argtypes = [C.c_uint32]
    lib.ora_mt_init_by_array.argtypes = [C.POINTER(C.c_uint32), i32]
    lib.ora_mt_int32.restype = C.c_uint32
    lib.ora_mt_real1.restype = d; lib.ora_mt_real3.restype = d
    lib.ora_philox4x32_10.argtypes = [C.POINTER(C.c_uint32)] * 3
    lib.ora_intersect_reflect.restype = i32
    lib.ora_intersect_reflect.argtypes = [C.c_void_p, d, d, d, d, pd, pd, pd, pd, pi, pi]
    lib.ora_mbounds.restype = i32; lib.ora_mbounds.argtypes = [C.c_void_p, d, d]
    lib.ora_ibounds.restype = i32; lib.ora_ibounds.argtypes = [C.c_void_p, d, d, pd]
    lib.ora_settle_point.restype = i32; lib.ora_settle_point.argtypes = [C.c_void_p, i32, d, d, d]
    lib.ora_behave_case.restype = None
    lib.ora_behave_case.argtypes = [C.c_void_p, pd, d, C.POINTER(C.c_uint32)] + [d] * 8 + [i32, d, pd]
    lib.ora_find_currents_column.restype = None
    lib.ora_find_currents_column.argtypes = [i32, i32, d, d, pd, pd, pd, pd, pd, d, d, d, pd, pd, i32, i32, pd, pi]
    lib.ora_interp_quad.restype = d; lib.ora_interp_quad.argtypes = [pd, pd, pd, d, d, i32]
    lib.ora_wcts_profile.restype = d
    lib.ora_wcts_profile.argtypes = [pd] * 6 + [d, d, d, pd, pd, i32, i32, pi]
    lib.ora_vturb_column.restype = d
    lib.ora_vturb_column.argtypes = [i32, i32, i32, pd, pd, pd, pd, pd, pd, pd, pd, d, d, d, pd, pi]
    return lib


def dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))
