"""ctypes binding of the C ABI declared in include/ltrans_b200.h.

This is the Python stand-in for the Fortran ``INTERFACE ... BIND(C)`` block shown
in INTEGRATION.md: every method is a 1:1 call of one ``ltgpu_*`` entry point with
host (NumPy) buffers.  There is NO CPU fallback: if the CUDA library is missing
``LtransLib()`` raises, and ``create`` fails when no CUDA device is usable.

The class is generic over (library path, symbol prefix) only so that the test
oracle (oracle/oracle.py, prefix ``ora_``) can be driven with the same arrays;
the product never loads anything but ``libltrans_b200.so``.
"""
import ctypes as C
import os

import numpy as np

LTGPU_OK, LTGPU_E_ARG, LTGPU_E_CUDA, LTGPU_E_NODEVICE, LTGPU_E_PARTICLE, LTGPU_W_EVENTS_LOST = 0, 1, 2, 3, 4, 5
LTGPU_F32, LTGPU_F64 = 4, 8
LTGPU_RNG_PHILOX = 1

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.environ.get("LTRANS_B200_LIB") or os.path.join(os.path.dirname(_HERE), "csrc", "libltrans_b200.so")


class LtransError(RuntimeError):
    pass


class Params(C.Structure):
    """``ltgpu_params`` (include/ltrans_b200.h); names follow reference LTRANS.h:45-269."""
    _fields_ = [
        ("numpar", C.c_int32), ("dt", C.c_int32), ("idt", C.c_int32), ("us", C.c_int32),
        ("ws", C.c_int32), ("hc", C.c_float), ("Vtransform", C.c_int32), ("z0", C.c_double),
        ("HTurbOn", C.c_int32), ("VTurbOn", C.c_int32), ("ConstantHTurb", C.c_double),
        ("Behavior", C.c_int32), ("OpenOceanBoundary", C.c_int32), ("mortality", C.c_int32),
        ("settlementon", C.c_int32),
        ("deadage", C.c_double), ("pediage", C.c_double), ("swimstart", C.c_double),
        ("swimslow", C.c_double), ("swimfast", C.c_double),
        ("Sgradient", C.c_double), ("sink", C.c_double), ("Hswimspeed", C.c_double),
        ("Swimdepth", C.c_double),
        ("twistart", C.c_double), ("twiend", C.c_double), ("daylength", C.c_double),
        ("Em", C.c_double), ("Kd", C.c_double), ("thresh", C.c_double),
        ("holesExist", C.c_int32), ("seed", C.c_int32), ("PI", C.c_double),
        ("ErrorFlag", C.c_int32), ("SaltTempOn", C.c_int32), ("TrackCollisions", C.c_int32),
        ("FreeSlip", C.c_int32), ("rng_mode", C.c_int32), ("field_dtype", C.c_int32),
        ("vturb_window_sigs", C.c_int32), ("vturb_fp32_walk", C.c_int32),
    ]

    @classmethod
    def shipped(cls, **over):
        """Values of the shipped LTRANS.data (reference Model/LTRANS.data:21-254)."""
        p = cls(numpar=608, dt=3600, idt=120, us=20, ws=21, hc=0.2, Vtransform=1, z0=0.0005,
                HTurbOn=1, VTurbOn=1, ConstantHTurb=1.0, Behavior=4, OpenOceanBoundary=1,
                mortality=1, settlementon=1, deadage=367200.0, pediage=302400.0, swimstart=0.0,
                swimslow=0.005, swimfast=0.005, Sgradient=1.0, sink=-0.0003, Hswimspeed=0.9,
                Swimdepth=2.0, twistart=4.801821, twiend=19.19956, daylength=14.39774,
                Em=1814.328, Kd=1.07, thresh=0.0166, holesExist=1, seed=9,
                PI=3.14159265358979, ErrorFlag=0, SaltTempOn=0, TrackCollisions=0, FreeSlip=0,
                rng_mode=LTGPU_RNG_PHILOX, field_dtype=LTGPU_F32)
        for k, v in over.items():
            if not hasattr(p, k):
                raise AttributeError(k)
            setattr(p, k, v)
        return p


class Event(C.Structure):
    _fields_ = [("particle", C.c_int32), ("code", C.c_int32), ("time", C.c_double)]


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class LtransLib:
    """One context = one GPU = one rank (ltgpu_create .. ltgpu_destroy)."""

    def __init__(self, path=None, prefix="ltgpu_"):
        path = path or DEFAULT_LIB
        if not os.path.exists(path):
            raise LtransError(
                f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`."
                " There is no CPU fallback for the particle step.")
        self.lib = C.CDLL(path)
        self.prefix = prefix
        self.ctx = C.c_void_p()
        self.n = 0
        self.events_lost = 0
        self._keep = []       # host buffers that must outlive an async push
        last = getattr(self.lib, prefix + "last_error", None)
        if last is not None:
            last.restype = C.c_char_p
        lc = getattr(self.lib, prefix + "launch_count", None)
        if lc is not None:
            lc.restype = C.c_int64
        st = getattr(self.lib, prefix + "stream", None)
        if st is not None:
            st.restype = C.c_void_p

    # -- plumbing -----------------------------------------------------------
    def _fn(self, name):
        f = getattr(self.lib, self.prefix + name)
        return f

    def _check(self, rc, what):
        if rc != LTGPU_OK:
            msg = ""
            last = getattr(self.lib, self.prefix + "last_error", None)
            if last is not None and self.ctx:
                m = last(self.ctx)
                msg = m.decode() if m else ""
            raise LtransError(f"{self.prefix}{what} failed: status {rc} {msg}")

    # -- lifecycle ----------------------------------------------------------
    def create(self, prm, device=0):
        self.prm = prm
        if self.prefix == "ltgpu_":
            rc = self._fn("create")(C.byref(prm), C.c_int32(device), C.byref(self.ctx))
        else:
            rc = self._fn("create")(C.byref(prm), C.byref(self.ctx))
        self._check(rc, "create")
        return self

    def destroy(self):
        if self.ctx:
            self._fn("destroy")(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    # -- uploads ------------------------------------------------------------
    def set_grid(self, g):
        """g: dict from host.world / the Fortran host's initGrid tables."""
        a = {k: _f64(g[k]) for k in ("rx", "ry", "ux", "uy", "vx", "vy", "depth", "angle",
                                     "SC", "CS", "SCW", "CSW")}
        b = {k: _i32(g[k]) for k in ("rho_mask", "u_mask", "v_mask", "RE", "UE", "VE",
                                     "rAdj", "uAdj", "vAdj")}
        rc = self._fn("set_grid")(
            self.ctx, C.c_int32(g["vi"]), C.c_int32(g["uj"]), C.c_int32(g["ui"]), C.c_int32(g["vj"]),
            _p(a["rx"]), _p(a["ry"]), _p(a["ux"]), _p(a["uy"]), _p(a["vx"]), _p(a["vy"]),
            _p(a["depth"]), _p(a["angle"]), _p(b["rho_mask"]), _p(b["u_mask"]), _p(b["v_mask"]),
            _p(a["SC"]), _p(a["CS"]), _p(a["SCW"]), _p(a["CSW"]),
            _p(b["RE"]), _p(b["UE"]), _p(b["VE"]),
            C.c_int32(g["nRE"]), C.c_int32(g["nUE"]), C.c_int32(g["nVE"]),
            _p(b["rAdj"]), _p(b["uAdj"]), _p(b["vAdj"]))
        self._check(rc, "set_grid")

    def set_bounds(self, b):
        bnd_x, bnd_y = _f64(b["bnd_x"]), _f64(b["bnd_y"])
        land = _i32(b["land"])
        bx, by, hx, hy = _f64(b["bx"]), _f64(b["by"]), _f64(b["hx"]), _f64(b["hy"])
        hid = _i32(b["hid"])
        rc = self._fn("set_bounds")(
            self.ctx, C.c_int32(len(land)), _p(bnd_x), _p(bnd_y), _p(land),
            C.c_int32(len(bx)), _p(bx), _p(by), C.c_int32(len(hx)), _p(hx), _p(hy), _p(hid))
        self._check(rc, "set_bounds")

    def set_habitat(self, h):
        polys, holes = _f64(h["polys"]), _f64(h["holes"])
        ii = {k: _i32(h[k]) for k in ("poly_id", "poly_start", "poly_size", "hole_id", "hole_start",
                                      "hole_size", "elepoly_ptr", "elepoly_idx", "polyhole_ptr",
                                      "polyhole_idx")}
        pm, hm = _f64(h["poly_maxdis"]), _f64(h["hole_maxdis"])
        rc = self._fn("set_habitat")(
            self.ctx, C.c_int32(h["pedges"]), _p(polys), C.c_int32(h["hedges"]), _p(holes),
            C.c_int32(len(ii["poly_id"])), _p(ii["poly_id"]), _p(ii["poly_start"]), _p(ii["poly_size"]), _p(pm),
            C.c_int32(len(ii["hole_id"])), _p(ii["hole_id"]), _p(ii["hole_start"]), _p(ii["hole_size"]), _p(hm),
            _p(ii["elepoly_ptr"]), _p(ii["elepoly_idx"]), _p(ii["polyhole_ptr"]), _p(ii["polyhole_idx"]))
        self._check(rc, "set_habitat")

    def set_particles(self, x, y, z, dob, startpoly=None, r_ele=None, u_ele=None, v_ele=None, first_id=1):
        x, y, z, dob = _f64(x), _f64(y), _f64(z), _f64(dob)
        sp, r, u, v = _i32(startpoly), _i32(r_ele), _i32(u_ele), _i32(v_ele)
        self.n = len(x)
        rc = self._fn("set_particles")(
            self.ctx, C.c_int32(self.n), C.c_int64(first_id), _p(x), _p(y), _p(z), _p(dob),
            _p(sp), _p(r), _p(u), _p(v))
        self._check(rc, "set_particles")

    def push_hydro(self, rec):
        """rec: dict zeta,u,v,w,aks[,salt,temp] in ROMS memory order (node fastest)."""
        dt = rec["zeta"].dtype
        code = LTGPU_F32 if dt == np.float32 else LTGPU_F64
        arrs = [np.ascontiguousarray(rec[k], dtype=dt) if rec.get(k) is not None else None
                for k in ("zeta", "u", "v", "w", "aks", "salt", "temp")]
        self._keep = arrs          # source must stay valid until the next step / push
        rc = self._fn("push_hydro")(self.ctx, C.c_int32(code), *[_p(a) for a in arrs])
        self._check(rc, "push_hydro")

    def rotate_hydro(self):
        self._check(self._fn("rotate_hydro")(self.ctx), "rotate_hydro")

    # -- stepping -----------------------------------------------------------
    def step(self, p, it):
        rc = self._fn("step")(self.ctx, C.c_int32(p), C.c_int32(it))
        if rc not in (LTGPU_OK, LTGPU_E_PARTICLE):
            self._check(rc, "step")
        return rc

    def run_external(self, p):
        rc = self._fn("run_external")(self.ctx, C.c_int32(p))
        if rc not in (LTGPU_OK, LTGPU_E_PARTICLE):
            self._check(rc, "run_external")
        return rc

    def screen_initial(self):
        """-> (rc, counts[5] for codes 11..15, lowest offending particle id or 0)"""
        c = (C.c_int64 * 5)()
        bad = C.c_int64(0)
        rc = self._fn("screen_initial")(self.ctx, c, C.byref(bad))
        if rc not in (LTGPU_OK, LTGPU_E_PARTICLE):
            self._check(rc, "screen_initial")
        return rc, np.array(list(c), dtype=np.int64), bad.value

    def sync(self):
        bad = C.c_int32(0)
        rc = self._fn("sync")(self.ctx, C.byref(bad))
        if rc not in (LTGPU_OK, LTGPU_E_PARTICLE):
            self._check(rc, "sync")
        return rc, bad.value

    # -- results ------------------------------------------------------------
    def fetch(self, fields=("x", "y", "z", "age", "status", "salt", "temp", "hitBottom", "hitLand",
                            "endpoly", "lifespan", "r_ele", "u_ele", "v_ele"), out=None):
        """`out` = {field: preallocated array}: filled in place; page-locked arrays (e.g.
        torch.empty(..., pin_memory=True).numpy()) receive the device copy directly."""
        n = self.n
        spec = [("x", np.float64), ("y", np.float64), ("z", np.float64), ("age", np.float64),
                ("status", np.int32), ("salt", np.float64), ("temp", np.float64),
                ("hitBottom", np.int32), ("hitLand", np.int32), ("endpoly", np.int32),
                ("lifespan", np.float64), ("r_ele", np.int32), ("u_ele", np.int32), ("v_ele", np.int32)]
        given = out or {}
        for k, t in spec:
            if k in given and (given[k].dtype != t or given[k].shape != (n,) or not given[k].flags.c_contiguous):
                raise ValueError("fetch: out[%r] must be a contiguous %s array of %d" % (k, np.dtype(t).name, n))
        out = {k: (given[k] if k in given else np.empty(n, dtype=t)) if k in fields else None for k, t in spec}
        rc = self._fn("fetch")(self.ctx, *[_p(out[k]) for k, _ in spec])
        self._check(rc, "fetch")
        return {k: v for k, v in out.items() if v is not None}

    def fetch_lonlat(self, proj, spherical=True):
        """(lon, lat) of every particle, converted on the device (x2lon / y2lat)"""
        lon, lat = np.zeros(self.n), np.zeros(self.n)
        rc = self._fn("fetch_lonlat")(self.ctx, C.c_int32(1 if spherical else 0), C.c_double(proj.lonmin), C.c_double(proj.latmin),
                                      C.c_double(proj.R), _p(lon), _p(lat))
        self._check(rc, "fetch_lonlat")
        return lon, lat

    def fetch_sigerr(self):
        """diagnostic: SigErr (linint) fall-backs per particle so far"""
        c = np.zeros(self.n, dtype=np.int32)
        self._check(self._fn("fetch_sigerr")(self.ctx, _p(c)), "fetch_sigerr")
        return c

    def reset_hits(self):
        self._check(self._fn("reset_hits")(self.ctx), "reset_hits")

    def stats(self):
        c = (C.c_int64 * 8)()
        self._check(self._fn("stats")(self.ctx, c), "stats")
        return np.array(list(c), dtype=np.int64)

    def drain_events(self, cap=4096, everything=False):
        """One ltgpu_drain_events call (at most `cap` events), or with everything=True repeated
        calls until the log is empty.  An overflow of the device log (LTGPU_W_EVENTS_LOST) is
        not an error here: the count is kept in `self.events_lost`."""
        out = []
        while True:
            buf = (Event * cap)()
            n = C.c_int32(0)
            rc = self._fn("drain_events")(self.ctx, buf, C.c_int32(cap), C.byref(n))
            if rc == LTGPU_W_EVENTS_LOST:
                self.events_lost = self.lost_events()
            else:
                self._check(rc, "drain_events")
            out += [(buf[i].particle, buf[i].code, buf[i].time) for i in range(n.value)]
            if not everything or n.value < cap:
                return out

    def lost_events(self):
        f = getattr(self.lib, self.prefix + "events_lost", None)
        if f is None:
            return 0
        lost = C.c_int64(0)
        self._check(f(self.ctx, C.byref(lost)), "events_lost")
        return int(lost.value)

    # -- device-only helpers (bench / NCCL gather) ---------------------------
    def timer_start(self):
        self._check(self._fn("timer_start")(self.ctx), "timer_start")

    def timer_stop(self):
        ms = C.c_float(0)
        self._check(self._fn("timer_stop")(self.ctx, C.byref(ms)), "timer_stop")
        return ms.value

    def kernel_times(self, enable=True):
        ms = (C.c_float * 4)()
        steps = C.c_int64(0)
        self._check(self._fn("kernel_times")(self.ctx, C.c_int32(1 if enable else 0), ms, C.byref(steps)), "kernel_times")
        return [float(v) for v in ms], int(steps.value)

    def launch_count(self):
        return int(self._fn("launch_count")(self.ctx))

    def device_ptr(self, which):
        p = C.c_void_p()
        self._check(self._fn("device_ptr")(self.ctx, C.c_int32(which), C.byref(p)), "device_ptr")
        return p.value

    def export_device(self, which, dst_ptr):
        """column `which` (0=x 1=y 2=z 3=age f64, 4=status i32) in particle order into device memory"""
        self._check(self._fn("export_device")(self.ctx, C.c_int32(which), C.c_void_p(dst_ptr)), "export_device")

    def fp64_peak(self):
        t = C.c_double(0)
        self._check(self._fn("fp64_peak")(self.ctx, C.byref(t)), "fp64_peak")
        return t.value

    def stream(self):
        return int(self._fn("stream")(self.ctx) or 0)
