"""Parity of the CUDA path (called through the C ABI) against the CPU oracle on the same
seeded inputs, against the committed golden fixtures, and full-size properties.

Tolerances (BASELINE.json north_star): turbulence off -> positions within 1e-9 relative
(to the domain extent for x, y; to the maximum depth for z), identical rho/u/v element
ids, status, hit counters, settlement polygons and events.  Turbulence on -> the oracle is
fed the same Philox stream; the random-displacement model amplifies 1-ulp libm differences
(DESIGN.md), so same-stream parity is asserted at 1e-9 over a short horizon and in
distribution over a long one."""
import os

import numpy as np
import pytest

from common import ROOT, SMALL, World, LtransLib, make_params, setup, run, compare, assert_parity

pytestmark = pytest.mark.gpu
PASSIVE = dict(HTurbOn=0, VTurbOn=0, Behavior=0, settlementon=0, mortality=0)


def _pair(n, nexternal, nint=None, world_kw=SMALL, dob_max=0.0, locate=True, dtype=np.float32, sigerr=False, **kw):
    from oracle.oracle import Oracle
    w = World(**world_kw)
    prm = make_params(w, n, **kw)
    hab = w.habitat() if prm.settlementon else None
    g, o = LtransLib(), Oracle()
    setup(g, w, prm, n, habitat=hab, dob_max=dob_max, locate=locate, dtype=dtype)
    setup(o, w, prm, n, habitat=hab, dob_max=dob_max, locate=True, dtype=dtype)
    o.set_threads(os.cpu_count() or 1)
    rg, ro = run(g, w, nexternal, nint=nint, dtype=dtype), run(o, w, nexternal, nint=nint, dtype=dtype)
    fg, fo = g.fetch(), o.fetch()
    res = compare(fg, fo, w)
    ev = (g.drain_events(), o.drain_events())
    st = (g.stats(), o.stats())
    sig = (g.fetch_sigerr(), o.fetch_sigerr()) if sigerr else ()
    g.destroy(); o.destroy()
    return (rg, ro, res, ev, st, fg, fo) + sig


CASES = {
    "passive": dict(PASSIVE),
    "passive_late_release": dict(PASSIVE, _dob=7200.0),
    "hturb": dict(PASSIVE, HTurbOn=1),
    "salttemp": dict(PASSIVE, SaltTempOn=1),
    "closed_basin_errorflag2": dict(PASSIVE, HTurbOn=1, ConstantHTurb=40.0, ErrorFlag=2, OpenOceanBoundary=0),
    "errorflag3": dict(PASSIVE, HTurbOn=1, ConstantHTurb=40.0, ErrorFlag=3),
    "vtransform2": dict(PASSIVE, Vtransform=2),
    "vtransform3": dict(PASSIVE, Vtransform=3),
    "freeslip": dict(PASSIVE, FreeSlip=1),
    "behav1": dict(PASSIVE, Behavior=1), "behav2": dict(PASSIVE, Behavior=2), "behav3": dict(PASSIVE, Behavior=3),
    "behav4": dict(PASSIVE, Behavior=4), "behav5": dict(PASSIVE, Behavior=5), "behav6": dict(PASSIVE, Behavior=6),
    "behav7": dict(PASSIVE, Behavior=7),
    "oyster_full": dict(Behavior=4, HTurbOn=1, VTurbOn=0, pediage=3600.0, deadage=9000.0),
    "ariakensis_full": dict(Behavior=5, HTurbOn=1, VTurbOn=0, pediage=5400.0, deadage=10000.0, Sgradient=0.05),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_trajectory_parity_turbulence_off(name):
    kw = dict(CASES[name]); dob = kw.pop("_dob", 0.0)
    rg, ro, res, ev, st, fg, fo = _pair(500, 4, dob_max=dob, **kw)
    assert rg == ro
    assert_parity(res, 1e-9)
    assert ev[0] == ev[1]
    assert np.array_equal(st[0], st[1])
    if name in ("salttemp",):
        assert np.allclose(fg["salt"], fo["salt"], rtol=1e-12, atol=0) and np.allclose(fg["temp"], fo["temp"], rtol=1e-12, atol=0)
    if name == "oyster_full":
        assert st[0][0] > 0 and st[0][1] > 0                 # some settled, some died
        assert np.array_equal(fg["lifespan"], fo["lifespan"])


def test_f64_field_storage_and_device_side_locate():
    rg, ro, res, ev, st, fg, fo = _pair(300, 3, locate=False, dtype=np.float64, field_dtype=8, **PASSIVE)
    assert_parity(res, 1e-9)


def test_errorflag0_stops_on_lowest_particle():
    from oracle.oracle import Oracle
    kw = dict(PASSIVE, HTurbOn=1, ConstantHTurb=2000.0, ErrorFlag=0)     # huge kicks: particles jump elements
    w = World(**SMALL); n = 300
    prm = make_params(w, n, **kw)
    g, o = LtransLib(), Oracle()
    setup(g, w, prm, n); setup(o, w, prm, n)
    bad_g = bad_o = 0
    for it in range(1, 31):
        g.step(1, it); rc_o = o.step(1, it)
        rg, bad_g = g.sync(); ro, bad_o = o.sync()
        assert rg == ro
        if ro:
            break
    assert bad_o > 0 and bad_g == bad_o


def test_vturb_same_stream_short_horizon():
    """Same Philox stream, 4 internal steps of HTurb + VTurb.  The only branch on which CUDA and
    oracle may part is the reference's SigErr fall-back (spline -> linint), whose verdict hangs on
    the last bits of T and exp(): particles that have not met it on either
    side must agree to 1e-9, and they are nearly all of them."""
    rg, ro, res, ev, st, fg, fo, sg, so = _pair(2000, 1, nint=4, sigerr=True, **dict(PASSIVE, HTurbOn=1, VTurbOn=1))
    H = 30.0
    dz = np.abs(fg["z"] - fo["z"]) / H
    clean = (sg == 0) & (so == 0)
    assert clean.mean() >= 0.99 and dz[clean].max() <= 1e-9, (float(clean.mean()), float(dz[clean].max()))
    assert np.array_equal(fg["r_ele"][clean], fo["r_ele"][clean])
    assert np.mean(dz <= 1e-9) >= 0.995, float(np.mean(dz <= 1e-9))
    assert np.median(dz) <= 1e-14
    assert dz.max() <= 1e-3
    for k in ("n_r_ele", "n_u_ele", "n_v_ele", "n_status"):
        assert res[k] == 0, res
    assert res["eage"] == 0.0


def test_vturb_long_horizon_distribution():
    """60 internal steps of HTurb + VTurb (full-size grid): particles without a SigErr fall-back on
    either side still agree to 1e-9 after 3,600 random displacements each; the others differ by
    the fall-back and its chaotic amplification, so the depth distributions must coincide."""
    from scipy import stats as ss
    rg, ro, res, ev, st, fg, fo, sg, so = _pair(4000, 2, sigerr=True, world_kw={}, **dict(PASSIVE, HTurbOn=1, VTurbOn=1))
    H = 30.0
    dz = np.abs(fg["z"] - fo["z"]) / H
    clean = (sg == 0) & (so == 0)
    assert clean.mean() >= 0.9 and dz[clean].max() <= 1e-9, (float(clean.mean()), float(dz[clean].max()))
    assert np.median(dz) <= 1e-12
    assert np.mean(dz <= 1e-9) >= 0.9
    assert np.mean(fg["r_ele"] == fo["r_ele"]) >= 0.995
    assert ss.ks_2samp(fg["z"], fo["z"]).pvalue > 0.2
    assert abs(np.mean(fg["z"]) - np.mean(fo["z"])) <= 1e-3 * H
    assert np.array_equal(st[0][:3], st[1][:3]) or np.abs(st[0][:3] - st[1][:3]).max() <= 2
    # fall-back rates: both sides examine every interval of the VTurb fit
    assert 0.6 <= sg.sum() / max(1, so.sum()) <= 1.5, (int(sg.sum()), int(so.sum()))


def test_vturb_sigerr_rate_matches_reference_and_window_option():
    """Default: SIGS examines every interval of the 4 ws-knot VTurb fit like the reference
    (ver_turb:278-279), so the fall-back RATE equals the oracle's.  ltgpu_params.vturb_window_sigs = 1
    (opt-in approximation) examines only the 32 knots around the particle and takes the branch less often."""
    n = 6000
    out = {}
    for window in (0, 1):
        rg, ro, res, ev, st, fg, fo, sg, so = _pair(n, 1, sigerr=True, world_kw={}, **dict(PASSIVE, HTurbOn=1, VTurbOn=1, vturb_window_sigs=window))
        dz = np.abs(fg["z"] - fo["z"]) / 30.0
        clean = (sg == 0) & (so == 0)
        assert dz[clean].max() <= 1e-9 and clean.mean() > 0.95
        out[window] = (int(sg.sum()), int(so.sum()))
    assert out[0][1] == out[1][1]                               # the oracle ignores the switch
    assert 0.7 <= out[0][0] / out[0][1] <= 1.4, out             # ~110 events: +-10 % is one sigma
    assert out[1][0] < out[0][0], out


@pytest.mark.parametrize("name", ["passive", "hturb_salttemp", "oyster4_settle", "tidal7"])
def test_golden_fixtures(name):
    from golden.make_golden import run_case
    want = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    got = run_case(name, LtransLib)
    for k in want.files:
        if want[k].dtype.kind == "f":
            assert np.allclose(got[k], want[k], rtol=0, atol=1e-9 * max(1.0, float(np.abs(want[k]).max()))), (name, k)
        else:
            assert np.array_equal(got[k], want[k]), (name, k)


def _stepwise(n, nsteps, **kw):
    """CUDA and oracle side by side, one internal step at a time; returns per step the relative
    differences and the SigErr fall-back counts of both sides."""
    from oracle.oracle import Oracle
    w = World()
    prm = make_params(w, n, **kw)
    g, o = LtransLib(), Oracle()
    setup(g, w, prm, n); setup(o, w, prm, n); o.set_threads(os.cpu_count() or 1)
    L = float(max(np.ptp(w.x_r), np.ptp(w.y_r))); H = float(w.h.max())
    out = []
    for it in range(1, nsteps + 1):
        g.step(1, it); o.step(1, it)
        fg, fo = g.fetch(("x", "y", "z", "status", "r_ele")), o.fetch(("x", "y", "z", "status", "r_ele"))
        d = np.maximum(np.maximum(np.abs(fg["x"] - fo["x"]), np.abs(fg["y"] - fo["y"])) / L, np.abs(fg["z"] - fo["z"]) / H)
        out.append((d, g.fetch_sigerr(), o.fetch_sigerr(), fg, fo))
    return out


def test_config1_parity_until_first_sigerr():
    """BASELINE configs[0] grid (130x130x20), 5,000 passive particles, 30 internal steps, compared
    after EVERY step.  The reference replaces a water-column spline by linint when SIGS raises
    SigErr; that verdict is decided by the last bits of T and of exp(), so
    CUDA and oracle take the branch at the same RATE but not at the same steps.  Claim checked
    here: a particle agrees within 1e-9 for as long as neither side has taken the branch, every
    larger difference belongs to a particle that has, and the two rates agree."""
    hist = _stepwise(5000, 30, **PASSIVE)
    for it, (d, sg, so, fg, fo) in enumerate(hist, 1):
        clean = (sg == 0) & (so == 0)
        assert d[clean].max() <= 1e-9, (it, float(d[clean].max()))
        assert np.array_equal(fg["status"], fo["status"]) and np.array_equal(fg["r_ele"][clean], fo["r_ele"][clean])
    d, sg, so, fg, fo = hist[-1]
    assert clean.mean() > 0.97                                   # ~4e-4 fall-backs per particle-step and side
    assert d.max() <= 1e-5                                       # a fall-back moves a particle by ~1e-7 of the depth
    ng, no = int(sg.sum()), int(so.sum())
    assert 20 <= ng <= 120 and 20 <= no <= 120 and 0.5 <= ng / no <= 2.0, (ng, no)


def test_config1_two_days_passive():
    """BASELINE configs[0] at full size: 130x130x20 grid, 5,000 passive particles, RK4 advection
    only, 2 days = 48 external x 30 internal steps (7.2e6 particle-steps), CUDA vs oracle.  Over
    1,440 steps most particles meet a SigErr fall-back on one side or the other (see the test
    above); the ones that never do must agree to 1e-9, all of them to 1e-4 of the domain, and
    statuses / boundary events must be identical."""
    from oracle.oracle import Oracle
    w = World(); n = 5000
    prm = make_params(w, n, **PASSIVE)
    g, o = LtransLib(), Oracle()
    setup(g, w, prm, n); setup(o, w, prm, n); o.set_threads(os.cpu_count() or 1)
    rg, ro = run(g, w, 48), run(o, w, 48)
    fg, fo = g.fetch(), o.fetch()
    sg, so = g.fetch_sigerr(), o.fetch_sigerr()
    stg, sto = g.stats(), o.stats()
    assert rg == ro and g.drain_events() == o.drain_events()
    assert np.array_equal(stg[[0, 1, 2, 5, 6, 7]], sto[[0, 1, 2, 5, 6, 7]]), (stg, sto)      # settled, dead, out, errors, active, unborn
    assert np.all(np.abs(stg[3:5] - sto[3:5]) <= 0.002 * sto[3:5] + 2), (stg, sto)            # land / bottom hit totals
    assert np.array_equal(fg["status"], fo["status"]) and np.array_equal(fg["age"], fo["age"])
    L = float(max(np.ptp(w.x_r), np.ptp(w.y_r))); H = float(w.h.max())
    d = np.maximum(np.maximum(np.abs(fg["x"] - fo["x"]), np.abs(fg["y"] - fo["y"])) / L, np.abs(fg["z"] - fo["z"]) / H)
    clean = (sg == 0) & (so == 0)
    assert clean.sum() >= 0.2 * n and d[clean].max() <= 1e-9, (int(clean.sum()), float(d[clean].max()))
    assert np.array_equal(fg["r_ele"][clean], fo["r_ele"][clean])
    assert np.median(d) <= 1e-9 and np.quantile(d, 0.99) <= 1e-5, (float(np.median(d)), float(np.quantile(d, 0.99)))
    assert (fg["r_ele"] != fo["r_ele"]).mean() <= 0.002
    assert 0.6 <= sg.sum() / so.sum() <= 1.6, (int(sg.sum()), int(so.sum()))


def test_full_size_properties():
    """BASELINE configs[1] size (1M turbulent particles, 130x130x20): determinism, sharding
    independence (Philox keyed on global id) and physical invariants."""
    w = World(); n = 1_000_000
    prm = make_params(w, n, Behavior=0, settlementon=0, mortality=0, TrackCollisions=0)
    x, y, z, dob, r, u, v = w.seed_particles(n)

    def go(sl, first):
        g = LtransLib().create(prm)
        g.set_grid(w.grid()); g.set_bounds(w.bounds())
        g.set_particles(x[sl], y[sl], z[sl], dob[sl], None, r[sl], u[sl], v[sl], first_id=first)
        for k in range(3):
            g.push_hydro(w.record(k))
        for it in range(1, 7):
            g.step(1, it)
        assert g.sync() == (0, 0)
        f = g.fetch(); st = g.stats(); g.destroy()
        return f, st
    fa, sa = go(slice(0, n), 1)
    fb, sb = go(slice(0, n), 1)
    for k in ("x", "y", "z", "status", "r_ele"):
        assert np.array_equal(fa[k], fb[k]), k                    # bit-identical reruns
    h = n // 2
    f1, s1 = go(slice(0, h), 1); f2, s2 = go(slice(h, n), h + 1)
    for k in ("x", "y", "z", "status", "r_ele"):
        assert np.array_equal(np.concatenate([f1[k], f2[k]]), fa[k]), k
    assert np.array_equal(s1 + s2, sa)
    act = fa["status"] == 0
    assert act.mean() > 0.95 and np.isfinite(fa["x"]).all() and np.isfinite(fa["z"]).all()
    assert fa["z"][act].max() < 1.0 and fa["z"][act].min() > -30.5
    assert np.all(fa["age"][act] == 6 * prm.idt)
    moved = np.hypot(fa["x"] - x, fa["y"] - y)[act]
    assert 1.0 < np.median(moved) < 6 * prm.idt * 1.5


def test_driver_outputs_match(tmp_path):
    """Whole pipeline (run loop, print intervals, para*.csv / endfile.csv / hit files) with the
    CUDA engine vs the oracle engine: same files, same numbers."""
    from oracle.oracle import Oracle
    from test_formats import run_driver
    og, oo = str(tmp_path / "gpu"), str(tmp_path / "ora")
    run_driver(LtransLib(), og)
    run_driver(Oracle(), oo)
    assert sorted(os.listdir(og)) == sorted(os.listdir(oo))
    for name in sorted(os.listdir(og)):
        if name.startswith("para") or name == "endfile.csv":
            a = np.loadtxt(os.path.join(og, name), delimiter=","); b = np.loadtxt(os.path.join(oo, name), delimiter=",")
            assert a.shape == b.shape and np.allclose(a, b, rtol=0, atol=1.01e-3), name
            assert np.mean(a == b) > 0.999, name
        else:
            assert open(os.path.join(og, name)).read().splitlines()[0] == open(os.path.join(oo, name)).read().splitlines()[0]


def test_driver_from_netcdf_files(tmp_path):
    """CUDA engine fed from ROMS grid / history NetCDF files (tdim = 2 records per file) and writing
    the particle NetCDF file: same CSVs as the run fed from memory."""
    from scipy.io import netcdf_file
    from test_formats import run_driver
    from ltrans_b200.host import roms_io
    w = World(**SMALL)
    roms_io.write_grid_nc(str(tmp_path / "grid.nc"), w)
    roms_io.write_history_nc(w, str(tmp_path / "his_"), ".nc", 1, 4, nrec=5, tdim=2, startfile=True)
    rw = roms_io.RomsWorld(str(tmp_path / "grid.nc"), str(tmp_path / "his_"), ".nc", 1, 4, tdim=2, startfile=True)
    a, b = str(tmp_path / "a"), str(tmp_path / "b")
    run_driver(LtransLib(), a)
    run_driver(LtransLib(), b, world=rw, run_kw=dict(write_nc=True, NCOutFile="out"))
    rw.close()
    for name in sorted(os.listdir(a)):
        assert open(os.path.join(a, name)).read() == open(os.path.join(b, name)).read(), name
    with netcdf_file(os.path.join(b, "out.nc"), "r", mmap=False) as f:
        assert f.variables["lon"][:].shape == (7, 120)
        last = np.loadtxt(os.path.join(b, "para10000007.csv"), delimiter=",")
        assert np.allclose(f.variables["depth"][-1], last[:, 0], rtol=0, atol=5.01e-4)
        assert np.array_equal(f.variables["color"][-1], last[:, 1])


GULF = dict(ni=48, nj=40, us=36, hmin=40.0, hmax=600.0, dlon=0.02, dlat=0.018, speed=0.9)


def test_gulf_like_buoyant_particles_turbulence_off():
    """BASELINE configs[3] shape at test size: us 36 / ws 37, deep water, Behavior 6 with a positive
    `sink` (buoyant droplets), open ocean boundary."""
    rg, ro, res, ev, st, fg, fo = _pair(600, 3, world_kw=GULF, **dict(PASSIVE, Behavior=6, sink=0.002, swimstart=0.0,
                                                                       HTurbOn=1, OpenOceanBoundary=1))
    assert rg == ro
    assert_parity(res, 1e-9)
    assert np.array_equal(st[0], st[1])


def test_gulf_like_vturb_windows():
    """ws = 37 -> 148 spline knots: the 32-knot window of k_vturb is re-centred by many particles."""
    rg, ro, res, ev, st, fg, fo = _pair(1500, 1, nint=3, world_kw=GULF, **dict(PASSIVE, Behavior=6, sink=0.002, HTurbOn=1, VTurbOn=1))
    dz = np.abs(fg["z"] - fo["z"]) / 600.0
    assert np.mean(dz <= 1e-9) >= 0.99 and np.median(dz) <= 1e-14, (float(np.mean(dz <= 1e-9)), float(np.median(dz)))
    assert res["n_status"] == 0 and res["n_r_ele"] == 0


def test_oyster_run_with_turbulence_statistics():
    """BASELINE configs[2] shape at test size: Behavior 4 (C. virginica), HTurb + VTurb, settlement
    polygons with holes, mortality.  Same Philox stream; compared in distribution."""
    from scipy import stats as ss
    rg, ro, res, ev, st, fg, fo = _pair(3000, 3, **dict(Behavior=4, HTurbOn=1, VTurbOn=1, pediage=3600.0, deadage=9000.0))
    assert rg == ro
    assert np.abs(st[0][:3] - st[1][:3]).max() <= max(3, 0.01 * 3000), (st[0], st[1])      # settled, dead, out of bounds
    assert np.mean(fg["status"] == fo["status"]) >= 0.99
    assert ss.ks_2samp(fg["z"], fo["z"]).pvalue > 0.1
    assert np.mean(np.abs(fg["x"] - fo["x"]) <= 1e-6 * 1.4e4) >= 0.95


def test_lonlat_on_device():
    from oracle.oracle import Oracle
    w = World(**SMALL); n = 3000
    prm = make_params(w, n, **PASSIVE)
    g, o = LtransLib(), Oracle()
    setup(g, w, prm, n); setup(o, w, prm, n)
    run(g, w, 1); run(o, w, 1)
    for sph in (True, False):
        a, b = g.fetch_lonlat(w.proj, spherical=sph), o.fetch_lonlat(w.proj, spherical=sph)
        assert np.allclose(a[0], b[0], rtol=0, atol=1e-11) and np.allclose(a[1], b[1], rtol=0, atol=1e-11)
    f = g.fetch(("x", "y"))
    assert np.allclose(a[0], f["x"] / w.proj.R * 180.0 / w.proj.pi, rtol=1e-13)


def test_abi_misc_calls():
    """reset_hits, stats, device_ptr (particle-order copies), kernel_times, launch_count."""
    import ctypes as C
    import torch
    w = World(**SMALL); n = 300
    prm = make_params(w, n, **dict(PASSIVE, HTurbOn=1, ConstantHTurb=30.0))
    g = LtransLib()
    setup(g, w, prm, n)
    l0 = g.launch_count()
    g.kernel_times(True)
    g.run_external(1)
    ms, steps = g.kernel_times(False)
    assert steps == 30 and ms[0] > 0 and ms[2] > 0 and ms[1] < 0.2 * ms[0]     # VTurb off: k_vturb not launched
    assert g.launch_count() - l0 >= 60
    f = g.fetch()
    st = g.stats()
    assert st[3] == f["hitLand"].sum() > 0 and st[6] + st[2] == n
    px = g.device_ptr(0)                               # x in particle order, device memory
    buf = torch.empty(n, dtype=torch.float64, device="cuda")
    rt = C.CDLL("libcudart.so.12")
    assert rt.cudaMemcpy(C.c_void_p(buf.data_ptr()), C.c_void_p(px), C.c_size_t(8 * n), C.c_int(3)) == 0
    assert np.array_equal(buf.cpu().numpy(), f["x"])
    g.reset_hits()
    assert g.fetch(("hitLand",))["hitLand"].sum() == 0
    g.destroy()


def test_odd_grid_f64_push_with_salt_temp():
    """every field of a pushed record is padded to 16 bytes in the staging buffers; an
    odd x odd grid with float64 records and salt / temp is the case whose padding used to run past
    the allocation."""
    odd = dict(ni=41, nj=37, us=10)
    rg, ro, res, ev, st, fg, fo = _pair(300, 3, world_kw=odd, dtype=np.float64, field_dtype=8, **dict(PASSIVE, SaltTempOn=1))
    assert rg == ro
    assert_parity(res, 1e-9)
    assert np.allclose(fg["salt"], fo["salt"], rtol=1e-12, atol=0)


def test_event_log_recycles_and_reports_overflow(monkeypatch):
    """The device event log is emptied at every synchronisation point, so a host that synchronises
    once per internal step loses nothing however long the run; when it is too small for the steps
    queued between two synchronisations the loss is counted and reported, not silent."""
    from oracle.oracle import Oracle
    from ltrans_b200.host.binding import LTGPU_W_EVENTS_LOST
    kw = dict(PASSIVE, HTurbOn=1, ConstantHTurb=2000.0, ErrorFlag=1)          # huge kicks: many particles raise events
    w = World(**SMALL); n = 300
    prm = make_params(w, n, **kw)
    monkeypatch.setenv("LTGPU_EVCAP", str(n))                                 # one step's worth
    g, o = LtransLib(), Oracle()
    setup(g, w, prm, n); setup(o, w, prm, n)
    for it in range(1, 31):
        g.step(1, it); o.step(1, it)
        assert g.sync() == o.sync()
    eg, eo = g.drain_events(64, everything=True), o.drain_events(1 << 16)
    assert len(eo) > n and eg == eo and g.lost_events() == 0                  # more events than the log holds, none lost
    assert g.stats()[5] == 0
    g.destroy(); o.destroy()
    monkeypatch.setenv("LTGPU_EVCAP", "8")
    g = LtransLib()
    setup(g, w, prm, n)
    g.run_external(1)                                                         # 30 steps queued without a synchronisation
    g.sync()
    got = g.drain_events(1 << 16)
    assert g.events_lost > 0 and len(got) + g.events_lost == len(eo)
    assert set(got) <= set(eo)
    import ctypes as C
    from ltrans_b200.host.binding import Event
    buf = (Event * 4)(); k = C.c_int32(0)
    assert g.lib.ltgpu_drain_events(g.ctx, buf, C.c_int32(4), C.byref(k)) == 0   # the warning is raised once
    g.destroy()


def test_unlocated_particle_does_not_fault():
    """a particle released outside every element keeps element id 0; with ErrorFlag = 2
    and mortality off it is still stepped.  It must raise 'not in rho element', not index the
    adjacency table out of bounds."""
    w = World(**SMALL); n = 64
    prm = make_params(w, n, **dict(PASSIVE, ErrorFlag=2))
    g = LtransLib().create(prm)
    g.set_grid(w.grid()); g.set_bounds(w.bounds())
    x, y, z, dob, r, u, v = w.seed_particles(n)
    x[:4] = w.x_r.min() - 5e4                                                 # far outside the grid
    g.set_particles(x, y, z, dob, None, None, None, None)
    rc, counts, bad = g.screen_initial()
    assert rc == 0 and counts[0] == 4
    for k in range(3):
        g.push_hydro(w.record(k))
    g.run_external(1)
    assert g.sync() == (0, 0)
    ev = g.drain_events(1 << 12)
    codes = {c for (p_, c, t) in ev if p_ <= 4}
    assert codes == {11, 23}          # start-up screen, then setEle's error (the v-grid test is the last to set it, hydro:1497)
    f = g.fetch(("x", "status"))
    assert np.all(f["status"][:4] == -1) and np.isfinite(f["x"]).all()
    g.destroy()


def _vturb_run(monkeypatch, n, env, world_kw=SMALL, nint=4, **kw):
    for k in ("LTGPU_VTURB_LEGACY", "LTGPU_VTURB_CHUNK"):
        monkeypatch.delenv(k, raising=False)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    w = World(**world_kw)
    prm = make_params(w, n, **dict(PASSIVE, HTurbOn=1, VTurbOn=1, **kw))
    g = LtransLib()
    setup(g, w, prm, n)
    run(g, w, 1, nint=nint)
    f, s = g.fetch(), g.fetch_sigerr()
    g.destroy()
    return f, s


def test_vturb_fit_walk_kernels_vs_round1_fused_kernel(monkeypatch):
    """k_vbuild (one warp per column) + k_vwalk against the round-1 per-thread kernel on the same
    particles: the same fit up to the rounding of the moving average (direct 8-term sums, like the
    reference, instead of a running sum), the same walk."""
    for wk, H in ((SMALL, 30.0), (GULF, 600.0)):
        fa, sa = _vturb_run(monkeypatch, 1500, {}, world_kw=wk)
        fb, sb = _vturb_run(monkeypatch, 1500, {"LTGPU_VTURB_LEGACY": "1"}, world_kw=wk)
        clean = (sa == 0) & (sb == 0)
        dz = np.abs(fa["z"] - fb["z"]) / H
        assert clean.mean() > 0.98 and dz[clean].max() <= 1e-10, (float(clean.mean()), float(dz[clean].max()))
        assert np.abs(fa["x"] - fb["x"])[clean].max() <= 1e-9 * 1e5 and np.array_equal(fa["status"], fb["status"])


def test_vturb_chunked_scratch_is_invisible(monkeypatch):
    """the fit -> walk scratch is sized for a chunk of particles; many small chunks = one big one, bit for bit"""
    fa, sa = _vturb_run(monkeypatch, 1500, {})
    fb, sb = _vturb_run(monkeypatch, 1500, {"LTGPU_VTURB_CHUNK": "96"})
    for k in ("x", "y", "z", "status", "r_ele"):
        assert np.array_equal(fa[k], fb[k]), k
    assert np.array_equal(sa, sb)


def test_slot_order_and_vturb_visiting_order_are_invisible(monkeypatch):
    """At Gulf scale the slots are ordered by (8 depth bins, element) and k_vwalk visits them through a second index
    ordered by (64 bins, element), k_vbuild in slot order when one scratch chunk holds all particles, else in the
    walk's order.  Particles do not interact and Philox is keyed by particle id: every combination gives the same
    bits as one order for all and as no re-sort at all."""
    for k in ("LTGPU_SORT", "LTGPU_SORT_MODE", "LTGPU_VT_BINS", "LTGPU_VB_SLOT_ORDER"):
        monkeypatch.delenv(k, raising=False)
    for wk in (SMALL, GULF):
        ref, sref = _vturb_run(monkeypatch, 3000, {"LTGPU_SORT": "0"}, world_kw=wk)
        monkeypatch.delenv("LTGPU_SORT", raising=False)
        for env in ({"LTGPU_SORT_MODE": "8", "LTGPU_VT_BINS": "64"},
                    {"LTGPU_SORT_MODE": "8", "LTGPU_VT_BINS": "64", "LTGPU_VB_SLOT_ORDER": "0"},
                    {"LTGPU_SORT_MODE": "8", "LTGPU_VT_BINS": "64", "LTGPU_VTURB_CHUNK": "416"},
                    {"LTGPU_SORT_MODE": "0", "LTGPU_VT_BINS": "127"},
                    {"LTGPU_SORT_MODE": "32", "LTGPU_VT_BINS": "0"}):
            f, sg = _vturb_run(monkeypatch, 3000, env, world_kw=wk)
            for k in ("x", "y", "z", "status", "r_ele", "age"):
                assert np.array_equal(ref[k], f[k]), (env, k)
            assert np.array_equal(sref, sg), env
            for k in env:
                monkeypatch.delenv(k, raising=False)


def _first_divergence(w, prm, n, nexternal, habitat=None, tol=1e-9):
    """CUDA and oracle side by side for nexternal external steps, compared after EVERY internal step.
    Returns (first step with a deviation > tol per particle or -1, whether the two SigErr fall-back
    counters had moved differently by then, fall-back totals, final fetches)."""
    from oracle.oracle import Oracle
    g, o = LtransLib(), Oracle()
    setup(g, w, prm, n, habitat=habitat); setup(o, w, prm, n, habitat=habitat); o.set_threads(os.cpu_count() or 1)
    L = float(max(np.ptp(w.x_r), np.ptp(w.y_r))); H = float(w.h.max())
    first = np.full(n, -1); explained = np.zeros(n, bool); moved_diff = np.zeros(n, bool)
    sg0 = np.zeros(n, np.int32); so0 = np.zeros(n, np.int32)
    stepIT = prm.dt // prm.idt
    k = 0
    for p in range(1, nexternal + 1):
        if p > 2:
            rec = w.record(p)
            g.push_hydro(rec); g.rotate_hydro(); o.push_hydro(rec); o.rotate_hydro()
        for it in range(1, stepIT + 1):
            k += 1
            g.step(p, it); o.step(p, it)
            fg, fo = g.fetch(("x", "y", "z", "status", "r_ele")), o.fetch(("x", "y", "z", "status", "r_ele"))
            sg, so = g.fetch_sigerr(), o.fetch_sigerr()
            moved_diff |= (sg - sg0) != (so - so0)
            fb_now = ((sg - sg0) > 0) | ((so - so0) > 0)              # a fall-back on either side in THIS step
            sg0, so0 = sg, so
            d = np.maximum(np.maximum(np.abs(fg["x"] - fo["x"]), np.abs(fg["y"] - fo["y"])) / L, np.abs(fg["z"] - fo["z"]) / H)
            new = (first < 0) & (d > tol)
            first[new] = k
            explained[new] = moved_diff[new] | fb_now[new]
            same_so_far = ~moved_diff
            assert np.array_equal(fg["status"][same_so_far], fo["status"][same_so_far]), k
            assert np.array_equal(fg["r_ele"][same_so_far & (first < 0)], fo["r_ele"][same_so_far & (first < 0)]), k
    out = (first, explained, moved_diff, int(sg0.sum()), int(so0.sum()), g.fetch(), o.fetch(), (g.drain_events(1 << 16), o.drain_events(1 << 16)))
    g.destroy(); o.destroy()
    return out


def test_config1_first_divergence_over_all_1440_steps():
    """BASELINE configs[0] in full (130x130x20, 5,000 passive particles, RK4 only, 2 days = 1,440 internal
    steps), CUDA vs oracle after EVERY step.  North-star gate: trajectories within 1e-9 with identical cell
    ids and events.  The one branch on which two correct implementations part is the reference's SigErr
    fall-back, so the claim checked is: a particle's FIRST deviation above 1e-9 appears
    either in a step in which one of the two sides took a fall-back (both may, on different splines of that
    step: seen 7 times in 7.2e6 particle-steps, each with a deviation of exactly 0 the step before), or after
    the two fall-back counters have moved differently; everything else agrees to 1e-9 for the whole run."""
    w = World(); n = 5000
    prm = make_params(w, n, **PASSIVE)
    first, explained, moved_diff, ng, no, fg, fo, ev = _first_divergence(w, prm, n, 48)
    deviated = first >= 0
    unexplained = int((deviated & ~explained).sum())
    print("config 1, 1440 steps: never deviated %.3f, fall-back histories identical %.3f, fall-backs %d (CUDA) / %d (oracle), "
          "unexplained first deviations %d" % (1 - deviated.mean(), 1 - moved_diff.mean(), ng, no, unexplained))
    assert unexplained == 0, unexplained
    same = ~moved_diff                                            # identical fall-back history => 1e-9 for the whole run
    L = float(max(np.ptp(w.x_r), np.ptp(w.y_r))); H = float(w.h.max())
    d = np.maximum(np.maximum(np.abs(fg["x"] - fo["x"]), np.abs(fg["y"] - fo["y"])) / L, np.abs(fg["z"] - fo["z"]) / H)
    assert d[same & ~deviated].max() <= 1e-9
    assert np.array_equal(fg["status"], fo["status"]) and ev[0] == ev[1]
    assert 0.7 <= ng / no <= 1.4, (ng, no)


CHES = dict(ni=120, nj=80, us=20, dlon=0.02, dlat=0.018)
GULF_MID = dict(ni=256, nj=192, us=36, hmin=50.0, hmax=3000.0, dlon=0.08, dlat=0.072, speed=0.9)


@pytest.mark.parametrize("name", ["config3_oysters", "config4_gulf"])
def test_configs_3_and_4_worlds_turbulence_off_first_divergence(name):
    """The worlds and switches of BASELINE configs[2] (Chesapeake-scale grid, Behavior 4, settlement polygons
    with holes, mortality) and configs[3] (Gulf-scale levels and depths, buoyant Behavior 6, open boundary)
    with turbulence off, 2,000 particles, 3 external steps, compared after every internal step."""
    if name == "config3_oysters":
        w = World(**CHES)
        prm = make_params(w, 2000, HTurbOn=0, VTurbOn=0, Behavior=4, settlementon=1, holesExist=1, mortality=1, ErrorFlag=3,
                          pediage=3600.0, deadage=2.5 * 3600.0)
        hab = w.habitat(npoly=64)
    else:
        w = World(**GULF_MID)
        prm = make_params(w, 2000, HTurbOn=0, VTurbOn=0, Behavior=6, sink=0.002, settlementon=0, mortality=0, ErrorFlag=3, OpenOceanBoundary=1)
        hab = None
    first, explained, moved_diff, ng, no, fg, fo, ev = _first_divergence(w, prm, 2000, 3, habitat=hab)
    deviated = first >= 0
    assert int((deviated & ~explained).sum()) == 0
    assert (1 - deviated.mean()) > 0.9
    same = ~moved_diff
    for k in ("status", "endpoly", "hitBottom", "hitLand"):
        assert np.array_equal(fg[k][same], fo[k][same]), k
    assert np.array_equal(fg["lifespan"][same], fo["lifespan"][same])
    if name == "config3_oysters":
        assert (fg["status"] == -2).sum() > 0 and (fg["status"] == -1).sum() > 0        # some settled, some died


def test_fp32_walk_option_statistical_parity(monkeypatch):
    """ltgpu_params.vturb_fp32_walk = 1 (opt-in, never the headline): the 60 random-displacement sub-steps
    in single precision.  North star, turbulent runs: dispersion statistics within a stated tolerance.
    Same Philox stream as the FP64 walk, so the two are compared particle by particle AND in distribution:
      * median |dz| / H <= 1e-5, 99 % of the particles within 1e-2 of the depth (differences amplify where K is
        small and K' large), identical cells and status for >= 99.5 %;
      * depth distributions: two-sample KS p > 0.2, means within 1e-3 H, standard deviations within 1 %;
      * the well-mixed condition: particles released uniformly over the column stay uniformly distributed
        (chi-square over 10 depth classes of the relative depth, FP32 against FP64)."""
    from scipy import stats as ss
    n = 20000
    f64, s64 = _vturb_run(monkeypatch, n, {}, world_kw={}, nint=30)
    f32, s32 = _vturb_run(monkeypatch, n, {}, world_kw={}, nint=30, vturb_fp32_walk=1)
    H = 30.0
    dz = np.abs(f32["z"] - f64["z"]) / H
    assert np.median(dz) <= 1e-5 and np.quantile(dz, 0.99) <= 1e-2, (float(np.median(dz)), float(np.quantile(dz, 0.99)))
    assert np.mean(f32["r_ele"] == f64["r_ele"]) >= 0.995 and np.mean(f32["status"] == f64["status"]) >= 0.999
    assert ss.ks_2samp(f32["z"], f64["z"]).pvalue > 0.2
    assert abs(f32["z"].mean() - f64["z"].mean()) <= 1e-3 * H
    assert abs(f32["z"].std() / f64["z"].std() - 1.0) <= 0.01
    w = World()
    x, y, z0, dob, r, u, v = w.seed_particles(n)
    i0 = np.clip(np.searchsorted(w.x_r[0], x) - 1, 0, w.ni - 1); j0 = np.clip(np.searchsorted(w.y_r[:, 0], y) - 1, 0, w.nj - 1)
    h = w.h[j0, i0]
    act = f64["status"] == 0
    rel32, rel64 = -f32["z"][act] / h[act], -f64["z"][act] / h[act]
    c32, _ = np.histogram(rel32, bins=10, range=(0, 1)); c64, _ = np.histogram(rel64, bins=10, range=(0, 1))
    chi2 = float(np.sum((c32 - c64) ** 2 / np.maximum(c32 + c64, 1)))
    assert chi2 < 27.9, chi2                                       # chi-square, 10 d.o.f. (two-sample form), p = 0.002
    assert abs(int(s32.sum()) - int(s64.sum())) <= 4.0 * np.sqrt(float(s32.sum() + s64.sum())) + 5      # same fit: SigErr fall-backs at the same rate (Poisson counts)
