"""Start-up screen of ini_LTRANS (LTRANS.f90:356-452) and the device-side particle location
(SURVEY.md section 8f row 3): particles released on land, in an island or off the grid."""
import numpy as np
import pytest

from common import SMALL, World, LtransLib, make_params
from ltrans_b200.host import formats


def _release(w, n, seed=5):
    """n good particles, then 4 bad ones: far outside, inside the block island, in the western
    land rim (outside the main polygon), in the diagonal-pair island."""
    x, y, z, dob, r, u, v = w.seed_particles(n, seed=seed)
    a, b = int(0.30 * w.nj), int(0.38 * w.ni)
    bad = [(w.x_r[0, 0] - 5e4, w.y_r[0, 0] - 5e4), (w.x_r[a + 1, b + 2] + 1.0, w.y_r[a + 1, b + 2] + 1.0),
           (0.5 * (w.x_r[w.nj // 2, 0] + w.x_r[w.nj // 2, 1]), w.y_r[w.nj // 2, 0] + 3.0)]
    a, b = int(0.62 * w.nj), int(0.55 * w.ni)
    bad.append((w.x_r[a + 1, b + 1] + 2.0, w.y_r[a + 1, b + 1] + 2.0))
    bx, by = np.array(bad).T
    return (np.concatenate([x, bx]), np.concatenate([y, by]), np.concatenate([z, np.full(4, -1.0)]),
            np.concatenate([dob, np.zeros(4)]))


def _screen(lib, w, ErrorFlag, n=200, OpenOceanBoundary=1, mortality=1):
    x, y, z, dob = _release(w, n)
    prm = make_params(w, len(x), Behavior=0, settlementon=0, ErrorFlag=ErrorFlag, OpenOceanBoundary=OpenOceanBoundary,
                      mortality=mortality)
    lib.create(prm); lib.set_grid(w.grid()); lib.set_bounds(w.bounds())
    lib.set_particles(x, y, z, dob, None, None, None, None)
    return lib.screen_initial(), lib


def _check(res, lib, n, ErrorFlag):
    rc, counts, bad = res
    assert rc == 0 and bad == 0
    assert list(counts) == [2, 2, 0, 0, 0], counts          # two outside the main polygon, two inside islands
    ev = sorted(lib.drain_events(64))
    assert ev == [(n + 1, 11, 0.0), (n + 2, 12, 0.0), (n + 3, 11, 0.0), (n + 4, 12, 0.0)]
    st = lib.fetch(("status",))["status"]
    assert np.all(st[:n] == 0) and np.all(st[n:] == (-1 if ErrorFlag == 2 else -3))
    return ev


@pytest.mark.parametrize("ErrorFlag", [2, 3])
def test_screen_oracle(tmp_path, ErrorFlag):
    from oracle.oracle import Oracle
    w = World(**SMALL)
    res, lib = _screen(Oracle(), w, ErrorFlag)
    ev = _check(res, lib, 200, ErrorFlag)
    log = str(tmp_path / "ErrorLog.txt")
    formats.append_errorlog(log, ev)
    lines = open(log).read().splitlines()
    assert lines[0] == "Particle        201 initially outside main bounds" and lines[1].endswith("initially inside island bounds")


def test_screen_stop_oracle():
    from oracle.oracle import Oracle
    rc, counts, bad = _screen(Oracle(), World(**SMALL), 0)[0]
    assert rc != 0 and bad == 201                           # the reference STOPs at the first bad particle


@pytest.mark.gpu
@pytest.mark.parametrize("ErrorFlag", [1, 2])
def test_screen_gpu_matches_oracle(ErrorFlag):
    from oracle.oracle import Oracle
    w = World(**SMALL)
    rg, g = _screen(LtransLib(), w, ErrorFlag)
    ro, o = _screen(Oracle(), w, ErrorFlag)
    assert np.array_equal(rg[1], ro[1])
    _check(rg, g, 200, ErrorFlag)
    fg, fo = g.fetch(("r_ele", "u_ele", "v_ele", "status")), o.fetch(("r_ele", "u_ele", "v_ele", "status"))
    for k in fg:
        assert np.array_equal(fg[k], fo[k]), k
    rc, counts, bad = _screen(LtransLib(), w, 0)[0]
    assert rc != 0 and bad == 201


@pytest.mark.gpu
def test_locate_buckets_equal_full_scan():
    """Device location through the element buckets == the oracle's whole-grid scan (hydro:1436-1457),
    including points on element edges / nodes and points off the grid (0)."""
    from oracle.oracle import Oracle
    w = World(ni=40, nj=36, us=6)
    rng = np.random.default_rng(3)
    x, y, z, dob, r, u, v = w.seed_particles(4000, seed=9)
    jj, ii = rng.integers(2, w.nj - 2, 300), rng.integers(2, w.ni - 3, 300)
    xe = np.concatenate([w.x_r[jj, ii], 0.5 * (w.x_r[jj, ii] + w.x_r[jj, ii + 1]), w.x_u[jj, ii], w.x_v[jj, ii]])
    ye = np.concatenate([w.y_r[jj, ii], 0.5 * (w.y_r[jj, ii] + w.y_r[jj, ii + 1]), w.y_u[jj, ii], w.y_v[jj, ii]])
    xo = w.x_r.min() + (w.x_r.max() - w.x_r.min()) * rng.uniform(-0.3, 1.3, 500)
    yo = w.y_r.min() + (w.y_r.max() - w.y_r.min()) * rng.uniform(-0.3, 1.3, 500)
    X, Y = np.concatenate([x, xe, xo]), np.concatenate([y, ye, yo])
    out = []
    for lib in (LtransLib(), Oracle()):
        prm = make_params(w, len(X), Behavior=0, settlementon=0)
        lib.create(prm); lib.set_grid(w.grid()); lib.set_bounds(w.bounds())
        lib.set_particles(X, Y, np.full(len(X), -1.0), np.zeros(len(X)), None, None, None, None)
        out.append(lib.fetch(("r_ele", "u_ele", "v_ele")))
    for k in out[0]:
        assert np.array_equal(out[0][k], out[1][k]), k
    assert np.array_equal(out[0]["r_ele"][:4000], r) and (out[0]["r_ele"][-500:] == 0).any()
