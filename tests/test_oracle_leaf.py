"""Pins the CPU oracle: published known answers (MT19937, Philox4x32-10) and analytic
invariants of the leaf numerics (SURVEY.md 8c mitigations 2 and 3).  The reference has
no tests or golden vectors for this path; see oracle/ltrans_oracle.h."""
import ctypes as C

import numpy as np
import pytest

from oracle.oracle import leaf, dptr

L = leaf()


def _d(*v):
    return (C.c_double * len(v))(*v)


def test_mt19937_published_vectors():
    # mt19937ar.out (Matsumoto & Nishimura); random_module.f90:108-246
    L.ora_mt_init_genrand(5489)
    assert L.ora_mt_int32() == 3499211612
    key = (C.c_uint32 * 4)(0x123, 0x234, 0x345, 0x456)
    L.ora_mt_init_by_array(key, 4)
    assert [L.ora_mt_int32() for _ in range(5)] == [1067595299, 955945823, 477289528, 4107218783, 4228976476]
    L.ora_mt_init_genrand(9)
    r1 = L.ora_mt_real1(); r3 = L.ora_mt_real3()
    assert 0.0 <= r1 <= 1.0 and 0.0 < r3 < 1.0


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    def ph(ctr, key):
        c = (C.c_uint32 * 4)(*ctr); k = (C.c_uint32 * 2)(*key); o = (C.c_uint32 * 4)()
        L.ora_philox4x32_10(c, k, o)
        return [int(x) for x in o]
    assert ph([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert ph([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert ph([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_polintd_exact_for_quadratics():
    rng = np.random.default_rng(0)
    for _ in range(200):
        a, b, c = rng.normal(size=3)
        xa = np.array([0.0, 3600.0, 7200.0]) + rng.integers(0, 5) * 3600.0
        ya = a + b * xa * 1e-3 + c * (xa * 1e-3) ** 2
        x = xa[0] + rng.random() * 7200.0
        want = a + b * x * 1e-3 + c * (x * 1e-3) ** 2
        got = L.ora_polintd(dptr(xa), dptr(ya), x)
        assert abs(got - want) <= 1e-10 * max(1.0, abs(want))


def test_linint():
    xa = np.array([0.0, 1.0, 3.0, 4.0]); ya = np.array([1.0, 3.0, -1.0, 0.0])
    y = C.c_double(); m = C.c_double()
    L.ora_linint(dptr(xa), dptr(ya), 4, 2.0, C.byref(y), C.byref(m))
    assert y.value == pytest.approx(1.0) and m.value == pytest.approx(-2.0)


def _winding(px, py, vx, vy):
    wn = 0
    n = len(vx)
    for i in range(n):
        x1, y1, x2, y2 = vx[i], vy[i], vx[(i + 1) % n], vy[(i + 1) % n]
        if y1 <= py:
            if y2 > py and (x2 - x1) * (py - y1) - (px - x1) * (y2 - y1) > 0:
                wn += 1
        elif y2 <= py and (x2 - x1) * (py - y1) - (px - x1) * (y2 - y1) < 0:
            wn -= 1
    return wn != 0


def test_gridcell_matches_winding_number_off_edges_and_edge_rules():
    rng = np.random.default_rng(1)
    for _ in range(300):
        c = rng.normal(size=2) * 100
        ex = c[0] + np.array([0, 10, 11, 1.0]) + rng.normal(size=4)
        ey = c[1] + np.array([0, 0.5, 9, 8.0]) + rng.normal(size=4)
        p = c + rng.random(2) * 14 - 2
        assert bool(L.ora_gridcell(dptr(ex), dptr(ey), p[0], p[1])) == _winding(p[0], p[1], ex, ey)
    ex = np.array([0.0, 2.0, 2.0, 0.0]); ey = np.array([0.0, 0.0, 2.0, 2.0])
    assert L.ora_gridcell(dptr(ex), dptr(ey), 0.0, 0.0) == 1        # on a node
    assert L.ora_gridcell(dptr(ex), dptr(ey), 1.0, 0.0) == 1        # on a horizontal edge
    assert L.ora_gridcell(dptr(ex), dptr(ey), 2.0, 1.0) == 1        # on a vertical edge
    assert L.ora_gridcell(dptr(ex), dptr(ey), 3.0, 0.0) == 0        # beyond the edge, same y
    assert L.ora_gridcell(dptr(ex), dptr(ey), 1.0, 1.0) == 1
    assert L.ora_gridcell(dptr(ex), dptr(ey), 1.0, 2.5) == 0


def test_inpoly_matches_winding_number_and_onin_switch():
    rng = np.random.default_rng(2)
    th = np.sort(rng.random(17) * 2 * np.pi)
    r = 5 + 3 * rng.random(17)
    vx, vy = r * np.cos(th), r * np.sin(th)
    cx, cy = np.append(vx, vx[0]), np.append(vy, vy[0])            # closed: first point repeated
    for _ in range(500):
        p = rng.random(2) * 20 - 10
        assert bool(L.ora_inpoly(p[0], p[1], len(cx), dptr(cx), dptr(cy), -1)) == _winding(p[0], p[1], vx, vy)
    sq_x = np.array([0.0, 0.0, 4.0, 4.0, 0.0]); sq_y = np.array([0.0, 4.0, 4.0, 0.0, 0.0])
    assert L.ora_inpoly(0.0, 0.0, 5, dptr(sq_x), dptr(sq_y), -1) == 1    # vertex: in by default
    assert L.ora_inpoly(0.0, 0.0, 5, dptr(sq_x), dptr(sq_y), 0) == 0     # onin = .FALSE.
    assert L.ora_inpoly(2.0, 4.0, 5, dptr(sq_x), dptr(sq_y), -1) in (0, 1)
    assert L.ora_inpoly(2.0, 2.0, 5, dptr(sq_x), dptr(sq_y), 0) == 1
    assert L.ora_inpoly(5.0, 2.0, 5, dptr(sq_x), dptr(sq_y), -1) == 0
    assert L.ora_inpoly(-1.0, 4.0, 5, dptr(sq_x), dptr(sq_y), -1) == 0   # ray through two vertices


def _fit(x, y):
    n = len(x); yp = np.zeros(n); sg = np.zeros(n)
    ier = C.c_int32(); se = C.c_int32(0)
    L.ora_tspsi(n, dptr(x), dptr(y), dptr(yp), dptr(sg), C.byref(ier), C.byref(se))
    return yp, sg, ier.value, se.value


def _hval(t, x, y, yp, sg):
    ier = C.c_int32()
    return L.ora_hval(t, len(x), dptr(x), dptr(y), dptr(yp), dptr(sg), C.byref(ier))


def _hpval(t, x, y, yp, sg):
    ier = C.c_int32()
    return L.ora_hpval(t, len(x), dptr(x), dptr(y), dptr(yp), dptr(sg), C.byref(ier))


def test_tension_spline_invariants():
    rng = np.random.default_rng(3)
    for n in (4, 84):
        for trial in range(20):
            x = np.cumsum(0.1 + rng.random(n))
            y = np.cumsum(rng.normal(size=n)) if trial % 2 else np.sin(x) + 0.1 * rng.normal(size=n)
            yp, sg, ier, se = _fit(x, y)
            assert ier == 0 and se == 0
            assert np.all(sg[:n - 1] >= 0) and np.all(sg[:n - 1] <= 85.0)
            for k in range(n):                                         # interpolates its knots
                assert abs(_hval(x[k], x, y, yp, sg) - y[k]) <= 1e-12 * max(1, abs(y[k]))
            for k in range(n - 1):                                     # shape preserving: stays within the
                t = x[k] + (x[k + 1] - x[k]) * rng.random()           # data range where data are monotone
                v = _hval(t, x, y, yp, sg)
                lo, hi = min(y[k], y[k + 1]), max(y[k], y[k + 1])
                if yp[k] * (y[k + 1] - y[k]) >= 0 and yp[k + 1] * (y[k + 1] - y[k]) >= 0:
                    assert lo - 1e-9 * (1 + abs(lo)) <= v <= hi + 1e-9 * (1 + abs(hi))
                h = 1e-6 * (x[k + 1] - x[k])                          # HPVAL is d/dt HVAL
                if x[k] + h < t < x[k + 1] - h:
                    num = (_hval(t + h, x, y, yp, sg) - _hval(t - h, x, y, yp, sg)) / (2 * h)
                    assert abs(num - _hpval(t, x, y, yp, sg)) <= 1e-5 * (1 + abs(num))
    x = np.array([0.0, 1.0, 2.5, 4.0]); y = 3.0 - 2.0 * x              # linear data reproduced exactly
    yp, sg, ier, se = _fit(x, y)
    assert np.allclose(yp, -2.0)
    assert _hval(1.7, x, y, yp, sg) == pytest.approx(3.0 - 3.4, abs=1e-13)


def test_snhcsh_matches_libm():
    s = C.c_double(); c = C.c_double(); cm = C.c_double()
    for v in (1e-3, 0.1, 0.49, 0.5, 0.51, 2.0, 10.0, -0.3, -3.0):
        L.ora_snhcsh(v, C.byref(s), C.byref(c), C.byref(cm))
        import math
        ser = lambda k0: sum(v ** k / math.factorial(k) for k in range(k0, 60, 2))   # noqa: E731
        assert s.value == pytest.approx(ser(3), rel=1e-12)
        assert c.value == pytest.approx(ser(2), rel=1e-12)
        assert cm.value == pytest.approx(ser(4), rel=1e-12)


def test_s_levels():
    # Vtransform 1: bottom w-level is the sea bed, top w-level is zeta (hydro:2709-2710)
    hc = np.float32(0.2)
    for vt in (1, 2, 3):
        zb = L.ora_slevel(0.3, -12.0, -1.0, -1.0, hc, vt)
        zt = L.ora_slevel(0.3, -12.0, 0.0, 0.0, hc, vt)
        assert zt == pytest.approx(0.3, abs=1e-12)
        if vt != 3:
            assert zb == pytest.approx(-12.0, abs=1e-12)
    # hc is REAL(4): 0.2f, not 0.2 (ledger 4)
    a = L.ora_slevel(0.0, -10.0, -0.5, -0.3, hc, 1)
    assert a == float(np.float32(0.2)) * -0.5 + (10.0 - float(np.float32(0.2))) * -0.3
