"""Second, independent restatement of the reference's LEAF numerics, in plain Python floats.

TEST INFRASTRUCTURE ONLY (same rule as the rest of oracle/).  Purpose (SURVEY.md 8c,
mitigation 1): the Fortran cannot be compiled here, so the C oracle (ora_leaf.c) is pinned by a
differential test against a second reading of the same Fortran lines, written separately in a
different language and shape.  Python floats are IEEE doubles, `math.exp / sqrt` are the C
library's, and the oracle is built with -ffp-contract=off, so the two must agree BIT FOR BIT
(tests/test_oracle_differential.py asserts equality, not closeness, wherever both sides follow
the reference's operation order).

Every function cites the reference lines it follows (/root/reference/Model/).
"""
import math

SBIG = 85.0
EPS_RTOL = 200.0 * 2.0 ** -53      # tension_module.f90:433-439: RTOL halves until 1 + RTOL <= 1, then x 200


# ---------------------------------------------------------------- interpolation_module.f90
def linint(xa, ya, x):
    """interpolation_module.f90:25-59 -> (y, m)"""
    jlo, jhi = 1, len(xa)
    while True:
        k = (jhi + jlo) // 2
        if xa[k - 1] > x:
            jhi = k
        else:
            jlo = k
        if jhi - jlo == 1:
            break
    m = (ya[jlo - 1] - ya[jhi - 1]) / (xa[jlo - 1] - xa[jhi - 1])
    b = ya[jlo - 1] - m * xa[jlo - 1]
    return m * x + b, m


def polintd(xa, ya, x):
    """interpolation_module.f90:70-107 (n = 3)"""
    ns, dif = 1, abs(x - xa[0])
    for i in range(1, 4):
        dift = abs(x - xa[i - 1])
        if dift < dif:
            ns, dif = i, dift
    c = (xa[1] - x) * ((ya[2] - ya[1]) / (xa[1] - xa[2]))
    c = c - (xa[1] - x) * ((ya[1] - ya[0]) / (xa[0] - xa[1]))
    c = c / (xa[0] - xa[2])
    if ns == 3:
        a = (ya[2] - ya[1]) / (xa[1] - xa[2]); b = xa[2] - x
    else:
        a = (ya[1] - ya[0]) / (xa[0] - xa[1]); b = xa[0] - x
    return ya[ns - 1] + (xa[ns - 1] - x) * a + b * c


# ---------------------------------------------------------------------- tension_module.f90
_P = (-3.51754964808151394800e5, -1.15614435765005216044e4, -1.63725857525983828727e2, -7.89474443963537015605e-1)
_Q = (-2.11052978884890840399e6, 3.61578279834431989373e4, -2.77711081420602794433e2, 1.0)


def snhcsh(x):
    """tension_module.f90:784-850 -> (sinh x - x, cosh x - 1, cosh x - 1 - x^2/2)"""
    ax = abs(x); xs = ax * ax
    if ax <= 0.5:
        xc = x * xs
        p = ((_P[3] * xs + _P[2]) * xs + _P[1]) * xs + _P[0]
        q = ((_Q[3] * xs + _Q[2]) * xs + _Q[1]) * xs + _Q[0]
        sinhm = xc * (p / q)
        xsd4 = 0.25 * xs; xsd2 = xsd4 + xsd4
        p = ((_P[3] * xsd4 + _P[2]) * xsd4 + _P[1]) * xsd4 + _P[0]
        q = ((_Q[3] * xsd4 + _Q[2]) * xsd4 + _Q[1]) * xsd4 + _Q[0]
        f = xsd4 * (p / q)
        coshmm = xsd2 * f * (f + 2.0)
        coshm = coshmm + xsd2
    else:
        expx = math.exp(ax)
        sinhm = -(((1.0 / expx + ax) + ax) - expx) / 2.0
        if x < 0.0:
            sinhm = -sinhm
        coshm = ((1.0 / expx - 2.0) + expx) / 2.0
        coshmm = coshm - xs / 2.0
    return sinhm, coshm, coshmm


def _sign(a, b):
    return abs(a) if (b > 0.0 or (b == 0.0 and math.copysign(1.0, b) > 0)) else -abs(a)


def ypc1(x, y):
    """tension_module.f90:852-978 -> (yp, ier)"""
    n = len(x); nm1 = n - 1
    yp = [0.0] * n
    dxi = x[1] - x[0]
    if dxi <= 0.0:
        return yp, 2
    si = (y[1] - y[0]) / dxi
    if nm1 == 1:
        return [si, si], 0
    dx2 = x[2] - x[1]
    if dx2 <= 0.0:
        return yp, 3
    s2 = (y[2] - y[1]) / dx2
    t = si + dxi * (si - s2) / (dxi + dx2)
    yp[0] = min(max(0.0, t), 3.0 * si) if si >= 0.0 else max(min(0.0, t), 3.0 * si)
    sim1 = dxim1 = 0.0
    for i in range(2, nm1 + 1):
        dxim1 = dxi
        dxi = x[i] - x[i - 1]
        if dxi <= 0.0:
            return yp, i + 1
        sim1 = si
        si = (y[i] - y[i - 1]) / dxi
        t = (dxim1 * si + dxi * sim1) / (dxim1 + dxi)
        asim1, asi = abs(sim1), abs(si)
        sgn = _sign(1.0, si)
        if asim1 > asi:
            sgn = _sign(1.0, sim1)
        if sgn > 0.0:
            yp[i - 1] = min(max(0.0, t), 3.0 * min(asim1, asi))
        else:
            yp[i - 1] = max(min(0.0, t), -3.0 * min(asim1, asi))
    t = si + dxi * (si - sim1) / (dxim1 + dxi)
    yp[n - 1] = min(max(0.0, t), 3.0 * si) if si >= 0.0 else max(min(0.0, t), 3.0 * si)
    return yp, 0


def convexity_newton(T, nit_cap=10000):
    """The Newton loop of SIGS' convexity branch (tension_module.f90:520-579) for one interval with
    T = max(D1/D2, D2/D1) > 2.  -> (SIG, failed): failed <=> NIT > 10000, the reference's SigErr."""
    rtol, ftol = EPS_RTOL, 0.0
    tp1 = T + 1.0
    sig = math.sqrt(10.0 * T - 20.0)
    nit = 0
    while True:
        if sig <= 0.5:
            sinhm, coshm, coshmm = snhcsh(sig)
            t1 = coshm / sinhm
            fp = t1 + sig * (sig / sinhm - t1 * t1 + 1.0)
        else:
            ems = math.exp(-sig)
            ssm = 1.0 - ems * (ems + sig + sig)
            t1 = (1.0 - ems) * (1.0 - ems) / ssm
            fp = t1 + sig * (2.0 * sig * ems / ssm - t1 * t1 + 1.0)
        f = sig * t1 - tp1
        nit += 1
        if nit > nit_cap:
            return sig, True
        if fp <= 0.0:
            return sig, False
        dsig = -f / fp
        if abs(dsig) <= rtol * sig or (0.0 <= f <= ftol) or abs(f) <= rtol:
            return sig, False
        sig = sig + dsig


def sigs(x, y, yp, secant_cap=100000):
    """tension_module.f90:314-782 with TOL = 0 -> (sigma, sigerr).  CONT starts .TRUE. in every
    interval (ledger 18: uninitialised in the reference on the SIG <= .5 secant branch).  The
    secant loop has no cap in the reference; `secant_cap` passes (or a NaN residual, which can never
    leave the loop) are reported as SigErr, as in the C oracle."""
    n = len(x); nm1 = n - 1
    sigma = [0.0] * n
    rtol, ftol = EPS_RTOL, 0.0
    for i in range(nm1):
        dx = x[i + 1] - x[i]
        if dx <= 0.0:
            return sigma, 0
        s1, s2 = yp[i], yp[i + 1]
        s = (y[i + 1] - y[i]) / dx
        d1 = s - s1; d2 = s2 - s; d1d2 = d1 * d2
        if (d1d2 == 0.0 and s1 != s2) or (s == 0.0 and s1 * s2 > 0.0):
            sigma[i] = SBIG
            continue
        if d1d2 >= 0.0:
            if d1d2 == 0.0:
                continue
            t = max(d1 / d2, d2 / d1)
            if t <= 2.0:
                continue
            sig, failed = convexity_newton(t)
            if failed:
                return sigma, 1                      # RETURN: the remaining intervals keep SIGMA = 0
            sig = min(sig, SBIG)
            if sig > 0.0:
                sigma[i] = sig
            continue
        # monotonicity, :638-760
        if s1 * s < 0.0 or s2 * s < 0.0:
            continue
        t0 = 3.0 * s - s1 - s2
        d0 = t0 * t0 - s1 * s2
        if d0 <= 0.0 or s * t0 >= 0.0:
            continue
        sgn = _sign(1.0, s)
        sig = SBIG
        fmax = sgn * (sig * s - s1 - s2) / (sig - 2.0)
        if fmax <= 0.0:
            sigma[i] = SBIG
            continue
        stol = rtol * sig
        f = fmax
        f0 = sgn * d0 / (3.0 * (d1 - d2))
        fneg = f0
        dsig = sig; dmax = sig
        d1pd2 = d1 + d2
        nit = 0
        cont = True
        a = e = c1 = c2 = 0.0
        err = False
        while True:
            dsig = -f * dsig / (f - f0)
            if abs(dsig) > abs(dmax) or dsig * dmax > 0.0:
                dsig = dmax; f0 = fneg
                nit += 1
                if nit > secant_cap:
                    err = True; break
                continue
            if abs(dsig) < stol / 2.0:
                dsig = -_sign(stol / 2.0, dmax)
            sig = sig + dsig
            f0 = f
            if sig <= 0.5:
                sinhm, coshm, coshmm = snhcsh(sig)
                c1 = sig * coshm * d2 - sinhm * d1pd2
                c2 = sig * (sinhm + sig) * d2 - coshm * d1pd2
                a = c2 - c1
                e = sig * sinhm - coshmm - coshmm
            else:
                ems = math.exp(-sig); ems2 = ems + ems; tm = 1.0 - ems
                ssinh = tm * (1.0 + ems); ssm = ssinh - sig * ems2; scm = tm * tm
                c1 = sig * scm * d2 - ssm * d1pd2
                c2 = sig * ssinh * d2 - scm * d1pd2
                f = fmax
                cont = True
                if c1 * (sig * scm * d1 - ssm * d1pd2) >= 0.0:
                    cont = False
                if cont:
                    a = ems2 * (sig * tm * d2 + (tm - sig) * d1pd2)
                if a * (c2 + c1) < 0.0:
                    cont = False
                if cont:
                    e = sig * ssinh - scm - scm
            if cont:
                arg = a * (c2 + c1)
                root = math.sqrt(arg) if arg >= 0.0 else float("nan")
                f = (sgn * (e * s2 - c2) + root) / e
            nit += 1
            if nit > secant_cap:
                err = True; break
            stol = rtol * sig
            if abs(dmax) <= stol or (0.0 <= f <= ftol) or abs(f) <= rtol:
                break
            if f != f:
                err = True; break
            dmax = dmax + dsig
            if f0 * f > 0.0 and abs(f) >= abs(f0):
                dsig = dmax; f0 = fneg
                continue
            if f0 * f <= 0.0:
                t1, t2 = dmax, fneg
                dmax = dsig; fneg = f0
                if abs(dsig) > abs(t1) and abs(f) < abs(t2):
                    dsig = t1; f0 = t2
        if err:
            return sigma, 1
        sig = min(sig, SBIG)
        if sig > 0.0:
            sigma[i] = sig
    return sigma, 0


def tspsi(x, y):
    """tension_module.f90:231-312 -> (yp, sigma, ier, sigerr)"""
    if len(x) < 2:
        return [], [], -1, 0
    yp, ierr = ypc1(x, y)
    if ierr != 0:
        return yp, [0.0] * len(x), -4, 0
    sigma, sigerr = sigs(x, y, yp)
    return yp, sigma, 0, sigerr


def intrvl(t, x):
    """tension_module.f90:1287-1354 without the SAVEd cache (ledger 20): 1-based I with X(I) <= T < X(I+1)"""
    il, ih = 1, len(x)
    while ih > il + 1:
        k = (il + ih) // 2
        if t < x[k - 1]:
            ih = k
        else:
            il = k
    return il


def _interval(t, x):
    n = len(x)
    if t < x[0]:
        return 1
    if t > x[n - 1]:
        return n - 1
    return intrvl(t, x)


def hval(t, x, y, yp, sigma):
    """tension_module.f90:1002-1119"""
    i = _interval(t, x) - 1
    dx = x[i + 1] - x[i]
    u = t - x[i]
    b2 = u / dx; b1 = 1.0 - b2
    y1 = y[i]; s1 = yp[i]
    s = (y[i + 1] - y1) / dx
    d1 = s - s1; d2 = yp[i + 1] - s
    sig = abs(sigma[i])
    if sig < 1.0e-9:
        return y1 + u * (s1 + b2 * (d1 + b1 * (d1 - d2)))
    if sig <= 0.5:
        sb2 = sig * b2
        sm, cm, cmm = snhcsh(sig)
        sm2, cm2, _ = snhcsh(sb2)
        e = sig * sm - cmm - cmm
        return y1 + s1 * u + dx * ((cm * sm2 - sm * cm2) * (d1 + d2) + sig * (cm * cm2 - (sm + sig) * sm2) * d1) / (sig * e)
    sb1 = sig * b1; sb2 = sig - sb1
    if -sb1 > SBIG or -sb2 > SBIG:
        return y1 + s * u
    e1 = math.exp(-sb1); e2 = math.exp(-sb2); ems = e1 * e2
    tm = 1.0 - ems; ts = tm * tm; tp = 1.0 + ems
    e = tm * (sig * tp - tm - tm)
    return y1 + s * u + dx * (tm * (tp - e1 - e2) * (d1 + d2) + sig * ((e2 + ems * (e1 - 2.0) - b1 * ts) * d1
                                                                           + (e1 + ems * (e2 - 2.0) - b2 * ts) * d2)) / (sig * e)


def hpval(t, x, y, yp, sigma):
    """tension_module.f90:1122-1251"""
    i = _interval(t, x) - 1
    dx = x[i + 1] - x[i]
    b1 = (x[i + 1] - t) / dx; b2 = 1.0 - b1
    s1 = yp[i]
    s = (y[i + 1] - y[i]) / dx
    d1 = s - s1; d2 = yp[i + 1] - s
    sig = abs(sigma[i])
    if sig < 1.0e-9:
        return s1 + b2 * (d1 + d2 - 3.0 * b1 * (d2 - d1))
    if sig <= 0.5:
        sb2 = sig * b2
        sm, cm, cmm = snhcsh(sig)
        sm2, cm2, _ = snhcsh(sb2)
        sinh2 = sm2 + sb2
        e = sig * sm - cmm - cmm
        return s1 + ((cm * cm2 - sm * sinh2) * (d1 + d2) + sig * (cm * sinh2 - (sm + sig) * cm2) * d1) / e
    sb1 = sig * b1; sb2 = sig - sb1
    if -sb1 > SBIG or -sb2 > SBIG:
        return s
    e1 = math.exp(-sb1); e2 = math.exp(-sb2); ems = e1 * e2
    tm = 1.0 - ems
    e = tm * (sig * (1.0 + ems) - tm - tm)
    return s + (tm * ((e2 - e1) * (d1 + d2) + tm * (d1 - d2)) + sig * ((e1 * ems - e2) * d1 + (e1 - e2 * ems) * d2)) / e


# ------------------------------------------- hydrodynamic_module.f90 (setInterp / getInterp / interp)
def _tri(xp, yp, xa, ya, xb, yb, xc, yc):
    """barycentric pair of (xp, yp) in the triangle with origin a and edges to b (t) and c (u), hydro:1706-1707"""
    t = ((xp - xa) * (yc - ya) + (ya - yp) * (xc - xa)) / ((xb - xa) * (yc - ya) - (yb - ya) * (xc - xa))
    u = ((xp - xa) * (yb - ya) + (ya - yp) * (xb - xa)) / ((xc - xa) * (yb - ya) - (yc - ya) * (xb - xa))
    return t, u


def _idw(x, y, xp, yp):
    dis = [1.0 / math.sqrt((x[i] - xp) * (x[i] - xp) + (y[i] - yp) * (y[i] - yp)) for i in range(4)]   # **2 is a product in Fortran
    td = dis[0] + dis[1] + dis[2] + dis[3]
    return [dis[0] / td, dis[1] / td, dis[2] / td, dis[3] / td]


def set_get_interp(x, y, v, xp, yp):
    """setInterp (hydro:1680-1740) then getInterp's combination (:1996-2004).  Note the reference's quirk: a point
    ON a node that lies outside both triangles' (t, u) tests gets Wgt = 1 for that node but tOK stays 2, so
    getInterp still combines with the second triangle's t and u."""
    wgt = [0.0, 0.0, 0.0, 0.0]
    t, u = _tri(xp, yp, x[0], y[0], x[1], y[1], x[2], y[2]); tok = 1
    if t < 0.0 or u < 0.0 or (t + u) > 1.0:
        t, u = _tri(xp, yp, x[2], y[2], x[3], y[3], x[0], y[0]); tok = 2
        if t < 0.0 or u < 0.0 or (t + u) > 1.0:
            on = [xp == x[i] and yp == y[i] for i in range(4)]
            if any(on):
                for i in range(4):
                    if on[i]:
                        wgt[i] = 1.0
            else:
                wgt = _idw(x, y, xp, yp); tok = 3
    if tok == 1:
        return v[0] + (v[1] - v[0]) * t + (v[2] - v[0]) * u
    if tok == 2:
        return v[2] + (v[3] - v[2]) * t + (v[0] - v[2]) * u
    return wgt[0] * v[0] + wgt[1] * v[1] + wgt[2] * v[2] + wgt[3] * v[3]


def interp_quad(x, y, v, xp, yp):
    """interp (hydro:2533-2565): the same three methods with the weights worked out per call; on a node: its value"""
    tt, uu = _tri(xp, yp, x[0], y[0], x[1], y[1], x[2], y[2])
    vp = v[0] + (v[1] - v[0]) * tt + (v[2] - v[0]) * uu
    if tt < 0.0 or uu < 0.0 or (tt + uu) > 1.0:
        tt, uu = _tri(xp, yp, x[2], y[2], x[3], y[3], x[0], y[0])
        vp = v[2] + (v[3] - v[2]) * tt + (v[0] - v[2]) * uu
        if tt < 0.0 or uu < 0.0 or (tt + uu) > 1.0:
            on = [xp == x[i] and yp == y[i] for i in range(4)]
            if any(on):
                for i in range(4):
                    if on[i]:
                        vp = v[i]
            else:
                w = _idw(x, y, xp, yp)
                vp = w[0] * v[0] + w[1] * v[1] + w[2] * v[2] + w[3] * v[3]
    return vp


# ------------------------------------------------------- hydrodynamic_module.f90 (WCTS_ITPI)
def wcts_profile(zb, zc, zf, vb, vc, vf, P_zb, P_zc, P_zf, ex, ix, p, v):
    """hydrodynamic_module.f90:2619-2689: the 4-knot tension spline of one field at the three hydro times (linint
    when SIGS reports SigErr), then the time polynomial.  -> (value, SigErr fall-backs among the profiles used)"""
    vals, nfall = [], 0
    for k, (z, y, T) in enumerate(((zb, vb, P_zb), (zc, vc, P_zc), (zf, vf, P_zf))):
        yp, sigm, ier, sigerr = tspsi(list(z), list(y))
        if sigerr == 0:
            vals.append(hval(T, list(z), list(y), yp, sigm))
        else:
            vals.append(linint(list(z), list(y), T)[0])
            if not (k == 2 and p == 1):
                nfall += 1
    ey = [vals[0], vals[0], vals[1]] if p == 1 else [vals[0], vals[1], vals[2]]
    b, c, f = polintd(ex, ey, ix[0]), polintd(ex, ey, ix[1]), polintd(ex, ey, ix[2])
    if v == 1:
        return b, nfall
    if v == 2:
        return c, nfall
    if v == 3:
        return f, nfall
    return (b + c * 4 + f) / 6.0, nfall


# ------------------------------------------------------------------------ behavior_module.f90
def behave(prm, state, P_S, words, Zpar, P_zc, P_zetac, P_age, P_depth, P_U, P_V, P_angle, it, daytime):
    """behavior_module.f90:181-551 for one particle.  prm: namelist values (dt, idt, twistart, twiend, Em, PI,
    daylength, Kd, thresh, Sgradient, swimfast, swimslow, swimstart, sink, Hswimspeed, Swimdepth, pediage, deadage);
    state = dict(behave, swim3, timer, Sprev, zprev, bottom) updated in place; words: genrand_int32 values in draw
    order (genrand_real1 = word / 4294967295, random_module.f90:213-220).  Single-precision literals of the source
    (0.80, 0.20, 0.1, 0.49, 0.05 ...) enter as the doubles they widen to; `1.5*24.*3600.` and the like are REAL(4)
    products that happen to be exact.  -> (XBehav, YBehav, ZBehav, bott, words used)"""
    f32 = _f32
    k = [0]

    def real1():
        w = words[k[0]]; k[0] += 1
        return float(w) / 4294967295.0

    XBehav = YBehav = ZBehav = 0.0
    bott = False
    # initBehave :118-131 (per-particle constants are the namelist values in v.2b)
    pediage, deadage = prm["pediage"], prm["deadage"]
    swim1 = (prm["swimfast"] - prm["swimslow"]) / (pediage - prm["swimstart"])
    swim2 = prm["swimfast"] - swim1 * pediage
    if P_age >= prm["swimstart"]:                                          # :214-215
        state["swim3"] = swim1 * P_age + swim2
    if P_age >= pediage:
        state["swim3"] = prm["swimfast"]
    if state["behave"] in (4, 5):                                          # :220-229
        if P_age >= pediage and P_age < deadage:
            state["behave"] = 2
        state["timer"] = max(0.0, state["timer"] - float(prm["dt"]))
    B = state["behave"]
    swim3 = state["swim3"]

    def updown(switch):
        """the recurring block: sign by one draw against `switch`, magnitude by a second draw"""
        negpos = 1.0
        dev1 = real1()
        if dev1 > switch:
            negpos = -1.0
        devB = real1()
        return negpos * devB * swim3

    parBehav = 0.0
    if B == 1:                                                             # :258-282
        parBehav = updown(f32(0.80)) if P_zc < (P_zetac - 1.0) else updown(0.5)
    if B == 2 or (B == 5 and state["timer"] > 0.0):                        # :286-311
        parBehav = updown(f32(0.20)) if P_zc > (P_depth + 1.0) else updown(0.5)
    if B == 3:                                                             # :314-353
        dtime = (daytime - math.trunc(daytime)) * 24.0
        if dtime > prm["twistart"] and dtime < prm["twiend"]:
            tst = (dtime - prm["twistart"]) * 3600.0
            sn = math.sin(prm["PI"] * tst / (prm["daylength"] * 3600.0))
            E0 = prm["Em"] * sn * sn
        else:
            E0 = 0.0
        P_light = E0 * math.exp(prm["Kd"] * P_zc)
        if P_light < prm["thresh"]:
            parBehav = updown(0.5)
        if P_light > prm["thresh"]:
            parBehav = updown(f32(0.20))
    if B == 4:                                                             # :356-407
        if it == 1:
            state["Sprev"] = P_S; state["zprev"] = P_zc
        Sslope = 0.0
        deltaS = state["Sprev"] - P_S
        deltaz = state["zprev"] - P_zc
        if it > 1:
            Sslope = (deltaS / deltaz) if deltaz != 0.0 else (math.copysign(math.inf, deltaS) * math.copysign(1.0, deltaz) if deltaS != 0.0 else math.nan)
        btest = 0
        if abs(Sslope) > prm["Sgradient"]:
            negpos = 1.0
            dev1 = real1()
            if dev1 > f32(0.80):
                negpos = -1.0
            parBehav = negpos * swim3
            btest = 1
        if btest == 0:
            negpos = 1.0
            dev1 = real1()
            if P_age < 1.5 * 24.0 * 3600.0:
                switch = f32(0.1)
            elif P_age < 5.0 * 24.0 * 3600.0:
                switch = f32(0.49)
            elif P_age < 8.0 * 24.0 * 3600.0:
                switch = f32(0.50)
            else:
                # DBLE(0.517) widens a REAL(4) literal: it is not the double 0.517
                switchslope = (f32(0.50) - f32(0.517)) / (8.0 * 24.0 * 3600.0 - pediage)
                switch = switchslope * P_age + f32(0.50) - switchslope * 8.0 * 24.0 * 3600.0
                if P_zc < P_depth + 1.0:
                    switch = f32(0.5)
            if dev1 > (1 - switch):
                negpos = -1.0
            devB = real1()
            parBehav = negpos * devB * swim3
        state["Sprev"] = P_S; state["zprev"] = P_zc
    if B == 5 and state["timer"] == 0.0:                                   # :410-463
        if it == 1:
            state["Sprev"] = P_S; state["zprev"] = P_zc
        Sslope = 0.0
        deltaS = state["Sprev"] - P_S
        deltaz = state["zprev"] - P_zc
        if it > 1:
            Sslope = (deltaS / deltaz) if deltaz != 0.0 else (math.copysign(math.inf, deltaS) * math.copysign(1.0, deltaz) if deltaS != 0.0 else math.nan)
        btest = 0
        if abs(Sslope) > prm["Sgradient"]:
            negpos = 1.0
            dev1 = real1()
            btest = 1
            state["timer"] = 2.0 * 3600.0
            if dev1 > f32(0.20):
                negpos = -1.0
            parBehav = negpos * swim3
            if P_age < 3.5 * 24.0 * 3600.0:
                btest = 0
                state["timer"] = 0.0
        if btest == 0:
            negpos = 1.0
            dev1 = real1()
            switch = f32(0.495)
            if P_age < 1.5 * 24.0 * 3600.0:
                switch = f32(0.9)
            if P_age > 2.0 * 24.0 * 3600.0 and P_age < 3.5 * 24.0 * 3600.0:
                switchslope = (f32(0.3) - f32(0.495)) / (2.0 * 24.0 * 3600.0 - 3.5 * 24.0 * 3600.0)
                switch = switchslope * P_age + f32(0.3) - switchslope * 2.0 * 24.0 * 3600.0
            if dev1 > switch:
                negpos = -1.0
            devB = real1()
            parBehav = negpos * devB * swim3
        state["Sprev"] = P_S; state["zprev"] = P_zc
    if B == 6:                                                             # :466-471
        parBehav = prm["sink"] if P_age >= prm["swimstart"] else swim3
    ZBehav = parBehav * prm["idt"]                                         # :491
    if B == 7:                                                             # :495-548
        if it == 1:
            state["Sprev"] = P_S
        ca, sa = math.cos(P_angle), math.sin(P_angle)
        X = P_U * ca - P_V * sa
        Y = P_U * sa + P_V * ca
        currentspeed = math.sqrt(X * X + Y * Y)
        if state["bottom"]:
            if state["Sprev"] < P_S:
                state["bottom"] = False
                ZBehav = P_depth + prm["Swimdepth"]
            else:
                ZBehav = -9999.0
        else:
            if currentspeed > f32(0.05):
                Hdistance = prm["Hswimspeed"] * prm["idt"]
                theta = math.atan(Y / X) if X != 0.0 else math.atan(math.copysign(math.inf, Y) if Y != 0.0 else math.nan)
                if X > 0.0:
                    XBehav = Hdistance * math.cos(theta); YBehav = Hdistance * math.sin(theta)
                if X < 0.0:
                    XBehav = -1.0 * Hdistance * math.cos(theta); YBehav = -1.0 * Hdistance * math.sin(theta)
                if X == 0 and Y >= 0.0:
                    XBehav = 0.0; YBehav = Hdistance
                if X == 0 and Y <= 0.0:
                    XBehav = 0.0; YBehav = -1.0 * Hdistance
                ZBehav = P_depth + prm["Swimdepth"]
            else:
                ZBehav = -9999.0
                state["bottom"] = True
        bott = state["bottom"]
    return XBehav, YBehav, ZBehav, bott, k[0]


# ------------------------------------------------------------------ LTRANS.f90 (find_currents)
def find_currents_column(us, ws, z0, Zpar, z, wz, u, v, w, P_zb, P_zc, P_zf, ex, ix, p, version):
    """LTRANS.f90:1422-1614 on a bare column.  z[t][k], wz[t][k]: rho- / w-level depths (t = 0, 1, 2 = back, centre,
    forward; k 0-based), u, v (us levels) and w (ws levels) the field values at the particle.
    -> (Uad, Vad, Wad, SigErr fall-backs)"""
    Z = lambda t, i: z[t][i - 1]          # 1-based level access like the Fortran
    WZ = lambda t, i: wz[t][i - 1]
    # lowest numbered level of the closest four, :1450-1467 (the DO variable is us - 1 / ws - 1 after a complete loop)
    i = 3
    while i <= us - 2:
        if Zpar < Z(0, i) or Zpar < Z(1, i) or Zpar < Z(2, i):
            break
        i += 1
    ii = i - 2
    i = 3
    while i <= ws - 2:
        if Zpar < WZ(0, i) or Zpar < WZ(1, i) or Zpar < WZ(2, i):
            break
        i += 1
    iii = i - 2
    nfall = 0
    if Zpar < WZ(0, 1) + z0 or Zpar < WZ(1, 1) + z0 or Zpar < WZ(2, 1) + z0:               # :1480-1486
        return 0.0, 0.0, 0.0, 0
    if Zpar < Z(0, 1) or Zpar < Z(1, 1) or Zpar < Z(2, 1):                                   # log layer :1489-1600
        num = lambda: math.log10((Zpar - WZ(0, 1)) / z0)
        out = []
        for fld, lev, top in ((u, 1, None), (v, 1, None), (w, 2, "w")):
            vals = []
            for t in range(3):
                # note: the reference measures every height from the BACK record's bed, Pwc_wzb(1)
                ref = (WZ(t, 2) if top == "w" else Z(t, 1)) - WZ(0, 1)
                vals.append(fld[t][lev - 1] * num() / math.log10(ref / z0))
            ey = [vals[0], vals[0], vals[1]] if p == 1 else vals
            out.append(polintd(ex, ey, ix[version - 1]))
        return out[0], out[1], out[2], 0
    res = []
    for fld, lvl, zz in ((u, ii, z), (v, ii, z), (w, iii, wz)):
        prof = [[fld[t][lvl - 1 + k] for k in range(4)] for t in range(3)]
        zs = [[zz[t][lvl - 1 + k] for k in range(4)] for t in range(3)]
        val, nf = wcts_profile(zs[0], zs[1], zs[2], prof[0], prof[1], prof[2], P_zb, P_zc, P_zf, ex, ix, p, version)
        res.append(val); nfall += nf
    return res[0], res[1], res[2], nfall


# --------------------------------------------------------------------- ver_turb_module.f90
def _f32(v):
    """a single-precision literal as the double it widens to"""
    import struct
    return struct.unpack("f", struct.pack("f", v))[0]


def vturb_column(ws, idt, p, ex, ix, khb, khc, khf, wzb, wzc, wzf, P_zc, P_depth, P_zetac, dev):
    """ver_turb_module.f90:30-380 from step ii. on (the KH gather of step i. is the caller's): the column fit and the
    random displacement loop for one particle, `dev` = the idt/2 normal deviates norm() would return, in draw
    order.  Written from the Fortran with 1-based lists (index 0 unused).  -> (TurbV, SigErr)"""
    background = _f32(1.0e-6)                                            # :43 REAL(4) PARAMETER
    p2 = ws * 4                                                          # :63
    one = lambda a: [None] + list(a)                                     # 1-based view
    KHb, KHc, KHf, Wb, Wc, Wf = one(khb), one(khc), one(khf), one(wzb), one(wzc), one(wzf)
    # ii.a :116-124 (float(j-4) is REAL(4): exact for these integers)
    newx = {}
    for t, W in (("b", Wb), ("c", Wc), ("f", Wf)):
        nx = [0.0] * (p2 + 8)
        for j in range(1, p2 + 8):
            nx[j] = W[1] + float(j - 4) * (W[ws] - W[1]) / float(p2)
        newx[t] = nx
    # :126-133
    slope, icpt = {}, {}
    for t, K, W in (("b", KHb, Wb), ("c", KHc, Wc), ("f", KHf, Wf)):
        sl, ic = [0.0] * ws, [0.0] * ws
        for i in range(1, ws):
            sl[i] = (K[i] - K[i + 1]) / (W[i] - W[i + 1])
            ic[i] = K[i] - sl[i] * W[i]
        slope[t], icpt[t] = sl, ic
    # :135-166 the walking jlo
    newy = {}
    for t, W in (("b", Wb), ("c", Wc), ("f", Wf)):
        ny = [0.0] * (p2 + 8)
        jlo = 1
        for j in range(5, p2 + 4):
            while not (W[jlo + 1] > newx[t][j]):
                jlo += 1
            ny[j] = slope[t][jlo] * newx[t][j] + icpt[t][jlo]
        newy[t] = ny
    # ii.b :173-180: the lower pads take KHb(1) at all three times
    for i in range(1, 5):
        newy["b"][i] = KHb[1]; newy["c"][i] = KHb[1]; newy["f"][i] = KHb[1]
        newy["b"][i + p2 + 3] = KHb[ws]; newy["c"][i + p2 + 3] = KHc[ws]; newy["f"][i + p2 + 3] = KHf[ws]
    # iii. :187-197, iv. :200-213
    movex, movey = {}, {}
    for t, K, W in (("b", KHb, Wb), ("c", KHc, Wc), ("f", KHf, Wf)):
        mx, my = [0.0] * (p2 + 2), [0.0] * (p2 + 2)
        ny, nx = newy[t], newx[t]
        for i in range(2, p2):
            my[i] = (ny[i] + ny[i + 1] + ny[i + 2] + ny[i + 3] + ny[i + 4] + ny[i + 5] + ny[i + 6] + ny[i + 7]) / 8.0
            mx[i] = nx[i] + (nx[i + 7] - nx[i]) / 2.0
        mx[1] = W[1]; my[1] = K[1]; mx[p2] = W[ws]; my[p2] = K[ws]
        movex[t], movey[t] = mx, my
    # v. :223-262, clamps :264-268, vi. :272-275
    ifitx, ifity = [0.0] * p2, [0.0] * p2                                # 0-based: handed to the spline routines
    for k in range(1, p2 + 1):
        if p == 1:
            eyx = [movex["b"][k], movex["b"][k], movex["c"][k]]; eyy = [movey["b"][k], movey["b"][k], movey["c"][k]]
        else:
            eyx = [movex["b"][k], movex["c"][k], movex["f"][k]]; eyy = [movey["b"][k], movey["c"][k], movey["f"][k]]
        xb, xc, xf = polintd(ex, eyx, ix[0]), polintd(ex, eyx, ix[1]), polintd(ex, eyx, ix[2])
        yb, yc, yf = polintd(ex, eyy, ix[0]), polintd(ex, eyy, ix[1]), polintd(ex, eyy, ix[2])
        if yb < 0.0:
            yb = 0.0
        if yc < 0.0:
            yc = 0.0
        if yf < 0.0:
            yf = 0.0
        ifity[k - 1] = (yb + 4.0 * yc + yf) / 6.0
        ifitx[k - 1] = (xb + 4.0 * xc + xf) / 6.0
    # vii. :278-279
    ypk, sigk, ier, sigerr = tspsi(ifitx, ifity)
    # viii. - x. :282-337
    deltat = 2.0
    loop = idt // int(deltat)
    ParZc = P_zc
    for i in range(1, loop + 1):
        if ParZc < P_depth or ParZc > P_zetac:
            Kprimec = 0.0
        elif sigerr == 0:
            Kprimec = hpval(ParZc, ifitx, ifity, ypk, sigk)
        else:
            _, Kprimec = linint(ifitx, ifity, ParZc)
        KprimeZc = -1.0 * Kprimec * deltat
        Z3rdc = ParZc + 0.5 * KprimeZc
        if Z3rdc < P_depth or Z3rdc > P_zetac:
            KH3rdc = background
        else:
            if sigerr == 0:
                KH3rdc = hval(Z3rdc, ifitx, ifity, ypk, sigk)
            else:
                KH3rdc, _ = linint(ifitx, ifity, Z3rdc)
            if KH3rdc < background:
                KH3rdc = background
        r = 1.0
        ParZc = ParZc + KprimeZc + dev[i - 1] * math.pow(2.0 / r * KH3rdc * deltat, 0.5)
    return P_zc - ParZc, sigerr                                          # xi. :342


# ---------------------------------------------------------------------- gridcell_module.f90
def gridcell(ex, ey, X, Y):
    """gridcell_module.f90:26-257 for ONE element (checkele form): True <=> triangle /= 0"""
    if all(Y < v for v in ey) or all(Y > v for v in ey):
        return False
    if all(X < v for v in ex) or all(X > v for v in ex):
        return False
    for k in range(4):
        if X == ex[k] and Y == ey[k]:
            return True
    for a, b in ((0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)):           # :72-143, in this order
        if ey[a] == ey[b] and Y == ey[a]:
            return (ex[a] > ex[b] and ex[b] < X < ex[a]) or (ex[b] > ex[a] and ex[a] < X < ex[b])
    if Y in (ey[0], ey[1], ey[2], ey[3]):                                   # :147-176
        if Y == max(ey) or Y == min(ey):
            return False
    hit = False
    counter = [0, 0, 0, 0]
    for p in range(4):                                                      # :178-243
        bx1, by1, bx2, by2 = ex[p], ey[p], ex[(p + 1) % 4], ey[(p + 1) % 4]
        if X <= bx1 or X <= bx2:
            if (by1 > by2 and by2 <= Y <= by1) or (by2 > by1 and by1 <= Y <= by2):
                if bx1 == bx2:
                    if X == bx1:
                        hit = True; break
                    counter[p] = 0 if Y == by2 else 1
                else:
                    slope = (by1 - by2) / (bx1 - bx2)
                    xi = (Y - by1 + (slope * bx1)) / slope
                    if xi > X:
                        counter[p] = 0 if Y == by2 else 1
                    if xi == X:
                        hit = True; break
    return hit or (sum(counter) % 2 != 0)


# -------------------------------------------------------------- point_in_polygon_module.f90
def inpoly(x, y, e, onin=None):
    """point_in_polygon_module.f90:25-167; e = [(x, y), ...] closed polygon (first point repeated)"""
    n = len(e)
    onout = (not onin) if onin is not None else False
    hilo = [0] * (n + 2)                      # 1-based, hilo[0] stands for hilo(0) (never reached: `first`)
    on = False
    for i in range(1, n + 1):
        ex_, ey_ = e[i - 1]
        if ey_ > y:
            hilo[i] = 1
        if ey_ < y:
            hilo[i] = -1
        if ey_ == y and ex_ > x:
            on = True
        if ex_ == x and ey_ == y:
            return not onout
    crossed = 0
    if on:
        first = True
        i = 1
        while True:
            if i > n:
                break
            if hilo[i] == 0 and e[i - 1][0] > x:
                if first:
                    i += 1
                    continue
                if hilo[i - 1] == 0:
                    return not onout
                j = 1
                while True:
                    if i + j == n + 1:
                        j = 2 - i
                    if hilo[i + j] != 0:
                        break
                    if e[i + j - 1][0] < x:
                        return not onout
                    j += 1
                if hilo[i - 1] + hilo[i + j] == 0:
                    crossed += 1
                if j < 0:
                    break
                i = i + j
            first = False
            i += 1
    for i in range(1, n):
        ax, ay = e[i - 1]; bx, by = e[i]
        if ax <= x and bx <= x:
            continue
        if ay <= y and by <= y:
            continue
        if ay >= y and by >= y:
            continue
        if ax > x and bx > x:
            crossed += 1
            continue
        m = (by - ay) / (bx - ax)
        b = ay - m * ax
        ix = (y - b) / m
        if ix == x:
            return not onout
        if ix > x:
            crossed += 1
    return crossed % 2 != 0


# ------------------------------------------------------------------------ boundary_module.f90
def _sq(v):
    return v * v


def intersect_reflect(bnd, land, Xpos, Ypos, nXpos, nYpos, skipbound=0):
    """boundary_module.f90:1620-1902.  bnd = [(x1, y1, x2, y2), ...]; skipbound 1-based (0 = none).
    -> (intersectf, fiX, fiY, frX, frY, skipbound, isWater).  rPxyz keeps its previous value when
    dist1 == dist2 (ledger 19): it starts undefined in the reference, here as NaN."""
    fiX = fiY = frX = frY = -999999.0
    dtest = 999999.0
    isWater = False
    intersectf = 0
    skipi = skipbound
    xhigh, xlow = (Xpos, nXpos) if Xpos >= nXpos else (nXpos, Xpos)
    yhigh, ylow = (Ypos, nYpos) if Ypos >= nYpos else (nYpos, Ypos)
    rPx = rPy = float("nan")
    ix = iy = float("nan")
    Mbc = 0.0
    for i in range(1, len(bnd) + 1):
        if i == skipbound:
            continue
        intersect = 0
        bcx1, bcy1, bcx2, bcy2 = bnd[i - 1]
        if ((bcx1 > xhigh and bcx2 > xhigh) or (bcx1 < xlow and bcx2 < xlow) or
                (bcy1 > yhigh and bcy2 > yhigh) or (bcy1 < ylow and bcy2 < ylow)):
            continue
        bxhigh, bxlow = (bcx1, bcx2) if bcx1 >= bcx2 else (bcx2, bcx1)
        byhigh, bylow = (bcy1, bcy2) if bcy1 >= bcy2 else (bcy2, bcy1)

        def inside():
            return (xlow <= ix <= xhigh and ylow <= iy <= yhigh and bxlow <= ix <= bxhigh and bylow <= iy <= byhigh)

        def pick(rx1, ry1, rx2, ry2):
            nonlocal rPx, rPy
            dist1 = math.sqrt(_sq(ix - rx1) + _sq(iy - ry1))
            dist2 = math.sqrt(_sq(ix - rx2) + _sq(iy - ry2))
            if dist1 < dist2:
                rPx, rPy = rx1, ry1
            elif dist1 > dist2:
                rPx, rPy = rx2, ry2

        def oblique():
            distBC = math.sqrt(_sq(bcx1 - bcx2) + _sq(bcy1 - bcy2))
            crossk = ((nXpos - bcx1) * (bcy2 - bcy1)) - ((bcx2 - bcx1) * (nYpos - bcy1))
            dPBC = math.sqrt(_sq(crossk)) / distBC
            mP = -1.0 / Mbc
            bP = nYpos - mP * nXpos
            r = math.sqrt(_sq(2.0 * dPBC) / (1.0 + _sq(mP)))
            rx1 = r + nXpos; ry1 = mP * rx1 + bP
            rx2 = r * -1.0 + nXpos; ry2 = mP * rx2 + bP
            pick(rx1, ry1, rx2, ry2)

        if bcx1 == bcx2 or nXpos == Xpos:
            if bcx1 == bcx2 and nXpos == Xpos:
                continue
            if bcx1 == bcx2 and nYpos == Ypos:
                ix, iy = bcx1, nYpos
                if inside():
                    dPBC = math.sqrt(_sq(ix - nXpos) + _sq(iy - nYpos))
                    pick(nXpos + 2.0 * dPBC, nYpos, nXpos - 2.0 * dPBC, nYpos)
                    intersect = 1
            elif nXpos == Xpos and bcy1 == bcy2:
                ix, iy = nXpos, bcy1
                if inside():
                    dPBC = math.sqrt(_sq(ix - nXpos) + _sq(iy - nYpos))
                    pick(nXpos, nYpos + 2.0 * dPBC, nXpos, nYpos - 2.0 * dPBC)
                    intersect = 1
            elif bcx1 == bcx2 and nYpos != Ypos:
                Mp = (nYpos - Ypos) / (nXpos - Xpos)
                Bp = Ypos - Mp * Xpos
                ix = bcx1; iy = Mp * ix + Bp
                if inside():
                    dPBC = nXpos - ix
                    pick(nXpos + 2.0 * dPBC, nYpos, nXpos - 2.0 * dPBC, nYpos)
                    intersect = 1
            elif nXpos == Xpos and bcy1 != bcy2:
                Mbc = (bcy2 - bcy1) / (bcx2 - bcx1)
                Bbc = bcy2 - Mbc * bcx2
                ix = nXpos; iy = Mbc * ix + Bbc
                if inside():
                    oblique()
                    intersect = 1
        else:
            Mbc = (bcy2 - bcy1) / (bcx2 - bcx1)
            Bbc = bcy2 - Mbc * bcx2
            Mp = (nYpos - Ypos) / (nXpos - Xpos)
            Bp = Ypos - Mp * Xpos
            ix = (Bbc - Bp) / (Mp - Mbc)
            iy = Mp * ix + Bp
            if Mbc == 0.0:
                iy = byhigh
            if inside():
                if Mbc == 0.0:
                    dPBC = nYpos - bcy1
                    pick(nXpos, nYpos + 2.0 * dPBC, nXpos, nYpos - 2.0 * dPBC)
                else:
                    oblique()
                intersect = 1
        d_P = math.sqrt(_sq(Xpos - ix) + _sq(Ypos - iy))
        if intersect == 1 and d_P < dtest:
            fiX, fiY, frX, frY = ix, iy, rPx, rPy
            intersectf = 1
            dtest = d_P
            skipi = i
            isWater = not land[i - 1]
    return intersectf, fiX, fiY, frX, frY, skipi, isWater


# -------------------------------------------------------------------- hydrodynamic_module.f90
def slevel(zeta, depth, sc, cs, hc32, vtransform):
    """getSlevel / getWlevel (hydrodynamic_module.f90:2691-2777); hc is REAL(4) (ledger 4),
    passed here already widened from float32"""
    h = -1.0 * depth
    if vtransform == 1:
        S = hc32 * sc + (h - hc32) * cs
        return S + zeta * (1.0 + S / h)
    if vtransform == 2:
        S = (hc32 * sc + h * cs) / (hc32 + h)
        return zeta + (zeta + h) * S
    if vtransform == 3:
        return zeta * (1.0 + sc) + hc32 * sc + (h - hc32) * cs
    raise ValueError("Illegal Vtransform number")


# ------------------------------------------------------------------ settlement_module.f90 (step)
def test_settlement_point(hab, R_ele, P_age, settletime, holes_exist, Px, Py):
    """settlement_module.f90:485-622 (testSettlement, psettle, hsettle) for one point.  hab: the lists createPolySpecs
    builds, in the layout of the ABI (polys / holes column-major (5 | 6, edges); poly_start 1-based; elepoly / polyhole
    as CSR over 0-based polygon / hole indices; maxdis per polygon / hole).  -> polygon id, 0 for none"""
    if not (P_age >= settletime):
        return 0

    def first_hit(rows, starts, sizes, maxdis, cand, onin):
        for pi in cand:
            start, size = int(starts[pi]), int(sizes[pi])
            cx, cy = rows[1][start - 1], rows[2][start - 1]
            dis = math.sqrt(_sq(Px - cx) + _sq(Py - cy))
            if dis > maxdis[pi]:
                continue
            e = [(rows[3][start - 1 + j], rows[4][start - 1 + j]) for j in range(size)]
            if inpoly(Px, Py, e, onin=onin):
                return int(round(rows[0][start - 1])), pi
        return 0, -1

    cand = [int(hab["elepoly_idx"][q]) for q in range(int(hab["elepoly_ptr"][R_ele - 1]), int(hab["elepoly_ptr"][R_ele]))]
    polyin, pidx = first_hit(hab["polys"], hab["poly_start"], hab["poly_size"], hab["poly_maxdis"], cand, None)    # psettle
    if polyin <= 0:
        return 0
    if holes_exist:                                                                                                # hsettle
        hc = [int(hab["polyhole_idx"][q]) for q in range(int(hab["polyhole_ptr"][pidx]), int(hab["polyhole_ptr"][pidx + 1]))]
        holein, _ = first_hit(hab["holes"], hab["hole_start"], hab["hole_size"], hab["hole_maxdis"], hc, False)
        if holein != 0:
            return 0
    return polyin


# ---------------------------------------------------------------- settlement_module.f90 (set-up)
def create_poly_specs_literal(r_ele_x, r_ele_y, polys, maxbdis):
    """createPolySpecs' element loop as written (settlement_module.f90:296-402): every element against every
    row of `polys` ([id, cx, cy, ex, ey] per row), with the `polytail%num` skip, the edge-point test, and the
    corner test on the polygon's last row.  maxbdis: {id: radius}.  -> list of polygon ids per element."""
    pedges = len(polys)
    first, count = {}, {}
    for i, row in enumerate(polys):
        q = int(round(row[0]))
        first.setdefault(q, i)
        count[q] = count.get(q, 0) + 1
    out = []
    for e in range(len(r_ele_x)):
        ex, ey = list(r_ele_x[e]), list(r_ele_y[e])
        lst = []
        for j in range(pedges):
            q = int(round(polys[j][0]))
            if lst and lst[-1] == q:
                continue
            if gridcell(ex, ey, polys[j][3], polys[j][4]):
                lst.append(q)
                continue
            if j == pedges - 1 or polys[j][0] != polys[j + 1][0]:
                near = any(math.sqrt(_sq(ex[k] - polys[j][1]) + _sq(ey[k] - polys[j][2])) < maxbdis[q] for k in range(4))
                if near:
                    pts = [(polys[first[q] + k][3], polys[first[q] + k][4]) for k in range(count[q])]
                    if any(inpoly(ex[k], ey[k], pts) for k in range(4)):
                        lst.append(q)
        out.append(lst)
    return out
