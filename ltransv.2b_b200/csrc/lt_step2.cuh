// lt_step2.cuh -- the production step: the kernels of one internal time step, one thread per
// particle except the VTurb fit (lt_vturb.cuh), per-particle scratch (SoA) handed from one to the next.
//
//   k_advect : gates, setEle, setInterp, vertical clamp, RK4 advection (find_currents x4),
//              salinity / temperature, HTurb                      LTRANS.f90:778-1087
//   k_vbuild : the water-column fit of the Visser random displacement model, one column per warp
//   k_vwalk  : its 60 sub-steps (both in lt_vturb.cuh)            ver_turb_module.f90:30-380
//   k_finish : behave, vertical + horizontal reflection, bounds checks, commit, setEle at
//              the new position, settlement                        LTRANS.f90:1110-1382
//   (k_vturb, the round-1 per-thread VTurb kernel built on VtCtx below, is kept as the A/B reference
//    of the tests: LTGPU_VTURB_LEGACY=1; VtCtx::build also serves k_vwalk's rare window refit)
//
// Why several: the phases have very different register / instruction footprints
// (the fused v1 kernel stalled on instruction fetch half of the time, see
// profiles/r01_notes.md); split, each keeps its own occupancy and i-cache working set.
//
// Differences from the reference that stay inside the 1e-9 parity budget (they change
// results by O(1e-16) relative; DESIGN.md section 6 has the full list):
//   * the 3-point time polynomial is applied as Lagrange weights computed once per step
//     on the host (LtDev::LW) instead of polintd's divided differences per value;
//   * reciprocals come from qrcp()/qdiv() (<= 1 ulp) in the smooth numerics;
//   * VTurb: the 4 ws spline knots are uniform in the interior (an identity of the
//     reference's construction: movex(i) = z1 + (i - 0.5) R / p2), so abscissae are
//     evaluated, not stored;
//   * SigErr (Newton did not converge in 10000 iterations in SOME interval of the column,
//     which makes the reference fall back to linint for the whole column) is examined on
//     every interval of the fit by k_vbuild, as the reference does; the legacy VtCtx path
//     does so only with its sweep() (vturb_window_sigs = 0).
#pragma once
#include "lt_device.cuh"

struct ColK { double zb, zc, zf, depth, h; };
// The s-level tables (SC | CS | SCW | CSW of getSlevel / getWlevel, hydro:2691-2777) in shared memory: every
// kernel on the path starts with lev_tables_load().  As kernel parameters they sit in the constant bank,
// where a lane-dependent index is served one distinct address at a time (5 % of k_vbuild's and 7 % of
// k_advect's stall samples, profiles/r02_notes.md).
__shared__ double s_levtab[4 * LT_MAXLEV];
LT_DEV void lev_tables_load(const LtDev& D)
{
    const int ws = D.P.ws;                                              // us = ws - 1 entries of SC / CS are used
    for (int i = threadIdx.x; i < 4 * ws; i += blockDim.x) {
        const int t = i / ws, k = i - t * ws;
        s_levtab[t * LT_MAXLEV + k] = t == 0 ? D.SC[k] : t == 1 ? D.CS[k] : t == 2 ? D.SCW[k] : D.CSW[k];
    }
    __syncthreads();
}
LT_DEV double zlev2(const LtDev& D, const ColK& c, double zeta, double sc, double cs)
{
    double hc = (double)D.P.hc, S;
    if (D.P.Vtransform == 1) { S = hc * sc + (c.h - hc) * cs; return S + zeta * (1.0 + qdiv(S, c.h)); }
    if (D.P.Vtransform == 2) { S = qdiv(hc * sc + c.h * cs, hc + c.h); return zeta + (zeta + c.h) * S; }
    return zeta * (1.0 + sc) + hc * sc + (c.h - hc) * cs;
}
template <bool W>
LT_DEV void zlev3(const LtDev& D, const ColK& c, int k, double& zb, double& zc, double& zf)
{   // level k (0-based) at the three hydro times; S and 1 + S/h are shared
    const double sc = s_levtab[(W ? 2 * LT_MAXLEV : 0) + k], cs = s_levtab[(W ? 3 * LT_MAXLEV : LT_MAXLEV) + k];
    double hc = (double)D.P.hc;
    if (D.P.Vtransform == 1) {
        double S = hc * sc + (c.h - hc) * cs, q = 1.0 + qdiv(S, c.h);
        zb = S + c.zb * q; zc = S + c.zc * q; zf = S + c.zf * q;
    } else if (D.P.Vtransform == 2) {
        double S = qdiv(hc * sc + c.h * cs, hc + c.h);
        zb = c.zb + (c.zb + c.h) * S; zc = c.zc + (c.zc + c.h) * S; zf = c.zf + (c.zf + c.h) * S;
    } else {
        double a = 1.0 + sc, b = hc * sc + (c.h - hc) * cs;
        zb = c.zb * a + b; zc = c.zc * a + b; zf = c.zf * a + b;
    }
}
template <bool W>
LT_DEV int level_window2(const LtDev& D, const ColK& c, double Z, int n)
{   // LTRANS.f90:1451-1467 as a lower_bound (levels increase with the index)
    int lo = 3, hi = n - 1;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        double zb, zc, zf; zlev3<W>(D, c, mid - 1, zb, zc, zf);
        if (Z < zb || Z < zc || Z < zf) hi = mid; else lo = mid + 1;
    }
    return lo - 2;
}

// Lagrange weights applied to DIFFERENCES from the centre value: like polintd, this
// returns the data exactly when the three values are equal (a particle resting on the
// bed must not drift by an ulp), and the weights need not sum to one in floating point.
LT_DEV double lag(const double* w, double b, double c, double f) { return c + (w[0] * (b - c) + w[2] * (f - c)); }

// 4-knot spline value (TSPSI + HVAL of WCTS_ITPI, hydro:2619-2644) with the knot
// reciprocals shared between the fields interpolated on the same knots.
struct Knots4 { double x[4], r1, r2, r3, r12, r23; };
LT_DEV void knots_prepare(Knots4& k)
{
    double d1 = k.x[1] - k.x[0], d2 = k.x[2] - k.x[1], d3 = k.x[3] - k.x[2];
    k.r1 = qrcp(d1); k.r2 = qrcp(d2); k.r3 = qrcp(d3); k.r12 = qrcp(d1 + d2); k.r23 = qrcp(d2 + d3);
}
// the interval of a 4-knot spline that contains T, with its end slopes (YPC1) -- everything
// SIGS and HVAL need (tension:852-978, 1026-1041)
struct Iv4 { double X1, X2, Y1, Y2, P1, P2; };
LT_DEV void spline4_prepare(const Knots4& k, double y0, double y1, double y2, double y3, double T, Iv4& v)
{
    const double d1 = k.x[1] - k.x[0], d2 = k.x[2] - k.x[1], d3 = k.x[3] - k.x[2];
    const double s1 = (y1 - y0) * k.r1, s2 = (y2 - y1) * k.r2, s3 = (y3 - y2) * k.r3;
    const int I = (T < k.x[0]) ? 0 : (T > k.x[3]) ? 2 : (T < k.x[2] ? (T < k.x[1] ? 0 : 1) : 2);
    // interior-knot slope needed by every case: knot 1 for I = 0,1; knot 2 for I = 2
    const bool hi = I == 2;
    const double mA = ypc1_mid_r(hi ? d2 : d1, hi ? d3 : d2, hi ? s2 : s1, hi ? s3 : s2, hi ? k.r23 : k.r12);
    if (I == 0) { v.X1 = k.x[0]; v.X2 = k.x[1]; v.Y1 = y0; v.Y2 = y1; v.P1 = ypc1_end(s1, s1 + d1 * (s1 - s2) * k.r12); v.P2 = mA; }
    else if (I == 1) { v.X1 = k.x[1]; v.X2 = k.x[2]; v.Y1 = y1; v.Y2 = y2; v.P1 = mA; v.P2 = ypc1_mid_r(d2, d3, s2, s3, k.r23); }
    else { v.X1 = k.x[2]; v.X2 = k.x[3]; v.Y1 = y2; v.Y2 = y3; v.P1 = mA; v.P2 = ypc1_end(s3, s3 + d3 * (s3 - s2) * k.r23); }
}
LT_DEV double linint4(const Knots4& k, double y0, double y1, double y2, double y3, double T)
{   // linint fallback when SigErr (interpolation_module.f90:25-59), n = 4
    const double Y[4] = {y0, y1, y2, y3};
    int jlo = 1, jhi = 4;
    for (;;) { int q = (jhi + jlo) / 2; if (k.x[q - 1] > T) jhi = q; else jlo = q; if (jhi - jlo == 1) break; }
    double m = (Y[jlo - 1] - Y[jhi - 1]) / (k.x[jlo - 1] - k.x[jhi - 1]);
    return m * T + (Y[jlo - 1] - m * k.x[jlo - 1]);
}

// The reference's SIGS sweeps ALL intervals of the profile and leaves at the first one whose
// convexity Newton loop does not converge (SigErr, tension:556-559); WCTS_ITPI then replaces the
// whole profile by linint (hydro:2621-2644) even when the failing interval is not the one that
// holds the particle.  That loop's verdict is a pure function of T = max(D1/D2, D2/D1); scanned
// over 8e7 values it fails only for T in [2.02498, 2.04625] (about 1e-3 of the doubles there,
// around SIG = 0.5 where the two branches of the iteration meet), never outside.  So the other
// intervals need the Newton loop only when their T lies in LT_BAND (1.4 % of intervals).
#define LT_BAND_LO 2.02
#define LT_BAND_HI 2.06
LT_DEV bool sigerr_candidate(double s, double ypa, double ypb, double& T)
{   // interval with chord slope s and end slopes ypa, ypb: convexity case with T in the band?
    const double D1 = s - ypa, D2 = ypb - s;
    if (!(D1 * D2 > 0.0)) return false;
    const double a = fabs(D1), b = fabs(D2), hi = a > b ? a : b, lo = a > b ? b : a;
    if (!(hi > LT_BAND_LO * lo && hi < LT_BAND_HI * lo)) return false;
    T = fmax(qdiv(D1, D2), qdiv(D2, D1));
    return T > 2.0;
}

// TSPSI(N=4) + HVAL (or the linint fallback) at T: the water-column profile value of
// WCTS_ITPI (hydro:2619-2644).  One Newton loop per lane runs the solve of the interval that
// holds T and the verdict-only solves of the other intervals.  (Flattening the solves of all 9
// splines of one find_currents was tried and lost: 9 x 7 doubles of state spill.)
LT_DEVN double spline4_eval2(const Knots4& k, double y0, double y1, double y2, double y3, double T, int& nsig)
{
    const double d1 = k.x[1] - k.x[0], d2 = k.x[2] - k.x[1], d3 = k.x[3] - k.x[2];
    const double s1 = (y1 - y0) * k.r1, s2 = (y2 - y1) * k.r2, s3 = (y3 - y2) * k.r3;
    const double p0 = ypc1_end(s1, s1 + d1 * (s1 - s2) * k.r12), p1 = ypc1_mid_r(d1, d2, s1, s2, k.r12);
    const double p2 = ypc1_mid_r(d2, d3, s2, s3, k.r23), p3 = ypc1_end(s3, s3 + d3 * (s3 - s2) * k.r23);
    const int I = (T < k.x[0]) ? 0 : (T > k.x[3]) ? 2 : (T < k.x[2] ? (T < k.x[1] ? 0 : 1) : 2);
    Iv4 v;
    v.X1 = I == 0 ? k.x[0] : I == 1 ? k.x[1] : k.x[2]; v.X2 = I == 0 ? k.x[1] : I == 1 ? k.x[2] : k.x[3];
    v.Y1 = I == 0 ? y0 : I == 1 ? y1 : y2; v.Y2 = I == 0 ? y1 : I == 1 ? y2 : y3;
    v.P1 = I == 0 ? p0 : I == 1 ? p1 : p2; v.P2 = I == 0 ? p1 : I == 1 ? p2 : p3;
    int err = 0, np = 0;
    double sig = 0.0, tq0 = 0.0, tq1 = 0.0, tq2 = 0.0, g0 = 0.0, g1 = 0.0, g2 = 0.0;       // pending (TP1, start) pairs
    bool own = false;                                   // slot 0 is the interval that holds T
    {
        double TP1, SIG0;
        if (!sigs_classify(v.X2 - v.X1, v.Y1, v.Y2, v.P1, v.P2, sig, TP1, SIG0, err)) { tq0 = TP1; g0 = SIG0; np = 1; own = true; }
    }
    {   // the two other intervals: verdict only
        const double sa = I == 0 ? s2 : s1, pa = I == 0 ? p1 : p0, pb = I == 0 ? p2 : p1;
        const double sb = I == 2 ? s2 : s3, pc = I == 2 ? p1 : p2, pd = I == 2 ? p2 : p3;
        double Tq;
        if (sigerr_candidate(sa, pa, pb, Tq)) { if (np == 0) { tq0 = Tq + 1.0; g0 = sig_guess(Tq); } else { tq1 = Tq + 1.0; g1 = sig_guess(Tq); } ++np; }
        if (sigerr_candidate(sb, pc, pd, Tq)) { if (np == 0) { tq0 = Tq + 1.0; g0 = sig_guess(Tq); } else if (np == 1) { tq1 = Tq + 1.0; g1 = sig_guess(Tq); } else { tq2 = Tq + 1.0; g2 = sig_guess(Tq); } ++np; }
    }
    if (np > 0) {
        int cur = 0; NewtonState ns; newton_start(ns, tq0, g0);
        while (cur < np) {
            double o; int e = 0;
            if (newton_step(ns, o, e)) {
                err |= e;
                if (cur == 0 && own) sig = o;
                if (++cur < np) newton_start(ns, cur == 1 ? tq1 : tq2, cur == 1 ? g1 : g2);
            }
        }
    }
    if (err == 0) return hval_interval(T, v.X1, v.X2, v.Y1, v.Y2, v.P1, v.P2, sig);
    ++nsig;
    return linint4(k, y0, y1, y2, y3, T);
}

struct Stage2 { Stencil r, u, v; };

// WCTS_ITPI (hydro:2577-2689) for NF fields sharing one set of knots (u and v share the
// rho-level knots).  v = 0,1,2: value at ix(v+1); v = 3: (b + 4c + f)/6.
#ifndef LT_WCTS_ATTR
#define LT_WCTS_ATTR LT_DEV
#endif
template <class T, int PH, bool W, int NF>
LT_WCTS_ATTR void wcts2(const LtDev& D, const T* const* fld, const Stencil* const* st, const int* grid, int4 und, int L,
                  const ColK& col, int deplvl, double P_zb, double P_zc, double P_zf, int v, double* out, int& nsig)
{
    double vb[NF][4], vc[NF][4], vf[NF][4];
    Knots4 kb, kc, kf;
#pragma unroll
    for (int f = 0; f < NF; ++f) gather4_bcf<T, PH>(D, fld[f], L, deplvl - 1, *st[f], grid[f], und, vb[f], vc[f], vf[f]);
#pragma unroll
    for (int i = 0; i < 4; ++i) zlev3<W>(D, col, deplvl - 1 + i, kb.x[i], kc.x[i], kf.x[i]);
    knots_prepare(kb); knots_prepare(kc);
    const bool first = D.p == 1;                     // (b,b,c): the forward profile is not used
    if (!first) knots_prepare(kf);
    const double* w = v < 3 ? D.LW[v] : D.LW4;
#pragma unroll
    for (int f = 0; f < NF; ++f) {
        double pb = spline4_eval2(kb, vb[f][0], vb[f][1], vb[f][2], vb[f][3], P_zb, nsig);
        double pc = spline4_eval2(kc, vc[f][0], vc[f][1], vc[f][2], vc[f][3], P_zc, nsig);
        double pf = first ? 0.0 : spline4_eval2(kf, vf[f][0], vf[f][1], vf[f][2], vf[f][3], P_zf, nsig);
        out[f] = lag(w, pb, pc, pf);
    }
}

// find_currents (LTRANS.f90:1422-1614)
template <class T, int PH>
LT_DEVN void find_currents2(const LtDev& D, const Stage2& s, const ColK& col, double Zpar,
                            double P_zb, double P_zc, double P_zf, int version,
                            double& Uad, double& Vad, double& Wad, int& nsig)
{
    const int us = D.P.us, ws = D.P.ws;
    const double z0 = D.P.z0;
    double wzb1, wzc1, wzf1; zlev3<true>(D, col, 0, wzb1, wzc1, wzf1);
    if (Zpar < wzb1 + z0 || Zpar < wzc1 + z0 || Zpar < wzf1 + z0) { Uad = 0.0; Vad = 0.0; Wad = 0.0; return; }
    double zb1, zc1, zf1; zlev3<false>(D, col, 0, zb1, zc1, zf1);
    const T* fu = (const T*)D.u; const T* fv = (const T*)D.v; const T* fw = (const T*)D.w;
    const double* lw = D.LW[version - 1];
    if (Zpar < zb1 || Zpar < zc1 || Zpar < zf1) {                      // log layer :1489-1600
        double Ub, Uc, Uf, Vb, Vc, Vf, Wb, Wc, Wf;
        gather_bcf<T, PH>(D, fu, us, 0, s.u, G_U, s.u.nd, Ub, Uc, Uf);
        gather_bcf<T, PH>(D, fv, us, 0, s.v, G_V, s.u.nd, Vb, Vc, Vf);
        gather_bcf<T, PH>(D, fw, ws, 1, s.r, G_RHO, s.u.nd, Wb, Wc, Wf);
        double rz0 = qrcp(z0);
        double num = log10_n((Zpar - wzb1) * rz0);
        double wzb2, wzc2, wzf2; zlev3<true>(D, col, 1, wzb2, wzc2, wzf2);
        double db = num * qrcp(log10_n((zb1 - wzb1) * rz0)), dc = num * qrcp(log10_n((zc1 - wzb1) * rz0)),
               df = num * qrcp(log10_n((zf1 - wzb1) * rz0));
        double wb = num * qrcp(log10_n((wzb2 - wzb1) * rz0)), wc = num * qrcp(log10_n((wzc2 - wzb1) * rz0)),
               wf = num * qrcp(log10_n((wzf2 - wzb1) * rz0));
        Uad = lag(lw, Ub * db, Uc * dc, Uf * df);
        Vad = lag(lw, Vb * db, Vc * dc, Vf * df);
        Wad = lag(lw, Wb * wb, Wc * wc, Wf * wf);
        return;
    }
    int ii = level_window2<false>(D, col, Zpar, us);
    int iii = level_window2<true>(D, col, Zpar, ws);
    LT_ASSERT(ii >= 1 && ii + 3 <= us && iii >= 1 && iii + 3 <= ws);
    {
        const T* f2[2] = {fu, fv}; const Stencil* s2[2] = {&s.u, &s.v}; const int g2[2] = {G_U, G_V};
        double o[2];
        wcts2<T, PH, false, 2>(D, f2, s2, g2, s.u.nd, us, col, ii, P_zb, P_zc, P_zf, version - 1, o, nsig);
        Uad = o[0]; Vad = o[1];
    }
    {
        const T* f1[1] = {fw}; const Stencil* s1[1] = {&s.r}; const int g1[1] = {G_RHO};
        double o[1];
        wcts2<T, PH, true, 1>(D, f1, s1, g1, s.u.nd, ws, col, iii, P_zb, P_zc, P_zf, version - 1, o, nsig);
        Wad = o[0];
    }
}

LT_DEV void load_stencils(const LtDev& D, int re, int ue, int ve, Stage2& st)
{
    st.r.q = D.R.ele + (size_t)(re - 1) * 8; st.r.nd = __ldg(D.R.node + (re - 1));
    st.u.q = D.U.ele + (size_t)(ue - 1) * 8; st.u.nd = __ldg(D.U.node + (ue - 1));
    st.v.q = D.V.ele + (size_t)(ve - 1) * 8; st.v.nd = __ldg(D.V.node + (ve - 1));
}
LT_DEV void stage_weights2(Stage2& s, double xp, double yp)
{
    s.r.xp = s.u.xp = s.v.xp = xp; s.r.yp = s.u.yp = s.v.yp = yp;
    s.r.w = make_weights(s.r.q, xp, yp, false);
    s.u.w = make_weights(s.u.q, xp, yp, false);
    s.v.w = make_weights(s.v.q, xp, yp, false);
}
LT_DEV Rng make_rng(const LtDev& D, int n)
{
    long long gid = D.first_id + D.pid[n];
    Rng g; g.id_lo = (unsigned)((unsigned long long)gid & 0xffffffffull); g.id_hi = (unsigned)((unsigned long long)gid >> 32);
    g.step = D.gstep; g.seed = (unsigned)D.P.seed;
    return g;
}

// ----------------------------------------------------------------- behave ---
// behavior_module.f90:181-551.  Per-particle constants of initBehave (:118-131) are
// uniform in v.2b, so P_swim(n,3) is a pure function of age.
struct BehavOut { double X, Y, Z; bool bott; };
template <class T, int PH>
LT_DEVN BehavOut behave(const LtDev& D, int n, const Stage2& s0, const ColK& col, const Rng& g, double Zpar,
                        double P_zb, double P_zc, double P_zf, double P_zetac, double P_age, double P_depth,
                        double P_U, double P_V, double P_angle)
{
    const ltgpu_params& P = D.P;
    BehavOut o; o.X = 0.0; o.Y = 0.0; o.Z = 0.0; o.bott = false;
    double swim1 = (P.swimfast - P.swimslow) / (P.pediage - P.swimstart);
    double swim2 = P.swimfast - swim1 * P.pediage;
    double swim3 = 0.0;
    if (P_age >= P.swimstart) swim3 = swim1 * P_age + swim2;
    if (P_age >= P.pediage) swim3 = P.swimfast;
    int bh = D.behave[n];
    double timer = 0.0;
    if (bh == 4 || bh == 5) {
        if (P_age >= P.pediage && P_age < P.deadage) bh = 2;
        timer = fmax(0.0, D.timer[n] - (double)P.dt);                   // ledger 15
        D.timer[n] = timer;
        D.behave[n] = (int8_t)bh;
    }
    double P_S = 0.0;
    if (bh == 4 || (bh == 5 && timer == 0.0) || bh == 7) {
        int deplvl = level_window2<false>(D, col, Zpar, P.us);
        const T* f1[1] = {(const T*)D.salt}; const Stencil* s1[1] = {&s0.r}; const int g1[1] = {G_RHO};
        double o1[1];
        int nsig = 0;
        wcts2<T, PH, false, 1>(D, f1, s1, g1, s0.u.nd, P.us, col, deplvl, P_zb, P_zc, P_zf, 3, o1, nsig);
        if (nsig) D.nsig[n] += nsig;
        P_S = o1[0];
    }
    uint4 rnd = philox(g, 0x80000000u);
    unsigned rw[3] = { rnd.x, rnd.y, rnd.z }; int w = 0;
    double parBehav = 0.0, negpos, dev1, devB, sw;
    const double f080 = (double)0.80f, f020 = (double)0.20f;
    auto rand_swim = [&](double swv) {          // dev1 / switch / devB pattern
        negpos = 1.0; dev1 = u_real1(rw[w++]);
        if (dev1 > swv) negpos = -1.0;
        devB = u_real1(rw[w++]);
        parBehav = negpos * devB * swim3;
    };
    if (bh == 1) { if (P_zc < (P_zetac - 1.0)) rand_swim(f080); else rand_swim(0.5); }
    if (bh == 2 || (bh == 5 && timer > 0.0)) { if (P_zc > (P_depth + 1.0)) rand_swim(f020); else rand_swim(0.5); }
    if (bh == 3) {
        double daytime = D.ix[2] / 86400.0;
        double dtime = (daytime - trunc(daytime)) * 24.0, E0 = 0.0;
        if (dtime > P.twistart && dtime < P.twiend) {
            double tst = (dtime - P.twistart) * 3600.0;
            double sn = sin(P.PI * tst / (P.daylength * 3600.0));
            E0 = P.Em * sn * sn;
        }
        double P_light = E0 * exp(P.Kd * P_zc);
        if (P_light < P.thresh) rand_swim(0.5);
        if (P_light > P.thresh) rand_swim(f020);
    }
    if (bh == 4 || (bh == 5 && timer == 0.0)) {
        double sprev = D.sprev[n], zprev = D.zprev[n];
        if (D.it == 1) { sprev = P_S; zprev = P_zc; }
        int btest = 0; double Sslope = 0.0;
        double deltaS = sprev - P_S, deltaz = zprev - P_zc;
        if (D.it > 1) Sslope = deltaS / deltaz;
        if (bh == 4) {
            if (fabs(Sslope) > P.Sgradient) {
                negpos = 1.0; dev1 = u_real1(rw[w++]);
                if (dev1 > f080) negpos = -1.0;
                parBehav = negpos * swim3; btest = 1;
            }
            if (btest == 0) {
                negpos = 1.0; dev1 = u_real1(rw[w++]);
                if (P_age < 1.5 * 24. * 3600.) sw = (double)0.1f;
                else if (P_age < 5. * 24. * 3600.) sw = (double)0.49f;
                else if (P_age < 8. * 24. * 3600.) sw = (double)0.50f;
                else {
                    double ss = ((double)0.50f - (double)0.517f) / (8.0 * 24.0 * 3600.0 - P.pediage);
                    sw = ss * P_age + (double)0.50f - ss * 8.0 * 24.0 * 3600.0;
                    if (P_zc < P_depth + 1.) sw = 0.5;
                }
                if (dev1 > (1 - sw)) negpos = -1.0;
                devB = u_real1(rw[w++]); parBehav = negpos * devB * swim3;
            }
        } else {
            if (fabs(Sslope) > P.Sgradient) {
                negpos = 1.0; dev1 = u_real1(rw[w++]); btest = 1;
                timer = 2.0 * 3600.0;
                if (dev1 > f020) negpos = -1.0;
                parBehav = negpos * swim3;
                if (P_age < 3.5 * 24. * 3600.) { btest = 0; timer = 0.; }
                D.timer[n] = timer;
            }
            if (btest == 0) {
                negpos = 1.0; dev1 = u_real1(rw[w++]); sw = (double)0.495f;
                if (P_age < 1.5 * 24. * 3600.) sw = (double)0.9f;
                if (P_age > 2.0 * 24. * 3600. && P_age < 3.5 * 24. * 3600.) {
                    double ss = ((double)0.3f - (double)0.495f) / (2.0 * 24.0 * 3600.0 - 3.5 * 24.0 * 3600.0);
                    sw = ss * P_age + (double)0.3f - ss * 2.0 * 24.0 * 3600.0;
                }
                if (dev1 > sw) negpos = -1.0;
                devB = u_real1(rw[w++]); parBehav = negpos * devB * swim3;
            }
        }
        D.sprev[n] = P_S; D.zprev[n] = P_zc;
    }
    if (bh == 6) parBehav = (P_age >= P.swimstart) ? P.sink : swim3;
    o.Z = parBehav * P.idt;
    if (bh == 7) {
        double sprev = D.sprev[n];
        if (D.it == 1) { sprev = P_S; D.sprev[n] = P_S; }
        double ca = cos(P_angle), sa = sin(P_angle);
        double X = (P_U * ca - P_V * sa), Y = (P_U * sa + P_V * ca);
        double currentspeed = sqrt(X * X + Y * Y);
        uint8_t fl = D.flags[n];
        if (fl & LT_F_BOTTOM) {
            if (sprev < P_S) { fl &= ~LT_F_BOTTOM; o.Z = P_depth + P.Swimdepth; }
            else o.Z = -9999;
        } else {
            if (currentspeed > (double)0.05f) {
                double Hd = P.Hswimspeed * P.idt;
                double theta = atan(Y / X);
                if (X > 0.0) { o.X = Hd * cos(theta); o.Y = Hd * sin(theta); }
                if (X < 0.0) { o.X = -1.0 * Hd * cos(theta); o.Y = -1.0 * Hd * sin(theta); }
                if (X == 0 && Y >= 0.0) { o.X = 0.0; o.Y = Hd; }
                if (X == 0 && Y <= 0.0) { o.X = 0.0; o.Y = -1.0 * Hd; }
                o.Z = P_depth + P.Swimdepth;
            } else { o.Z = -9999; fl |= LT_F_BOTTOM; }
        }
        D.flags[n] = fl;
        o.bott = (fl & LT_F_BOTTOM) != 0;
    }
    return o;
}

// ------------------------------------------------------------ error sites ---
// The four check sites of update_particles share this (LTRANS.f90:834-879 etc.).
// Returns nothing: the caller `cycle`s afterwards.
LT_DEV void particle_error(const LtDev& D, int n, int code, double revertZ)
{
    int EF = D.P.ErrorFlag;
    int gid = (int)(D.first_id + D.pid[n]);
    if (EF < 1 || EF > 3) atomicMin(D.bad, gid);                         // STOP: lowest id wins
    else if (EF == 1) D.z[n] = revertZ;                                  // pn = p (x, y unchanged)
    else if (EF == 2) D.flags[n] |= LT_F_DEAD;
    else D.flags[n] |= LT_F_OOB;
    int k = atomicAdd(D.nev, 1);
    if (k < D.evcap) { D.ev[k].particle = gid; D.ev[k].code = code; D.ev[k].time = D.ix[2]; }
}


LT_DEV void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
// setEle looks the particle's three elements up one after the other, each with two dependent loads (adjacency row,
// then corner record): what the three lookups will read first is requested together (67 % of k_finish's stall
// samples were long-scoreboard waits, 38 % at these two loads).
LT_DEV void prefetch_elements(const LtDev& D, int re, int ue, int ve)
{
    prefetch_l1(D.R.adj + (size_t)(max(re, 1) - 1) * 10); prefetch_l1(D.R.ele + (size_t)(max(re, 1) - 1) * 8);
    prefetch_l1(D.U.adj + (size_t)(max(ue, 1) - 1) * 10); prefetch_l1(D.U.ele + (size_t)(max(ue, 1) - 1) * 8);
    prefetch_l1(D.V.adj + (size_t)(max(ve, 1) - 1) * 10); prefetch_l1(D.V.ele + (size_t)(max(ve, 1) - 1) * 8);
}

// ============================================================ kernel 1: advect ==
// State carried through the advect kernel.  The kernel body is prologue -> 4 x stage ->
// epilogue with a block barrier before every stage: particles that drop out at a gate keep
// reaching the barriers, and the warps of a block enter each RK stage together, so they fetch
// the same instructions (instruction fetch was k_advect's top stall).
struct AdvS {
    Stage2 st; ColK col;
    double Xpar, Ypar, Zpar, P_zb, P_zc, P_zf, P_depth, P_angle, ca, sa, minpd, maxpd, sU, sV, sW, xs, ys, zs;
    int nsig;                       // SigErr fall-backs of this step's WCTS_ITPI calls
};

template <class T, int PH>
LT_DEV bool advect_prologue(const LtDev& D, int n, AdvS& S)
{
    const ltgpu_params& P = D.P;
    const int idt = P.idt;
    D.s_act[n] = 0;
    if (D.ix[2] <= D.dob[n]) return false;                               // :790-795
    double age = D.age[n] + (double)(float)idt;                          // :798
    D.age[n] = age;
    uint8_t fl = D.flags[n];
    if (age >= P.deadage && P.mortality) {                               // updateStatus behavior:162-179
        if (!(P.settlementon && (fl & LT_F_SETTLED))) { fl |= LT_F_DEAD; D.flags[n] = fl; }
    }
    if (P.settlementon && (fl & LT_F_SETTLED)) return false;             // :804-816
    if (P.mortality && (fl & LT_F_DEAD)) return false;
    if (P.OpenOceanBoundary && (fl & LT_F_OOB)) return false;

    const double Xpar = D.x[n], Ypar = D.y[n], Zold = D.z[n];
    int re = D.r_ele[n], ue = D.u_ele[n], ve = D.v_ele[n];
    prefetch_elements(D, re, ue, ve);
    {                                                                    // setEle :830
        int err = 0, re0 = re, ue0 = ue, ve0 = ve;
        if (!find_element(D.R, Xpar, Ypar, re)) err = 4;
        if (!find_element(D.U, Xpar, Ypar, ue)) err = 5;
        if (!find_element(D.V, Xpar, Ypar, ve)) err = 6;
        if (re != re0) D.r_ele[n] = re;
        if (ue != ue0) D.u_ele[n] = ue;
        if (ve != ve0) D.v_ele[n] = ve;
        if (err) {
            particle_error(D, n, err == 4 ? LTGPU_EV_NOT_IN_RHO : err == 5 ? LTGPU_EV_NOT_IN_U : LTGPU_EV_NOT_IN_V, Zold);
            return false;
        }
    }
    load_stencils(D, re, ue, ve, S.st);
    Stencil s0 = S.st.r; s0.xp = Xpar; s0.yp = Ypar; s0.w = make_weights(S.st.r.q, Xpar, Ypar, true);    // setInterp :882
    const double P_depth = -1.0 * gather_static(D, D.depth, s0);         // :892-896
    const double P_angle = gather_static(D, D.angle, s0);
    double P_zetab, P_zetac, P_zetaf;
    gather_bcf<T, PH>(D, (const T*)D.zeta, 1, 0, s0, G_RHO, s0.nd, P_zetab, P_zetac, P_zetaf);
    double Zp = Zold;
    if (Zp < P_depth) { Zp = P_depth + (double)kF32_1em3; if (P.TrackCollisions) D.hitB[n] += 1; }   // :900-903
    double P_zb = Zp, P_zc = Zp, P_zf = Zp;
    if (Zp > P_zetab) P_zb = P_zetab - (double)kF32_1em3;
    if (Zp > P_zetac) P_zc = P_zetac - (double)kF32_1em3;
    if (Zp > P_zetaf) P_zf = P_zetaf - (double)kF32_1em3;
    const double Zpar = lag(D.LWz, P_zb, P_zc, P_zf);                    // :914 (raw b,c,f triplet: ledger 8)
    ColK& col = S.col; col.zb = P_zetab; col.zc = P_zetac; col.zf = P_zetaf; col.depth = P_depth; col.h = -1.0 * P_depth;
    double a, b, c;
    zlev3<true>(D, col, 0, a, b, c);        S.maxpd = fmax(a, fmax(b, c));                      // :981-987
    zlev3<true>(D, col, P.ws - 1, a, b, c); S.minpd = fmin(a, fmin(b, c));
    S.ca = cos(P_angle); S.sa = sin(P_angle);
    S.Xpar = Xpar; S.Ypar = Ypar; S.Zpar = Zpar; S.P_zb = P_zb; S.P_zc = P_zc; S.P_zf = P_zf;
    S.P_depth = P_depth; S.P_angle = P_angle;
    S.nsig = 0; S.sU = 0.0; S.sV = 0.0; S.sW = 0.0; S.xs = Xpar; S.ys = Ypar; S.zs = Zpar;
    return true;
}


// one RK stage; stage times are (t-h, t, t, t+h) = versions 1,2,2,3 (ledger 6)
template <class T, int PH>
LT_DEV void advect_stage(const LtDev& D, AdvS& S, int stg)
{
    const double eps6 = (double)kF32_1em6;
    double Uad, Vad, Wad;
    // the u / v element records are on their way while the rho weights are worked out (-0.6 .. -3 %); asking for
    // the field windows of the previous stage's levels the same way cost 13 % (they evict the thread-local
    // lines the stage lives on) and a level search started from the previous stage's window changed nothing
    prefetch_l1(S.st.u.q); prefetch_l1(S.st.v.q);
    stage_weights2(S.st, S.xs, S.ys);
    find_currents2<T, PH>(D, S.st, S.col, S.zs, S.P_zb, S.P_zc, S.P_zf, stg == 0 ? 1 : (stg == 3 ? 3 : 2), Uad, Vad, Wad, S.nsig);
    const double wgt = (stg == 0 || stg == 3) ? 1.0 : 2.0;
    S.sU += wgt * Uad; S.sV += wgt * Vad; S.sW += wgt * Wad;
    if (stg < 3) {
        const double f = (double)D.P.idt, ca = S.ca, sa = S.sa;
        if (stg < 2) {
            S.xs = S.Xpar + (Uad * ca - Vad * sa) * f / 2.0; S.ys = S.Ypar + (Uad * sa + Vad * ca) * f / 2.0; S.zs = S.Zpar + Wad * f / 2.0;
        } else {
            S.xs = S.Xpar + (Uad * ca - Vad * sa) * f; S.ys = S.Ypar + (Uad * sa + Vad * ca) * f; S.zs = S.Zpar + Wad * f;
        }
        if (S.zs > S.minpd) S.zs = S.minpd - eps6;
        if (S.zs < S.maxpd) S.zs = S.maxpd + eps6;
    }
}

template <class T, int PH>
LT_DEV void advect_epilogue(const LtDev& D, int n, AdvS& S)
{
    const ltgpu_params& P = D.P;
    const int idt = P.idt;
    const double Xpar = S.Xpar, Ypar = S.Ypar, ca = S.ca, sa = S.sa;
    const double P_U = S.sU / 6.0, P_V = S.sV / 6.0, P_W = S.sW / 6.0;   // :1047-1049
    double newX = Xpar + idt * (P_U * ca - P_V * sa);                    // :1051-1053, :1128-1130
    double newY = Ypar + idt * (P_U * sa + P_V * ca);
    if (P.SaltTempOn) {                                                  // :1062-1076
        stage_weights2(S.st, Xpar, Ypar);
        int deplvl = level_window2<false>(D, S.col, S.Zpar, P.us);
        const T* f2[2] = {(const T*)D.salt, (const T*)D.temp}; const Stencil* s2[2] = {&S.st.r, &S.st.r}; const int g2[2] = {G_RHO, G_RHO};
        double o[2];
        wcts2<T, PH, false, 2>(D, f2, s2, g2, S.st.u.nd, P.us, S.col, deplvl, S.P_zb, S.P_zc, S.P_zf, 3, o, S.nsig);
        D.psalt[n] = o[0]; D.ptemp[n] = o[1];
    }
    if (P.HTurbOn) {                                                     // hor_turb_module.f90:29-50
        Rng g = make_rng(D, n);
        uint4 r = philox(g, 0u);
        double sd = sqrt(2.0 * P.ConstantHTurb * idt);
        newX = (Xpar + idt * (P_U * ca - P_V * sa)) + box_muller_n(D, r.x, r.y) * sd;
        newY = (Ypar + idt * (P_U * sa + P_V * ca)) + box_muller_n(D, r.z, r.w) * sd;
    }
    D.s_depth[n] = S.P_depth; D.s_angle[n] = S.P_angle;
    D.s_zeb[n] = S.col.zb; D.s_zec[n] = S.col.zc; D.s_zef[n] = S.col.zf;
    D.s_pzb[n] = S.P_zb; D.s_pzc[n] = S.P_zc; D.s_pzf[n] = S.P_zf; D.s_zpar[n] = S.Zpar;
    D.s_nx[n] = newX; D.s_ny[n] = newY; D.s_advz[n] = idt * P_W;
    D.s_pu[n] = P_U; D.s_pv[n] = P_V; D.s_turbv[n] = 0.0;
    D.s_act[n] = 1;
    if (S.nsig) D.nsig[n] += S.nsig;
}

// ============================================================= kernel 2: VTurb ==
// ver_turb_module.f90:30-380.  Everything a warp does here is arranged to be CONVERGENT:
//   1. the KH profile (ws levels x 3 times) and the w-level depths are gathered once into
//      thread-local arrays (21 iterations, identical for every lane);
//   2. knot values fy(k), knot slopes yp(k) (YPC1) and tension factors sg(k) (SIGS) are built
//      for a window of VW knots around the particle in lock-step loops;
//   3. the 60 random-displacement sub-steps then only look knots up.
// A particle that walks out of its window rebuilds it around the new position (rare: the
// window spans ~ +-15 knots = +-18% of the water column, the RDM step is ~1%).
#define VW 32                      // knots held per window

struct VtCtx {
    const LtDev& D;
    int ws, p2;
    double khp[3][LT_MAXLEV];             // KH at the particle, per hydro time and w-level (:102-108)
    double zl[3][LT_MAXLEV];              // w-level depths at the particle
    double hs[3];                         // newx spacing per hydro time (:120-124)
    double Z1, ZN, H, rH;                 // time-combined knot line: x(k) = Z1 + (k - 0.5) H
    double fy[VW], yp[VW], sg[VW];        // knot value, YPC1 slope, SIGS tension of interval (k, k+1)
    double tp[VW]; unsigned char pend[VW]; // pending convexity solves of build()
    int ka, kb, ia, ib;                   // knots [ka, kb] held; intervals [ia, ib] fully defined
    bool sigerr;
    LT_DEV VtCtx(const LtDev& D_) : D(D_) {}

    LT_DEV double knot_x(int k) const { double x = fma((double)k - 0.5, H, Z1); x = k <= 1 ? Z1 : x; return k >= p2 ? ZN : x; }
    double z1[3];
    LT_DEV double newx(int t, int j) const { return z1[t] + (double)(j - 4) * hs[t]; }
    // piecewise-linear KH profile at newx(j): smallest jlo >= 1 with wz(jlo+1) > x (:135-166), pads
    // (:169-177).  The current segment lives in registers (slope, intercept, upper end); the
    // thread-local profile arrays are only read when a pointer moves to the next level.
    struct Seg { double slope, icpt, znext; int lev; };
    LT_DEV void seg_load(Seg& g, int t, int lev) const
    {   // lev = jlo (1-based)
        LT_ASSERT(lev >= 1 && lev <= ws - 1);
        double zlo = zl[t][lev - 1], zhi = zl[t][lev], klo = khp[t][lev - 1], khi = khp[t][lev];
        g.lev = lev; g.slope = qdiv(klo - khi, zlo - zhi); g.icpt = klo - g.slope * zlo; g.znext = zhi;   // :126-133
    }
    LT_DEV double newy(int t, int j, Seg& g) const
    {
        if (j <= 4) return khp[0][0];                                  // ledger 11: KHb(1) for all three times
        if (j >= p2 + 4) return khp[t][ws - 1];
        double x = newx(t, j);
        while (!(g.znext > x) && g.lev < ws - 1) seg_load(g, t, g.lev + 1);
        return g.slope * x + g.icpt;
    }
    // knot values, slopes and tension factors for knots [ka_, ka_ + VW - 1] /\ [1, p2]
    LT_DEVN void build(int ka_)
    {
        ka = ka_; kb = min(p2, ka + VW - 1);
        LT_ASSERT(ka >= 1 && kb - ka >= 3 && kb - ka < VW);
#ifdef LT_DEBUG_TRACE
        atomicAdd(&g_dbgcnt2[0], 1ull);                  // builds
#endif
        const int k0 = max(ka, 2), k1 = min(kb, p2 - 1);
        double S[3] = {0.0, 0.0, 0.0}; Seg hi[3], lo[3];
#pragma unroll
        for (int t = 0; t < 3; ++t) { seg_load(hi[t], t, 1); lo[t] = hi[t]; }
        if (k0 <= k1) {
#pragma unroll
            for (int t = 0; t < 3; ++t) {
                double acc = 0.0;
                for (int j = k0; j <= k0 + 7; ++j) acc += newy(t, j, hi[t]);       // :184-194
                S[t] = acc;
                double d = newy(t, k0, lo[t]); (void)d;                             // position the trailing pointer
            }
        }
        for (int k = ka; k <= kb; ++k) {
            double my[3];
            if (k == 1) { my[0] = khp[0][0]; my[1] = khp[1][0]; my[2] = khp[2][0]; }                      // :197-210
            else if (k == p2) { my[0] = khp[0][ws - 1]; my[1] = khp[1][ws - 1]; my[2] = khp[2][ws - 1]; }
            else {
#pragma unroll
                for (int t = 0; t < 3; ++t) {
                    if (k > k0) S[t] += newy(t, k + 7, hi[t]) - newy(t, k - 1, lo[t]);   // running 8-point sum
                    my[t] = S[t] / 8.0;
                }
            }
            double fb = lag(D.LW[0], my[0], my[1], my[2]), fc = lag(D.LW[1], my[0], my[1], my[2]), ff = lag(D.LW[2], my[0], my[1], my[2]);
            fb = fb < 0.0 ? 0.0 : fb; fc = fc < 0.0 ? 0.0 : fc; ff = ff < 0.0 ? 0.0 : ff;     // :264-268
            fy[k - ka] = (fb + 4.0 * fc + ff) / 6.0;                                          // :272-275
        }
        // YPC1 (tension:852-978): needs both neighbours, or the profile end
        const int pa = ka == 1 ? 1 : ka + 1, pb = kb == p2 ? p2 : kb - 1;
        if (pa == 1) {
            double d1 = knot_x(2) - knot_x(1), d2 = knot_x(3) - knot_x(2);
            double s1 = qdiv(fy[2 - ka] - fy[1 - ka], d1), s2 = qdiv(fy[3 - ka] - fy[2 - ka], d2);
            yp[1 - ka] = ypc1_end(s1, s1 + qdiv(d1 * (s1 - s2), d1 + d2));
        }
        {   // interior knots: abscissae, values and the left chord slope slide along in registers
            const int m0 = max(pa, 2), m1 = min(pb, p2 - 1);
            double x0 = knot_x(m0 - 1), x1 = knot_x(m0), f0 = fy[m0 - 1 - ka], f1 = fy[m0 - ka];
            double sl = qdiv(f1 - f0, x1 - x0);
            for (int k = m0; k <= m1; ++k) {
                const double x2 = knot_x(k + 1), f2 = fy[k + 1 - ka];
                const double d1 = x1 - x0, d2 = x2 - x1, sr = qdiv(f2 - f1, d2);
                yp[k - ka] = ypc1_mid(d1, d2, sl, sr);
                x0 = x1; x1 = x2; f0 = f1; f1 = f2; sl = sr;
            }
        }
        if (pb == p2) {
            double d1 = knot_x(p2 - 1) - knot_x(p2 - 2), d2 = knot_x(p2) - knot_x(p2 - 1);
            double s1 = qdiv(fy[p2 - 1 - ka] - fy[p2 - 2 - ka], d1), s2 = qdiv(fy[p2 - ka] - fy[p2 - 1 - ka], d2);
            yp[p2 - ka] = ypc1_end(s2, s2 + qdiv(d2 * (s2 - s1), d1 + d2));
        }
        // SIGS (tension:314-782) for every interval with both slopes known
        ia = pa; ib = pb - 1;
        int np = 0;
        for (int k = ia; k <= ib; ++k) {
            double sigma, TP1, SIG0; int e = 0;
            if (sigs_classify(knot_x(k + 1) - knot_x(k), fy[k - ka], fy[k + 1 - ka], yp[k - ka], yp[k + 1 - ka], sigma, TP1, SIG0, e))
                sg[k - ka] = sigma;
            else { LT_ASSERT(np < VW && k - ka >= 0 && k - ka < VW); sg[k - ka] = SIG0; tp[np] = TP1; pend[np] = (unsigned char)(k - ka); ++np; }
            if (e) sigerr = true;
        }
        // one Newton loop per lane over all its pending intervals (see wcts2)
        int cur = 0; NewtonState ns;
        if (np > 0) newton_start(ns, tp[0], sg[pend[0]]);
#ifdef LT_DEBUG_TRACE
        atomicAdd(&g_dbgcnt2[1], (unsigned long long)np);            // convexity solves
        atomicAdd(&g_dbgcnt2[3], (unsigned long long)(ib - ia + 1)); // intervals classified
#endif
        while (cur < np) {
            double o; int e = 0;
#ifdef LT_DEBUG_TRACE
            atomicAdd(&g_dbgcnt2[2], 1ull);              // Newton iterations
#endif
            if (newton_step(ns, o, e)) {
                sg[pend[cur]] = o; if (e) sigerr = true;
                if (++cur < np) newton_start(ns, tp[cur], sg[pend[cur]]);
            }
        }
    }
    // Round-1 fused kernel only (unless ltgpu_params.vturb_window_sigs): the reference's SIGS sweeps all p2 - 1 intervals of
    // the fit and any SigErr sends the whole particle-step to linint (ver_turb:278-279, 300-336),
    // the window above only sees its own intervals.  This pass streams over every knot (values,
    // YPC1 slopes, T), keeps nothing, and runs the Newton loop for the intervals whose T lies in
    // the band where it can fail (LT_BAND_*), so that `sigerr` is the reference's verdict.
    LT_DEVN void sweep()
    {
        double S[3] = {0.0, 0.0, 0.0}; Seg hi[3], lo[3];
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            seg_load(hi[t], t, 1); lo[t] = hi[t];
            double acc = 0.0;
            for (int j = 2; j <= 9; ++j) acc += newy(t, j, hi[t]);
            S[t] = acc;
            double d = newy(t, 2, lo[t]); (void)d;
        }
        double fm2 = 0.0, fm1 = 0.0, f0 = 0.0, ypm2 = 0.0;
        int np = 0;
        auto flush = [&]() {
            for (int c = 0; c < np; ++c) {
                NewtonState ns; newton_start(ns, tp[c] + 1.0, sig_guess(tp[c]));
                double o; int e = 0;
                while (!newton_step(ns, o, e)) {}
                if (e) sigerr = true;
            }
            np = 0;
        };
        for (int k = 1; k <= p2; ++k) {
            double my[3];
            if (k == 1) { my[0] = khp[0][0]; my[1] = khp[1][0]; my[2] = khp[2][0]; }
            else if (k == p2) { my[0] = khp[0][ws - 1]; my[1] = khp[1][ws - 1]; my[2] = khp[2][ws - 1]; }
            else {
#pragma unroll
                for (int t = 0; t < 3; ++t) {
                    if (k > 2) S[t] += newy(t, k + 7, hi[t]) - newy(t, k - 1, lo[t]);
                    my[t] = S[t] / 8.0;
                }
            }
            double fb = lag(D.LW[0], my[0], my[1], my[2]), fc = lag(D.LW[1], my[0], my[1], my[2]), ff = lag(D.LW[2], my[0], my[1], my[2]);
            fb = fb < 0.0 ? 0.0 : fb; fc = fc < 0.0 ? 0.0 : fc; ff = ff < 0.0 ? 0.0 : ff;
            fm2 = fm1; fm1 = f0; f0 = (fb + 4.0 * fc + ff) / 6.0;          // fy(k-2), fy(k-1), fy(k)
            if (k < 3) continue;
            const double d1 = knot_x(k - 1) - knot_x(k - 2), d2 = knot_x(k) - knot_x(k - 1);
            const double s1 = qdiv(fm1 - fm2, d1), s2 = qdiv(f0 - fm1, d2);
            const double ypk1 = ypc1_mid(d1, d2, s1, s2);                  // slope at knot k-1
            if (k == 3) ypm2 = ypc1_end(s1, s1 + qdiv(d1 * (s1 - s2), d1 + d2));       // slope at knot 1
            double Tq;
            if (sigerr_candidate(s1, ypm2, ypk1, Tq)) { tp[np++] = Tq; if (np == VW) flush(); }   // interval (k-2, k-1)
            ypm2 = ypk1;
            if (k == p2) {
                const double ypN = ypc1_end(s2, s2 + qdiv(d2 * (s2 - s1), d1 + d2));
                if (sigerr_candidate(s2, ypk1, ypN, Tq)) { tp[np++] = Tq; if (np == VW) flush(); }   // interval (p2-1, p2)
            }
        }
        flush();
    }
    // HVAL / HPVAL interval choice incl. INTRVL (tension:1026-1041, 1287-1354)
    LT_DEV int interval(double Tq) const
    {
        if (Tq < Z1) return 1;
        if (Tq > ZN) return p2 - 1;
        int k = (int)floor((Tq - Z1) * rH + 0.5);
        k = max(1, min(p2 - 1, k));
        while (k > 1 && Tq < knot_x(k)) --k;
        while (k < p2 - 1 && !(Tq < knot_x(k + 1))) ++k;
        return k;
    }
    // the same choice, also returning the interval's end knots (the walk needs them anyway, so
    // the common case costs two knot_x instead of four)
    LT_DEV int interval_x(double Tq, double& X1, double& X2) const
    {
        int k;
        if (Tq < Z1) k = 1;
        else if (Tq > ZN) k = p2 - 1;
        else {
            k = (int)floor((Tq - Z1) * rH + 0.5);
            k = max(1, min(p2 - 1, k));
            X1 = knot_x(k); X2 = knot_x(k + 1);
            while (k > 1 && Tq < X1) { --k; X2 = X1; X1 = knot_x(k); }
            while (k < p2 - 1 && !(Tq < X2)) { ++k; X1 = X2; X2 = knot_x(k + 1); }
            return k;
        }
        X1 = knot_x(k); X2 = knot_x(k + 1);
        return k;
    }
    LT_DEV void need(int I) { if (I < ia || I > ib) build(max(1, min(I - VW / 2 + 1, p2 - VW + 1))); }
};

template <class T, int PH>
LT_DEV void vturb_particle(const LtDev& D, int n)
{
    if (!D.s_act[n]) return;
    const double background = (double)1.0E-6f;                          // ledger 2
    const double Xpar = D.x[n], Ypar = D.y[n];
    const int re = D.r_ele[n];
    Stencil s0; s0.q = D.R.ele + (size_t)(re - 1) * 8; s0.nd = __ldg(D.R.node + (re - 1));
    s0.xp = Xpar; s0.yp = Ypar; s0.w = make_weights(s0.q, Xpar, Ypar, true);      // getInterp uses setInterp's weights
    ColK col; col.zb = D.s_zeb[n]; col.zc = D.s_zec[n]; col.zf = D.s_zef[n]; col.depth = D.s_depth[n]; col.h = -1.0 * col.depth;
    const double P_zc = D.s_pzc[n], P_depth = col.depth, P_zetac = col.zc;
    VtCtx V(D);
    V.ws = D.P.ws; V.p2 = 4 * V.ws; V.sigerr = false;
    const T* fk = (const T*)D.kh;
#pragma unroll 1
    for (int l = 0; l < V.ws; ++l) {
        gather_bcf_inl<T, PH>(D, fk, V.ws, l, s0, G_RHO, s0.nd, V.khp[0][l], V.khp[1][l], V.khp[2][l]);
        zlev3<true>(D, col, l, V.zl[0][l], V.zl[1][l], V.zl[2][l]);
    }
    const double rp2 = 1.0 / (double)V.p2;
#pragma unroll
    for (int t = 0; t < 3; ++t) { V.z1[t] = V.zl[t][0]; V.hs[t] = (V.zl[t][V.ws - 1] - V.zl[t][0]) * rp2; }
    V.Z1 = lag(D.LW4, V.zl[0][0], V.zl[1][0], V.zl[2][0]);
    V.ZN = lag(D.LW4, V.zl[0][V.ws - 1], V.zl[1][V.ws - 1], V.zl[2][V.ws - 1]);
    V.H = (V.ZN - V.Z1) * rp2; V.rH = qrcp(V.H);
    V.ka = 1; V.kb = 0; V.ia = 1; V.ib = 0;
    if (!D.P.vturb_window_sigs) V.sweep();
    const Rng g = make_rng(D, n);
    const double deltat = 2.0;
    const int loop = D.P.idt / 2;                                       // :282-283
    double ParZc = P_zc;
    uint4 rnd = make_uint4(0, 0, 0, 0);
    // the current interval's seven numbers live in registers; the thread-local knot arrays
    // are only touched when the particle changes interval (long-scoreboard stalls on those
    // arrays were k_vturb's top stall)
    int cI = -1; double cX1 = 0, cX2 = 0, cY1 = 0, cY2 = 0, cP1 = 0, cP2 = 0, cSG = 0;
#ifdef LT_DEBUG_TRACE
    int dI0 = -1, dImin = 1 << 30, dImax = -1;
#endif
    auto load_iv = [&](double zq) {
        if (cI >= 0 && zq >= cX1 && zq < cX2) return;                  // still inside [X(I), X(I+1)): INTRVL gives I
        int I = V.interval_x(zq, cX1, cX2); V.need(I);
        int q = I - V.ka;
        LT_ASSERT(I >= V.ia && I <= V.ib && q >= 0 && q + 1 < VW && I >= 1 && I <= V.p2 - 1);
        cI = I;
#ifdef LT_DEBUG_TRACE
        if (dI0 < 0) dI0 = I; dImin = min(dImin, I); dImax = max(dImax, I);
#endif
        cY1 = V.fy[q]; cY2 = V.fy[q + 1]; cP1 = V.yp[q]; cP2 = V.yp[q + 1]; cSG = V.sg[q];
    };
#pragma unroll 1
    for (int i = 0; i < loop; ++i) {                                    // :291-337
        double Kprimec = 0.0;
        if (!(ParZc < P_depth || ParZc > P_zetac)) {
            load_iv(ParZc);
            if (!V.sigerr) Kprimec = hpval_interval(ParZc, cX1, cX2, cY1, cY2, cP1, cP2, cSG);
            else Kprimec = qdiv(cY1 - cY2, cX1 - cX2);                  // linint slope
        }
        const double KprimeZc = -1.0 * Kprimec * deltat;
        const double Z3rdc = ParZc + 0.5 * KprimeZc;
        double KH3rdc;
        if (Z3rdc < P_depth || Z3rdc > P_zetac) KH3rdc = background;
        else {
            load_iv(Z3rdc);
            if (!V.sigerr) KH3rdc = hval_interval(Z3rdc, cX1, cX2, cY1, cY2, cP1, cP2, cSG);
            else { double m = qdiv(cY1 - cY2, cX1 - cX2); KH3rdc = m * Z3rdc + (cY1 - m * cX1); }
            if (KH3rdc < background) KH3rdc = background;
        }
        if ((i & 1) == 0) rnd = philox(g, 1u + (unsigned)(i >> 1));
        const double DEV = (i & 1) ? box_muller(D, rnd.z, rnd.w) : box_muller(D, rnd.x, rnd.y);
        ParZc = ParZc + KprimeZc + DEV * sqrt(2.0 * KH3rdc * deltat);  // (...)**0.5, ledger 12
#ifdef LT_DEBUG_TRACE
        if (D.first_id + D.pid[n] == D.dbg_id) { D.dbg[4 * i] = ParZc; D.dbg[4 * i + 1] = Kprimec; D.dbg[4 * i + 2] = KH3rdc; D.dbg[4 * i + 3] = DEV; }
#endif
    }
#ifdef LT_DEBUG_TRACE
    if (dI0 >= 0) { atomicAdd(&g_dbghist[min(dImax - dI0, 31)], 1ull); atomicAdd(&g_dbghist[32 + min(dI0 - dImin, 31)], 1ull); }
#endif
    if (V.sigerr) D.nsig[n] += 1;
    D.s_turbv[n] = P_zc - ParZc;                                        // :342
}

// ============================================================ kernel 3: finish ==
// intersect_reflect through the segment buckets: candidates are the segments whose
// bounding box meets a bucket touched by the move's bounding box -- a superset of what
// passes the reference's reject test (boundary:1677-1680) -- tested with the reference's
// formulas; nearest hit, lowest index on ties (= the ascending strict-< scan of :1887-1897).
LT_DEVN bool intersect_segment(const LtDev& D, int i, double Xpos, double Ypos, double nXpos, double nYpos,
                               double xlow, double xhigh, double ylow, double yhigh, double& dtest, Hit& h)
{
    const double2* sp = reinterpret_cast<const double2*>(D.seg + i);
    double2 s1_ = __ldg(sp), s2_ = __ldg(sp + 1);
    double bcx1 = s1_.x, bcy1 = s1_.y, bcx2 = s2_.x, bcy2 = s2_.y;
    if ((bcx1 > xhigh && bcx2 > xhigh) || (bcx1 < xlow && bcx2 < xlow) ||
        (bcy1 > yhigh && bcy2 > yhigh) || (bcy1 < ylow && bcy2 < ylow)) return false;
    double bxhigh = fmax(bcx1, bcx2), bxlow = fmin(bcx1, bcx2), byhigh = fmax(bcy1, bcy2), bylow = fmin(bcy1, bcy2);
    double ix, iy, rx1, ry1, rx2, ry2, Mbc = 0.0, dPBC = 0.0;
    int kind;
    if (bcx1 == bcx2 || nXpos == Xpos) {
        if (bcx1 == bcx2 && nXpos == Xpos) return false;
        if (bcx1 == bcx2 && nYpos == Ypos) {
            ix = bcx1; iy = nYpos; kind = 1;
            dPBC = sqrt((ix - nXpos) * (ix - nXpos) + (iy - nYpos) * (iy - nYpos));
        } else if (nXpos == Xpos && bcy1 == bcy2) {
            ix = nXpos; iy = bcy1; kind = 2;
            dPBC = sqrt((ix - nXpos) * (ix - nXpos) + (iy - nYpos) * (iy - nYpos));
        } else if (bcx1 == bcx2 && nYpos != Ypos) {
            double Mp = (nYpos - Ypos) / (nXpos - Xpos), Bp = Ypos - Mp * Xpos;
            ix = bcx1; iy = Mp * ix + Bp; kind = 1; dPBC = nXpos - ix;
        } else if (nXpos == Xpos && bcy1 != bcy2) {
            Mbc = (bcy2 - bcy1) / (bcx2 - bcx1);
            double Bbc = bcy2 - Mbc * bcx2;
            ix = nXpos; iy = Mbc * ix + Bbc; kind = 3;
        } else return false;
    } else {
        Mbc = (bcy2 - bcy1) / (bcx2 - bcx1);
        double Bbc = bcy2 - Mbc * bcx2;
        double Mp = (nYpos - Ypos) / (nXpos - Xpos), Bp = Ypos - Mp * Xpos;
        ix = (Bbc - Bp) / (Mp - Mbc);
        iy = Mp * ix + Bp;
        if (Mbc == 0.0) { iy = byhigh; kind = 2; dPBC = nYpos - bcy1; } else kind = 3;
    }
    if (!(ix <= xhigh && ix >= xlow && iy <= yhigh && iy >= ylow &&
          ix <= bxhigh && ix >= bxlow && iy <= byhigh && iy >= bylow)) return false;
    if (kind == 1) { rx1 = nXpos + (2.0 * dPBC); ry1 = nYpos; rx2 = nXpos - (2.0 * dPBC); ry2 = nYpos; }
    else if (kind == 2) { rx1 = nXpos; ry1 = nYpos + (2.0 * dPBC); rx2 = nXpos; ry2 = nYpos - (2.0 * dPBC); }
    else {
        double distBC = sqrt((bcx1 - bcx2) * (bcx1 - bcx2) + (bcy1 - bcy2) * (bcy1 - bcy2));
        double crossk = ((nXpos - bcx1) * (bcy2 - bcy1)) - ((bcx2 - bcx1) * (nYpos - bcy1));
        dPBC = sqrt(crossk * crossk) / distBC;
        double mP = -1.0 / Mbc, bP = nYpos - mP * nXpos;
        double rr = sqrt(((2.0 * dPBC) * (2.0 * dPBC)) / (1.0 + mP * mP));
        rx1 = rr + nXpos; ry1 = mP * rx1 + bP;
        rx2 = rr * -1.0 + nXpos; ry2 = mP * rx2 + bP;
    }
    double dist1 = sqrt((ix - rx1) * (ix - rx1) + (iy - ry1) * (iy - ry1));
    double dist2 = sqrt((ix - rx2) * (ix - rx2) + (iy - ry2) * (iy - ry2));
    double d = sqrt((Xpos - ix) * (Xpos - ix) + (Ypos - iy) * (Ypos - iy));
    if (d < dtest || (d == dtest && d < 999999. && i < h.seg)) {
        // ledger 19 (dist1 == dist2 keeps a stale value in the reference): only when the end
        // point lies on the boundary line; the end point itself is then its own mirror image
        if (dist1 < dist2) { h.rx = rx1; h.ry = ry1; } else if (dist1 > dist2) { h.rx = rx2; h.ry = ry2; } else { h.rx = rx1; h.ry = ry1; }
        h.ix = ix; h.iy = iy; h.seg = i; h.water = !__ldg(D.land + i);
        dtest = d;
        return true;
    }
    return false;
}

LT_DEV bool intersect_reflect2(const LtDev& D, double Xpos, double Ypos, double nXpos, double nYpos, int skipbound, Hit& h)
{
    double xhigh = fmax(Xpos, nXpos), xlow = fmin(Xpos, nXpos), yhigh = fmax(Ypos, nYpos), ylow = fmin(Ypos, nYpos);
    double dtest = 999999.;
    bool found = false;
    h.seg = 0x7fffffff;
    int cx0 = (int)floor((xlow - D.sg_x0) * D.sg_rcs), cx1 = (int)floor((xhigh - D.sg_x0) * D.sg_rcs);
    int cy0 = (int)floor((ylow - D.sg_y0) * D.sg_rcs), cy1 = (int)floor((yhigh - D.sg_y0) * D.sg_rcs);
    if (cx1 < 0 || cy1 < 0 || cx0 >= D.sg_nx || cy0 >= D.sg_ny) return false;     // no segment bbox out there
    cx0 = max(cx0, 0); cy0 = max(cy0, 0); cx1 = min(cx1, D.sg_nx - 1); cy1 = min(cy1, D.sg_ny - 1);
    if ((long long)(cx1 - cx0 + 1) * (cy1 - cy0 + 1) > 64) {             // very long move: scan everything
        for (int i = 0; i < D.nbounds; ++i)
            if (i != skipbound) found |= intersect_segment(D, i, Xpos, Ypos, nXpos, nYpos, xlow, xhigh, ylow, yhigh, dtest, h);
        return found;
    }
    for (int cy = cy0; cy <= cy1; ++cy)
        for (int cx = cx0; cx <= cx1; ++cx) {
            int c = cy * D.sg_nx + cx;
            for (int q = __ldg(D.sg_ptr + c); q < __ldg(D.sg_ptr + c + 1); ++q) {
                int i = __ldg(D.sg_idx + q);
                if (i != skipbound) found |= intersect_segment(D, i, Xpos, Ypos, nXpos, nYpos, xlow, xhigh, ylow, yhigh, dtest, h);
            }
        }
    return found;
}

// inpoly through y-bands (point_in_polygon_module.f90:25-167): only edges whose y-range
// contains y can (a) carry a vertex on the ray, (b) be the vertex the point sits on,
// (c) be crossed by the ray.  If a vertex lies on the ray the reference's vertex walk
// (:60-117) depends on polygon order, so that (measure-zero) case runs the full routine.
// `poly` holds several closed polygons back to back; idx lists edge start vertices i
// (edge i -> i+1); edges joining two polygons are never listed.
LT_DEV int inpoly_banded(double x, double y, const double2* __restrict__ poly, double y0, double rbh, int nband,
                         const int* __restrict__ ptr, const int* __restrict__ idx)
{   // returns 0 out, 1 in, 2 = needs the full routine
    int b = (int)floor((y - y0) * rbh);
    if (b < 0 || b >= nband) return 0;
    int crossed = 0;
    for (int q = __ldg(ptr + b); q < __ldg(ptr + b + 1); ++q) {
        int i = __ldg(idx + q);
        double2 a = __ldg(poly + i), c = __ldg(poly + i + 1);
        if ((a.y == y && a.x > x) || (c.y == y && c.x > x)) return 2;
        if ((a.x == x && a.y == y) || (c.x == x && c.y == y)) return 1;
        if ((a.x <= x && c.x <= x) || (a.y <= y && c.y <= y) || (a.y >= y && c.y >= y)) continue;
        if (a.x > x && c.x > x) { crossed++; continue; }
        double m = (c.y - a.y) / (c.x - a.x);
        double bb = a.y - m * a.x;
        double ix = (y - bb) / m;
        if (ix == x) return 1;
        if (ix > x) crossed++;
    }
    return crossed & 1;
}

template <class T, int PH>
LT_DEV void finish_particle(const LtDev& D, int n)
{
    if (!D.s_act[n]) return;
    const ltgpu_params& P = D.P;
    const double Xpar = D.x[n], Ypar = D.y[n];
    const double P_depth = D.s_depth[n], P_zetac = D.s_zec[n], Zpar = D.s_zpar[n];
    const double eps6 = (double)kF32_1em6;
    const double age = D.age[n];
    int re = D.r_ele[n], ue = D.u_ele[n], ve = D.v_ele[n];
    prefetch_elements(D, re, ue, ve);                                    // for the setEle at the end
    BehavOut bo; bo.X = bo.Y = bo.Z = 0.0; bo.bott = false;
    if (P.Behavior != 0) {                                               // :1110
        Stage2 st; ColK col;
        st.r.q = D.R.ele + (size_t)(re - 1) * 8; st.r.nd = __ldg(D.R.node + (re - 1));
        st.u.nd = __ldg(D.U.node + (ue - 1));
        st.r.xp = Xpar; st.r.yp = Ypar; st.r.w = make_weights(st.r.q, Xpar, Ypar, false);
        col.zb = D.s_zeb[n]; col.zc = P_zetac; col.zf = D.s_zef[n]; col.depth = P_depth; col.h = -1.0 * P_depth;
        bo = behave<T, PH>(D, n, st, col, make_rng(D, n), Zpar, D.s_pzb[n], D.s_pzc[n], D.s_pzf[n], P_zetac, age, P_depth,
                       D.s_pu[n], D.s_pv[n], D.s_angle[n]);
    }
    double newXpos = D.s_nx[n], newYpos = D.s_ny[n];
    double newZpos = Zpar + D.s_advz[n] + D.s_turbv[n];                  // :1130
    int hitB = 0;
    if (newZpos > P_zetac) { double r = P_zetac - newZpos; newZpos = P_zetac + r; }
    if (newZpos < P_depth) { double r = P_depth - newZpos; newZpos = P_depth + r; hitB++; }
    newZpos = newZpos + bo.Z;
    if (P.Behavior == 7) {
        if (bo.bott) { newXpos = Xpar; newYpos = Ypar; newZpos = P_depth; }
        else { newXpos = newXpos + bo.X; newYpos = newYpos + bo.Y; newZpos = P_depth + P.Swimdepth; }
    }
    if (newZpos > P_zetac) newZpos = P_zetac - eps6;
    if (newZpos < P_depth) { newZpos = P_depth + eps6; hitB++; }
    if (P.TrackCollisions && hitB) D.hitB[n] += hitB;

    double Xpos = Xpar, Ypos = Ypar, nXpos = newXpos, nYpos = newYpos;  // :1180-1232
    int skip = -1, reflects = 0, hitL = 0;
    for (;;) {
        Hit h;
        if (!intersect_reflect2(D, Xpos, Ypos, nXpos, nYpos, skip, h)) break;
        skip = h.seg;
        hitL++;
        if (P.OpenOceanBoundary && h.water) {
            D.x[n] = h.ix; D.y[n] = h.iy; D.z[n] = newZpos;
            D.flags[n] |= LT_F_OOB;
            if (P.TrackCollisions) D.hitL[n] += hitL;
            return;
        }
        if (++reflects > 3) {
            if (P.TrackCollisions) D.hitL[n] += hitL;
            particle_error(D, n, LTGPU_EV_OUT_3RD, Zpar);
            return;
        }
        Xpos = h.ix; Ypos = h.iy; nXpos = h.rx; nYpos = h.ry;
    }
    if (P.TrackCollisions && hitL) D.hitL[n] += hitL;
    newXpos = nXpos; newYpos = nYpos;
    {                                                                    // mbounds :1240
        int in = inpoly_banded(newXpos, newYpos, D.bxy, D.mb_y0, D.mb_rbh, D.mb_n, D.mb_ptr, D.mb_idx);
        if (in == 2) in = inpoly(newXpos, newYpos, D.maxbound, D.bxy, false) ? 1 : 0;
        if (!in) { particle_error(D, n, LTGPU_EV_OUT_MAIN, Zpar); return; }
    }
    if (D.maxisland > 0) {                                               // ibounds :1275
        int in = D.ib_ok ? inpoly_banded(newXpos, newYpos, D.hxy, D.ib_y0, D.ib_rbh, D.ib_n, D.ib_ptr, D.ib_idx) : 2;
        if (in == 2) in = in_any_island(D, newXpos, newYpos) ? 1 : 0;
        if (in) { particle_error(D, n, LTGPU_EV_IN_ISLAND, Zpar); return; }
    }
    D.x[n] = newXpos; D.y[n] = newYpos; D.z[n] = newZpos;                // commit :1312-1314, :1407-1414
    {                                                                    // setEle at the new position :1317
        int err = 0, re0 = re, ue0 = ue, ve0 = ve;
        if (!find_element(D.R, newXpos, newYpos, re)) err = 4;
        if (!find_element(D.U, newXpos, newYpos, ue)) err = 5;
        if (!find_element(D.V, newXpos, newYpos, ve)) err = 6;
        if (re != re0) D.r_ele[n] = re;
        if (ue != ue0) D.u_ele[n] = ue;
        if (ve != ve0) D.v_ele[n] = ve;
        if (err) {
            if (P.ErrorFlag == 1) { D.x[n] = Xpar; D.y[n] = Ypar; }
            particle_error(D, n, err == 4 ? LTGPU_EV_JUMP_RHO : err == 5 ? LTGPU_EV_JUMP_U : LTGPU_EV_JUMP_V, Zpar);
            return;
        }
    }
    if (P.settlementon) {                                                // :1373-1382, ledger 14
        int inp = test_settlement(D, age, re, Xpar, Ypar);
        if (inp > 0) {
            D.flags[n] |= LT_F_SETTLED;
            D.z[n] = P_depth; D.endpoly[n] = inp; D.lifespan[n] = age;
        }
    }
}
