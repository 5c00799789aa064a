// lt_step.cuh -- one particle, one internal time step (update_particles,
// LTRANS.f90:778-1400) as device code.  One thread per particle.
#pragma once
#include "lt_device.cuh"

// per-stage stencils of the three grids at a point (xp,yp) inside the START element
// (ledger 7: RK sub-stage points are evaluated with the start element's nodes).
struct Stage { Stencil r, u, v; };

LT_DEV void stage_weights(Stage& s, double xp, double yp)
{
    s.r.xp = s.u.xp = s.v.xp = xp; s.r.yp = s.u.yp = s.v.yp = yp;
    s.r.w = make_weights(s.r.q, xp, yp, false);
    s.u.w = make_weights(s.u.q, xp, yp, false);
    s.v.w = make_weights(s.v.q, xp, yp, false);
}

// find_currents (LTRANS.f90:1422-1614)
template <class T>
LT_DEVN void find_currents(const LtDev& D, const Stage& s, const Column& col, double Zpar,
                           double P_zb, double P_zc, double P_zf, int version,
                           double& Uad, double& Vad, double& Wad)
{
    const int us = D.P.us, ws = D.P.ws;
    const double z0 = D.P.z0;
    double wzb1 = zw(D, col, col.zb, 0), wzc1 = zw(D, col, col.zc, 0), wzf1 = zw(D, col, col.zf, 0);
    if (Zpar < wzb1 + z0 || Zpar < wzc1 + z0 || Zpar < wzf1 + z0) { Uad = 0.0; Vad = 0.0; Wad = 0.0; return; }
    double zb1 = zr(D, col, col.zb, 0), zc1 = zr(D, col, col.zc, 0), zf1 = zr(D, col, col.zf, 0);
    const T* fu = (const T*)D.u; const T* fv = (const T*)D.v; const T* fw = (const T*)D.w;
    if (Zpar < zb1 || Zpar < zc1 || Zpar < zf1) {                      // log layer :1489-1600
        double Ub, Uc, Uf, Vb, Vc, Vf, Wb, Wc, Wf;
        gather_bcf<T>(D, fu, us, 0, s.u, G_U, s.u.nd, Ub, Uc, Uf);
        gather_bcf<T>(D, fv, us, 0, s.v, G_V, s.u.nd, Vb, Vc, Vf);
        gather_bcf<T>(D, fw, ws, 1, s.r, G_RHO, s.u.nd, Wb, Wc, Wf);
        double num = log10((Zpar - wzb1) / z0);
        double wzb2 = zw(D, col, col.zb, 1), wzc2 = zw(D, col, col.zc, 1), wzf2 = zw(D, col, col.zf, 1);
        double db = log10((zb1 - wzb1) / z0), dc = log10((zc1 - wzb1) / z0), df = log10((zf1 - wzb1) / z0);
        double wb = log10((wzb2 - wzb1) / z0), wc = log10((wzc2 - wzb1) / z0), wf = log10((wzf2 - wzb1) / z0);
        double xt = D.ix[version - 1];
        Uad = time_poly(D, Ub * num / db, Uc * num / dc, Uf * num / df, xt);
        Vad = time_poly(D, Vb * num / db, Vc * num / dc, Vf * num / df, xt);
        Wad = time_poly(D, Wb * num / wb, Wc * num / wc, Wf * num / wf, xt);
        return;
    }
    int ii = level_window<false>(D, col, Zpar, us);
    int iii = level_window<true>(D, col, Zpar, ws);
    Uad = wcts<T, false>(D, fu, us, s.u, G_U, s.u.nd, col, ii, P_zb, P_zc, P_zf, version);
    Vad = wcts<T, false>(D, fv, us, s.v, G_V, s.u.nd, col, ii, P_zb, P_zc, P_zf, version);
    Wad = wcts<T, true>(D, fw, ws, s.r, G_RHO, s.u.nd, col, iii, P_zb, P_zc, P_zf, version);
}

// ------------------------------------------------------------------ VTurb ---
// ver_turb_module.f90:30-380, first device version: the reference's profile
// construction kept step for step, work arrays in a per-thread strided scratch.
#define VT(a, k) D.vt[((size_t)(a) * (size_t)(D.vt_p2 + 8) + (size_t)(k)) * (size_t)D.vt_stride + tslot]
template <class T>
LT_DEVN double vturb(const LtDev& D, const Stencil& sr, const Column& col, const Rng& g,
                     double P_zc, double P_depth, double P_zetac, size_t tslot)
{
    const double background = (double)1.0E-6f;                         // ledger 2
    const int ws = D.P.ws, p2 = ws * 4;
    const T* fk = (const T*)D.kh;
    enum { A_NYB = 0, A_NYC, A_NYF, A_FX, A_FY, A_YP, A_SG };
    // arrays are 0-based: index j-1 for the reference's j
    const double wzb1 = zw(D, col, col.zb, 0), wzc1 = zw(D, col, col.zc, 0), wzf1 = zw(D, col, col.zf, 0);
    const double wzbN = zw(D, col, col.zb, ws - 1), wzcN = zw(D, col, col.zc, ws - 1), wzfN = zw(D, col, col.zf, ws - 1);
    double KHb1, KHc1, KHf1, KHbN, KHcN, KHfN;
    gather_bcf<T>(D, fk, ws, 0, sr, G_RHO, sr.nd, KHb1, KHc1, KHf1);
    gather_bcf<T>(D, fk, ws, ws - 1, sr, G_RHO, sr.nd, KHbN, KHcN, KHfN);
    auto newx = [&](double z1, double zN, int j) { return z1 + ((double)(float)(j - 4)) * (zN - z1) / (double)p2; };
    // ii.a  proliferate: newy(j), j = 5 .. p2+3 (:135-166); one running level pointer
    // per time level, as the reference's jlo
    {
        int jb = 1, jc = 1, jf = 1;
        double lb, lc, lf, hb, hc_, hf;                                 // KH at level jlo / jlo+1
        double kb0, kc0, kf0, kb1, kc1, kf1;
        gather_bcf<T>(D, fk, ws, 0, sr, G_RHO, sr.nd, kb0, kc0, kf0);
        gather_bcf<T>(D, fk, ws, 1, sr, G_RHO, sr.nd, kb1, kc1, kf1);
        lb = kb0; lc = kc0; lf = kf0; hb = kb1; hc_ = kc1; hf = kf1;
        double zlb = wzb1, zlc = wzc1, zlf = wzf1;
        double zhb = zw(D, col, col.zb, 1), zhc = zw(D, col, col.zc, 1), zhf = zw(D, col, col.zf, 1);
        for (int j = 5; j <= p2 + 3; ++j) {
            double xb = newx(wzb1, wzbN, j), xc = newx(wzc1, wzcN, j), xf = newx(wzf1, wzfN, j);
            while (!(zhb > xb)) { jb++; zlb = zhb; lb = hb; double t1, t2; gather_bcf<T>(D, fk, ws, jb, sr, G_RHO, sr.nd, hb, t1, t2); zhb = zw(D, col, col.zb, jb); }
            while (!(zhc > xc)) { jc++; zlc = zhc; lc = hc_; double t1, t2; gather_bcf<T>(D, fk, ws, jc, sr, G_RHO, sr.nd, t1, hc_, t2); zhc = zw(D, col, col.zc, jc); }
            while (!(zhf > xf)) { jf++; zlf = zhf; lf = hf; double t1, t2; gather_bcf<T>(D, fk, ws, jf, sr, G_RHO, sr.nd, t1, t2, hf); zhf = zw(D, col, col.zf, jf); }
            double sb = (lb - hb) / (zlb - zhb), ib = lb - sb * zlb;      // :126-133
            double sc = (lc - hc_) / (zlc - zhc), ic = lc - sc * zlc;
            double sf = (lf - hf) / (zlf - zhf), iff = lf - sf * zlf;
            VT(A_NYB, j - 1) = sb * xb + ib;
            VT(A_NYC, j - 1) = sc * xc + ic;
            VT(A_NYF, j - 1) = sf * xf + iff;
        }
    }
    for (int i = 1; i <= 4; ++i) {                                      // :169-177 (ledger 11)
        VT(A_NYB, i - 1) = KHb1; VT(A_NYC, i - 1) = KHb1; VT(A_NYF, i - 1) = KHb1;
        VT(A_NYB, i + p2 + 2) = KHbN; VT(A_NYC, i + p2 + 2) = KHcN; VT(A_NYF, i + p2 + 2) = KHfN;
    }
    // iii-vi  moving average, end points, time polynomial, clamp, (b+4c+f)/6
    for (int k = 1; k <= p2; ++k) {
        double mxb, mxc, mxf, myb, myc, myf;
        if (k == 1) { mxb = wzb1; mxc = wzc1; mxf = wzf1; myb = KHb1; myc = KHc1; myf = KHf1; }
        else if (k == p2) { mxb = wzbN; mxc = wzcN; mxf = wzfN; myb = KHbN; myc = KHcN; myf = KHfN; }
        else {
            myb = (VT(A_NYB, k - 1) + VT(A_NYB, k) + VT(A_NYB, k + 1) + VT(A_NYB, k + 2) + VT(A_NYB, k + 3) + VT(A_NYB, k + 4) + VT(A_NYB, k + 5) + VT(A_NYB, k + 6)) / 8.0;
            myc = (VT(A_NYC, k - 1) + VT(A_NYC, k) + VT(A_NYC, k + 1) + VT(A_NYC, k + 2) + VT(A_NYC, k + 3) + VT(A_NYC, k + 4) + VT(A_NYC, k + 5) + VT(A_NYC, k + 6)) / 8.0;
            myf = (VT(A_NYF, k - 1) + VT(A_NYF, k) + VT(A_NYF, k + 1) + VT(A_NYF, k + 2) + VT(A_NYF, k + 3) + VT(A_NYF, k + 4) + VT(A_NYF, k + 5) + VT(A_NYF, k + 6)) / 8.0;
            double a = newx(wzb1, wzbN, k), b = newx(wzb1, wzbN, k + 7); mxb = a + (b - a) / 2.0;
            a = newx(wzc1, wzcN, k); b = newx(wzc1, wzcN, k + 7); mxc = a + (b - a) / 2.0;
            a = newx(wzf1, wzfN, k); b = newx(wzf1, wzfN, k + 7); mxf = a + (b - a) / 2.0;
        }
        double fxb = time_poly(D, mxb, mxc, mxf, D.ix[0]), fxc = time_poly(D, mxb, mxc, mxf, D.ix[1]), fxf = time_poly(D, mxb, mxc, mxf, D.ix[2]);
        double fyb = time_poly(D, myb, myc, myf, D.ix[0]), fyc = time_poly(D, myb, myc, myf, D.ix[1]), fyf = time_poly(D, myb, myc, myf, D.ix[2]);
        if (fyb < 0.0) fyb = 0.0;
        if (fyc < 0.0) fyc = 0.0;
        if (fyf < 0.0) fyf = 0.0;
        VT(A_FY, k - 1) = (fyb + 4.0 * fyc + fyf) / 6.0;
        VT(A_FX, k - 1) = (fxb + 4.0 * fxc + fxf) / 6.0;
    }
    // vii  TSPSI(p2): YPC1 (tension:852-978) then SIGS on every interval
    int sigerr = 0;
    {
        double X0 = VT(A_FX, 0), X1 = VT(A_FX, 1), X2 = VT(A_FX, 2), Y0 = VT(A_FY, 0), Y1 = VT(A_FY, 1), Y2 = VT(A_FY, 2);
        double DXI = X1 - X0, SI = (Y1 - Y0) / DXI, DX2 = X2 - X1, S2 = (Y2 - Y1) / DX2;
        VT(A_YP, 0) = ypc1_end(SI, SI + DXI * (SI - S2) / (DXI + DX2));
        double DXIM1 = 0.0, SIM1 = 0.0, xp_ = X1, yp_ = Y1;
        for (int I = 2; I <= p2 - 1; ++I) {
            double xn = VT(A_FX, I), yn = VT(A_FY, I);
            DXIM1 = DXI; DXI = xn - xp_; SIM1 = SI; SI = (yn - yp_) / DXI;
            VT(A_YP, I - 1) = ypc1_mid(DXIM1, DXI, SIM1, SI);
            xp_ = xn; yp_ = yn;
        }
        VT(A_YP, p2 - 1) = ypc1_end(SI, SI + DXI * (SI - SIM1) / (DXIM1 + DXI));
        double xa = X0, ya = Y0, da = VT(A_YP, 0);
        for (int I = 1; I <= p2 - 1; ++I) {
            double xb = VT(A_FX, I), yb = VT(A_FY, I), db = VT(A_YP, I);
            double sg = 0.0;
            if (!sigerr) { int e = 0; sg = sigs_interval(xb - xa, ya, yb, da, db, e); if (e) { sigerr = 1; sg = 0.0; } }
            VT(A_SG, I - 1) = sg;
            xa = xb; ya = yb; da = db;
        }
    }
    // ix  random displacement model, deltat = 2 s (:282-337)
    const double deltat = 2.0;
    const int loop = D.P.idt / 2;
    double ParZc = P_zc;
    const double Xfirst = VT(A_FX, 0), Xlast = VT(A_FX, p2 - 1);
    auto interval = [&](double Tq) {            // HVAL/HPVAL interval choice incl. INTRVL (tension:1287-1354)
        if (Tq < Xfirst) return 1;
        if (Tq > Xlast) return p2 - 1;
        int IL = 1, IH = p2;
        while (IH > IL + 1) { int K = (IL + IH) / 2; if (Tq < VT(A_FX, K - 1)) IH = K; else IL = K; }
        return IL;
    };
    auto lin = [&](double Tq, double& yv, double& mv) {      // linint fallback (:25-59)
        int jlo = 1, jhi = p2;
        for (;;) { int k = (jhi + jlo) / 2; if (VT(A_FX, k - 1) > Tq) jhi = k; else jlo = k; if (jhi - jlo == 1) break; }
        mv = (VT(A_FY, jlo - 1) - VT(A_FY, jhi - 1)) / (VT(A_FX, jlo - 1) - VT(A_FX, jhi - 1));
        double b = VT(A_FY, jlo - 1) - mv * VT(A_FX, jlo - 1);
        yv = mv * Tq + b;
    };
    uint4 rnd = make_uint4(0, 0, 0, 0);
    for (int i = 0; i < loop; ++i) {
        double Kprimec = 0.0;
        if (!(ParZc < P_depth || ParZc > P_zetac)) {
            if (!sigerr) {
                int I = interval(ParZc);
                Kprimec = hpval_interval(ParZc, VT(A_FX, I - 1), VT(A_FX, I), VT(A_FY, I - 1), VT(A_FY, I),
                                         VT(A_YP, I - 1), VT(A_YP, I), VT(A_SG, I - 1));
            } else { double yv; lin(ParZc, yv, Kprimec); }
        }
        double KprimeZc = -1.0 * Kprimec * deltat;
        double Z3rdc = ParZc + 0.5 * KprimeZc;
        double KH3rdc;
        if (Z3rdc < P_depth || Z3rdc > P_zetac) KH3rdc = background;
        else {
            if (!sigerr) {
                int I = interval(Z3rdc);
                KH3rdc = hval_interval(Z3rdc, VT(A_FX, I - 1), VT(A_FX, I), VT(A_FY, I - 1), VT(A_FY, I),
                                       VT(A_YP, I - 1), VT(A_YP, I), VT(A_SG, I - 1));
            } else { double mv; lin(Z3rdc, KH3rdc, mv); }
            if (KH3rdc < background) KH3rdc = background;
        }
        if ((i & 1) == 0) rnd = philox(g, 1u + (unsigned)(i >> 1));
        double DEV = (i & 1) ? box_muller(D, rnd.z, rnd.w) : box_muller(D, rnd.x, rnd.y);
        ParZc = ParZc + KprimeZc + DEV * sqrt(2.0 * KH3rdc * deltat);   // (...)**0.5, ledger 12
    }
    return P_zc - ParZc;                                                // :342
}
#undef VT

// ----------------------------------------------------------------- behave ---
// behavior_module.f90:181-551.  Per-particle constants of initBehave (:118-131) are
// uniform in v.2b, so P_swim(n,3) is a pure function of age.
struct BehavOut { double X, Y, Z; bool bott; };
template <class T>
LT_DEVN BehavOut behave(const LtDev& D, int n, const Stage& s0, const Column& col, const Rng& g, double Zpar,
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
        int deplvl = level_window<false>(D, col, Zpar, P.us);
        P_S = wcts<T, false>(D, (const T*)D.salt, P.us, s0.r, G_RHO, s0.u.nd, col, deplvl, P_zb, P_zc, P_zf, 4);
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
    int gid = (int)(D.first_id + n);
    if (EF < 1 || EF > 3) atomicMin(D.bad, gid);                         // STOP: lowest id wins
    else if (EF == 1) D.z[n] = revertZ;                                  // pn = p (x, y unchanged)
    else if (EF == 2) D.flags[n] |= LT_F_DEAD;
    else D.flags[n] |= LT_F_OOB;
    int k = atomicAdd(D.nev, 1);
    if (k < D.evcap) { D.ev[k].particle = gid; D.ev[k].code = code; D.ev[k].time = D.ix[2]; }
}

// ------------------------------------------------------------ the step ------
template <class T>
LT_DEV void step_particle(const LtDev& D, int n, size_t tslot)
{
    const ltgpu_params& P = D.P;
    const int idt = P.idt;
    if (D.ix[2] <= D.dob[n]) return;                                     // :790-795
    double age = D.age[n] + (double)(float)idt;                          // :798
    D.age[n] = age;
    uint8_t fl = D.flags[n];
    if (age >= P.deadage && P.mortality) {                               // updateStatus behavior:162-179
        if (!(P.settlementon && (fl & LT_F_SETTLED))) { fl |= LT_F_DEAD; D.flags[n] = fl; }
    }
    if (P.settlementon && (fl & LT_F_SETTLED)) return;                   // :804-816
    if (P.mortality && (fl & LT_F_DEAD)) return;
    if (P.OpenOceanBoundary && (fl & LT_F_OOB)) return;

    const double Xpar = D.x[n], Ypar = D.y[n];
    double Zold = D.z[n];
    int re = D.r_ele[n], ue = D.u_ele[n], ve = D.v_ele[n];
    {                                                                    // setEle :830
        int err = 0;
        int re0 = re, ue0 = ue, ve0 = ve;
        if (!find_element(D.R, Xpar, Ypar, re)) err = 4;
        if (!find_element(D.U, Xpar, Ypar, ue)) err = 5;
        if (!find_element(D.V, Xpar, Ypar, ve)) err = 6;
        if (re != re0) D.r_ele[n] = re;
        if (ue != ue0) D.u_ele[n] = ue;
        if (ve != ve0) D.v_ele[n] = ve;
        if (err) {
            particle_error(D, n, err == 4 ? LTGPU_EV_NOT_IN_RHO : err == 5 ? LTGPU_EV_NOT_IN_U : LTGPU_EV_NOT_IN_V, Zold);
            return;
        }
    }
    Stage st;
    st.r.q = D.R.ele + (size_t)(re - 1) * 8; st.r.nd = __ldg(D.R.node + (re - 1));
    st.u.q = D.U.ele + (size_t)(ue - 1) * 8; st.u.nd = __ldg(D.U.node + (ue - 1));
    st.v.q = D.V.ele + (size_t)(ve - 1) * 8; st.v.nd = __ldg(D.V.node + (ve - 1));
    // setInterp (:882): rho weights at the particle, with its on-node quirk
    Stencil s0 = st.r; s0.xp = Xpar; s0.yp = Ypar; s0.w = make_weights(st.r.q, Xpar, Ypar, true);
    const T* fz = (const T*)D.zeta;
    double P_depth = -1.0 * gather_static(D, D.depth, s0);               // :892-896
    double P_angle = gather_static(D, D.angle, s0);
    double P_zetab, P_zetac, P_zetaf;
    gather_bcf<T>(D, fz, 1, 0, s0, G_RHO, s0.nd, P_zetab, P_zetac, P_zetaf);
    double Zp = Zold;
    int hitB = 0;
    if (Zp < P_depth) { Zp = P_depth + (double)kF32_1em3; hitB++; }      // :900-903
    double P_zb = Zp, P_zc = Zp, P_zf = Zp;
    if (Zp > P_zetab) P_zb = P_zetab - (double)kF32_1em3;
    if (Zp > P_zetac) P_zc = P_zetac - (double)kF32_1em3;
    if (Zp > P_zetaf) P_zf = P_zetaf - (double)kF32_1em3;
    const double Zpar = polintd(D.ex, P_zb, P_zc, P_zf, D.ix[1]);        // :914 (not the p==1 triplet: ledger 8)
    Column col; col.zb = P_zetab; col.zc = P_zetac; col.zf = P_zetaf; col.depth = P_depth;

    const int ws = P.ws;
    double maxpartdepth = fmax(zw(D, col, P_zetab, 0), fmax(zw(D, col, P_zetac, 0), zw(D, col, P_zetaf, 0)));          // :981-987
    double minpartdepth = fmin(zw(D, col, P_zetab, ws - 1), fmin(zw(D, col, P_zetac, ws - 1), zw(D, col, P_zetaf, ws - 1)));
    const double ca = cos(P_angle), sa = sin(P_angle);
    const double eps6 = (double)kF32_1em6;
    double Uad, Vad, Wad, sU, sV, sW;
    // RK4 with stage times (t-h, t, t, t+h) = versions 1,2,2,3 (ledger 6)
    stage_weights(st, Xpar, Ypar);
    find_currents<T>(D, st, col, Zpar, P_zb, P_zc, P_zf, 1, Uad, Vad, Wad);
    sU = Uad; sV = Vad; sW = Wad;
    double xs = Xpar + (Uad * ca - Vad * sa) * (double)idt / 2.0;
    double ys = Ypar + (Uad * sa + Vad * ca) * (double)idt / 2.0;
    double zs = Zpar + Wad * (double)idt / 2.0;
    if (zs > minpartdepth) zs = minpartdepth - eps6;
    if (zs < maxpartdepth) zs = maxpartdepth + eps6;
    stage_weights(st, xs, ys);
    find_currents<T>(D, st, col, zs, P_zb, P_zc, P_zf, 2, Uad, Vad, Wad);
    sU += 2.0 * Uad; sV += 2.0 * Vad; sW += 2.0 * Wad;
    xs = Xpar + (Uad * ca - Vad * sa) * (double)idt / 2.0;
    ys = Ypar + (Uad * sa + Vad * ca) * (double)idt / 2.0;
    zs = Zpar + Wad * (double)idt / 2.0;
    if (zs > minpartdepth) zs = minpartdepth - eps6;
    if (zs < maxpartdepth) zs = maxpartdepth + eps6;
    stage_weights(st, xs, ys);
    find_currents<T>(D, st, col, zs, P_zb, P_zc, P_zf, 2, Uad, Vad, Wad);
    sU += 2.0 * Uad; sV += 2.0 * Vad; sW += 2.0 * Wad;
    xs = Xpar + (Uad * ca - Vad * sa) * (double)idt;
    ys = Ypar + (Uad * sa + Vad * ca) * (double)idt;
    zs = Zpar + Wad * (double)idt;
    if (zs > minpartdepth) zs = minpartdepth - eps6;
    if (zs < maxpartdepth) zs = maxpartdepth + eps6;
    stage_weights(st, xs, ys);
    find_currents<T>(D, st, col, zs, P_zb, P_zc, P_zf, 3, Uad, Vad, Wad);
    const double P_U = (sU + Uad) / 6.0, P_V = (sV + Vad) / 6.0, P_W = (sW + Wad) / 6.0;     // :1047-1049
    const double AdvectX = idt * (P_U * ca - P_V * sa);
    const double AdvectY = idt * (P_U * sa + P_V * ca);
    const double AdvectZ = idt * P_W;

    // back to the particle's own position for everything that follows
    stage_weights(st, Xpar, Ypar);
    if (P.SaltTempOn) {                                                  // :1062-1076
        int deplvl = level_window<false>(D, col, Zpar, P.us);
        D.psalt[n] = wcts<T, false>(D, (const T*)D.salt, P.us, st.r, G_RHO, st.u.nd, col, deplvl, P_zb, P_zc, P_zf, 4);
        D.ptemp[n] = wcts<T, false>(D, (const T*)D.temp, P.us, st.r, G_RHO, st.u.nd, col, deplvl, P_zb, P_zc, P_zf, 4);
    }
    long long gid = D.first_id + n;
    Rng g; g.id_lo = (unsigned)((unsigned long long)gid & 0xffffffffull); g.id_hi = (unsigned)((unsigned long long)gid >> 32);
    g.step = D.gstep; g.seed = (unsigned)P.seed;
    double TurbHx = 0.0, TurbHy = 0.0, TurbV = 0.0;
    if (P.HTurbOn) {                                                     // hor_turb_module.f90:29-50
        uint4 r = philox(g, 0u);
        double sd = sqrt(2.0 * P.ConstantHTurb * idt);
        TurbHx = box_muller(D, r.x, r.y) * sd;
        TurbHy = box_muller(D, r.z, r.w) * sd;
    }
    if (P.VTurbOn) TurbV = vturb<T>(D, s0, col, g, P_zc, P_depth, P_zetac, tslot);          // :1098
    BehavOut bo; bo.X = bo.Y = bo.Z = 0.0; bo.bott = false;
    if (P.Behavior != 0)                                                 // :1110
        bo = behave<T>(D, n, st, col, g, Zpar, P_zb, P_zc, P_zf, P_zetac, age, P_depth, P_U, P_V, P_angle);

    double newXpos = Xpar + AdvectX + TurbHx;                            // :1128-1130
    double newYpos = Ypar + AdvectY + TurbHy;
    double newZpos = Zpar + AdvectZ + TurbV;
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

    // horizontal boundary: up to 3 reflections (:1180-1232)
    double Xpos = Xpar, Ypos = Ypar, nXpos = newXpos, nYpos = newYpos;
    int skip = -1, reflects = 0, hitL = 0;
    for (;;) {
        Hit h;
        if (!intersect_reflect(D, Xpos, Ypos, nXpos, nYpos, skip, h)) break;
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
    if (!inpoly(newXpos, newYpos, D.maxbound, D.bxy, false)) { particle_error(D, n, LTGPU_EV_OUT_MAIN, Zpar); return; }   // :1240
    if (in_any_island(D, newXpos, newYpos)) { particle_error(D, n, LTGPU_EV_IN_ISLAND, Zpar); return; }                 // :1275
    // commit (:1312-1314, and the swap :1407-1414 -- particles are independent)
    D.x[n] = newXpos; D.y[n] = newYpos; D.z[n] = newZpos;
    {                                                                    // setEle at the new position :1317
        int err = 0, re0 = re, ue0 = ue, ve0 = ve;
        if (!find_element(D.R, newXpos, newYpos, re)) err = 4;
        if (!find_element(D.U, newXpos, newYpos, ue)) err = 5;
        if (!find_element(D.V, newXpos, newYpos, ve)) err = 6;
        if (re != re0) D.r_ele[n] = re;
        if (ue != ue0) D.u_ele[n] = ue;
        if (ve != ve0) D.v_ele[n] = ve;
        if (err) {
            int EF = P.ErrorFlag;
            if (EF == 1) { D.x[n] = Xpar; D.y[n] = Ypar; }               // revert x,y too (already committed)
            particle_error(D, n, err == 4 ? LTGPU_EV_JUMP_RHO : err == 5 ? LTGPU_EV_JUMP_U : LTGPU_EV_JUMP_V, Zpar);
            return;
        }
    }
    if (P.settlementon) {                                                // :1373-1382, ledger 14: OLD x,y, NEW element
        int inp = test_settlement(D, age, re, Xpar, Ypar);
        if (inp > 0) {
            D.flags[n] |= LT_F_SETTLED;
            D.z[n] = P_depth; D.endpoly[n] = inp; D.lifespan[n] = age;
        }
    }
}
