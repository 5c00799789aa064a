// lt_vturb.cuh -- vertical turbulence (ver_turb_module.f90:30-380) as two kernels.
//
//   k_vbuild : the water-column fit of one particle -- KH at ws levels x 3 hydro times,
//              resampled to 4 ws + 7 points, 8-point moving average, time polynomial, clamp,
//              (b + 4c + f)/6, then TSPSI on the 4 ws knots (YPC1 slopes + SIGS tension factors,
//              ver_turb:102-279) -- is data-parallel over LEVELS, SAMPLES, KNOTS and INTERVALS,
//              not over particles.  So one warp fits the columns of its 32 particles one after
//              the other with the 32 lanes spread over the column: lane l gathers level l (the
//              four corner columns are contiguous in the [node][level][slot] layout: coalesced),
//              finds the segment of sample j by bisection in shared memory, averages knot k,
//              classifies interval k.  No lane walks a level pointer, nothing lives in thread-local
//              arrays, and every interval of the fit is examined for SigErr exactly as the
//              reference's SIGS sweep does (tension:433-782) -- the round-1 per-thread build
//              could only afford a 32-knot window.
//              Output per particle: the VW knots around its start depth (value, slope, tension)
//              in a [row][particle] scratch, the knot line (Z1, ZN), the window origin and the
//              SigErr verdict.
//   k_vwalk  : the 60 random-displacement sub-steps (ver_turb:291-337), one thread per particle,
//              register-resident; HVAL / HPVAL look the current interval up in the scratch.  A
//              particle that leaves its window (0.1 % of the walks in 5-30 m of water) refits a
//              window around its new depth with the per-thread routine of round 1 (VtCtx).
#pragma once
#include "lt_step2.cuh"

#define VT_SG 4                         // particles staged per window flush (one 32-byte sector per row)
#define VT_STAGE_LD (3 * VW + 1)        // padded: the flush reads [particle][row] transposed
#define VT_QCAP 64                      // pending tension-factor solves queued per warp
#define VT_FULL 0xffffffffu

LT_DEV double shfl_d(double v, int src) { return __shfl_sync(VT_FULL, v, src); }

// Shared memory of one warp, in doubles:
//   A  [12 ws]        column: depth, KH, segment slope, intercept ([3][ws] each); later the knot values
//   B  [3 (4 ws + 8)] resampled KH of the three hydro times; later chord slopes and knot slopes
//   M  [3 ws / 2 + 2] (3 ws ints) first sample at or above each level, per hydro time
//   Q  [3 VT_QCAP + VT_QCAP / 2 + 8]  queue of pending solves (three operands + tag) + per-particle words
//   O  [VT_SG][20 + 3]                the staged particles' own scalars (weights, zeta, depth, knot line; node ids)
//   S  [VT_SG][VT_STAGE_LD]           staged windows
__host__ __device__ __forceinline__ int vb_smem_doubles(int ws)
{
    return 12 * ws + 3 * (4 * ws + 8) + (3 * ws / 2 + 2) + (3 * VT_QCAP + VT_QCAP / 2 + 8) + VT_SG * (20 + 3) + VT_SG * VT_STAGE_LD;
}

// The time-combined knot line of one particle: x(1) = Z1, x(k) = Z1 + (k - 0.5) H, x(p2) = ZN -- an
// identity of the reference's construction (movex(i) = newx(i) + (newx(i+7) - newx(i))/2, ver_turb:187).
// So every interior interval has length H and the two end intervals 1.5 H; the reference's own
// abscissae differ from this by their rounding (1e-16 |x|), i.e. 1e-14 of H.
struct VbKnots {
    double Z1, ZN, H; int p2;
    LT_DEV double x(int k) const { double v = fma((double)k - 0.5, H, Z1); v = k <= 1 ? Z1 : v; return k >= p2 ? ZN : v; }
};

// Up to 32 queued tension-factor solves, one per lane: the convexity Newton loop (tension:528-579,
// operand T) or the monotonicity secant loop (tension:638-760, operands S, S1, S2).
// tag = kind << 24 | particle slot << 16 | window row of the tension factor (0xffff: outside the window).
LT_DEV void vb_drain(const double* __restrict__ qa, const double* __restrict__ qb, const double* __restrict__ qc,
                     const int* __restrict__ qtag, int first, int cnt, double* __restrict__ stage, int* __restrict__ errm)
{
    const int lane = threadIdx.x & 31;
    if (lane < cnt) {
        const int tag = qtag[first + lane];
        const int pp = (tag >> 16) & 0xff, row = tag & 0xffff;
        double sigma = 0.0; int e = 0;
        if ((tag >> 24) == 0) {
            const double T = qa[first + lane];
            NewtonState ns; newton_start(ns, T + 1.0, sig_guess(T));
            while (!newton_step(ns, sigma, e)) {}
        } else {
            const double S = qa[first + lane], S1 = qb[first + lane], S2 = qc[first + lane];
            const double T0 = 3.0 * S - S1 - S2;
            sigma = sigs_monotone_solve(S, S1, S2, S - S1, S2 - S, T0 * T0 - S1 * S2, e);
        }
        if (row != 0xffff) stage[pp * VT_STAGE_LD + 2 * VW + row] = sigma;
        if (e) atomicOr(errm, 1 << pp);
    }
    __syncwarp();
}

#define VT_OWN_D 20                     // doubles / ints of one staged particle's own scalars
#define VT_OWN_I 6

template <class T, int PH>
LT_DEV void vbuild_warp(const LtDev& D, double* __restrict__ sm, int warp_first, int base, int count)
{
    const int lane = threadIdx.x & 31;
    const int L = D.P.ws, P2 = 4 * L, NSP = P2 + 8;  // levels, knots; samples j = 1 .. P2 + 7
    double* zc = sm;                    // [3][L]  w-level depths at the particle        (Pwc_wz*)
    double* kc = zc + 3 * L;            // [3][L]  KH at the particle                    (Pwc_KH*)
    double* sl = kc + 3 * L;            // [3][L]  segment slope                         (slopek*)
    double* ic = sl + 3 * L;            // [3][L]  segment intercept                     (intercept*)
    double* ys = ic + 3 * L;            // [3][NSP] resampled KH, index j                (newy*)
    int* cm = (int*)(ys + 3 * NSP);     // [3][L]  first sample index j with newx(j) >= level m
    double* qa = (double*)cm + (3 * L / 2 + 2);      // [VT_QCAP] x 3 operands of the queued solves
    double* qb = qa + VT_QCAP;
    double* qc = qb + VT_QCAP;
    int* qtag = (int*)(qc + VT_QCAP);                // [VT_QCAP]
    int* kas = qtag + VT_QCAP;                       // [VT_SG] window origins of the staged particles
    int* errm = kas + VT_SG;                         // SigErr bits of the staged particles
    double* od = (double*)(qtag + VT_QCAP) + 8;      // [VT_SG][VT_OWN_D] the staged particles' own scalars
    int* oi = (int*)(od + VT_SG * VT_OWN_D);         // [VT_SG][VT_OWN_I]
    double* stage = (double*)oi + VT_SG * VT_OWN_I / 2;   // [VT_SG][VT_STAGE_LD]
    double* fy = zc;                    // [P2 + 2] knot values, over A (dead after step 3)                (ifity)
    double* sk = ys;                    // [P2 + 2] chord slope of interval (k, k + 1), over B (dead after step 4)
    double* yp = sk + (P2 + 2);         // [P2 + 2] knot slopes                                            (YPKc)

    const int ci = warp_first - base + lane;                           // index within the chunk (scratch column)
    const bool valid = ci < count;
    const int n = !valid ? 0 : (D.vorder && !D.vb_slot_order ? D.vorder[base + ci] : base + ci);   // the particle's slot
    const bool act = valid && D.s_act[n] != 0;
    // this lane's own particle: what the other lanes need to fit its column
    int4 o_nd = make_int4(0, 0, 0, 0); Wt o_w; o_w.mode = 1; o_w.t = o_w.u = o_w.w2 = o_w.w3 = 0.0;
    double o_zb = 0.0, o_zc = 0.0, o_zf = 0.0, o_depth = 0.0, o_pzc = 0.0;
    if (act) {
        const int re = D.r_ele[n];
        const double* q = D.R.ele + (size_t)(re - 1) * 8;
        o_nd = __ldg(D.R.node + (re - 1));
        o_w = make_weights(q, D.x[n], D.y[n], true);                   // getInterp uses setInterp's weights
        o_zb = D.s_zeb[n]; o_zc = D.s_zec[n]; o_zf = D.s_zef[n]; o_depth = D.s_depth[n]; o_pzc = D.s_pzc[n];
    }
    const unsigned actmask = __ballot_sync(VT_FULL, act);
    const T* fk = (const T*)D.kh;
    const double rp2 = 1.0 / (double)P2;
    const bool window_only = D.P.vturb_window_sigs != 0;
    // Vtransform 1 and 2: z(level) - z(bottom) = (S(level) - S(bottom)) x (a factor of zeta and h only), so the
    // levels sit at the same RELATIVE depths at the three hydro times and sample j falls into the same segment
    // at all of them; one set of level marks then serves the three times (a sample that lies on a level to
    // within rounding may take the neighbouring segment, which gives the same value to within rounding).
    const bool shared_geom = D.P.Vtransform != 3;
    const double w0b = D.LW[0][0], w2b = D.LW[0][2], w0c = D.LW[1][0], w2c = D.LW[1][2], w0f = D.LW[2][0], w2f = D.LW[2][2];
    const int RS = (P2 - 1 + 31) >> 5;                                 // interior samples 5 .. P2 + 3 per lane (blocked)
    if (lane == 0) *errm = 0;
    int qlen = 0;                                                      // queued solves (warp-uniform)

    for (int g0 = 0; g0 < 32; g0 += VT_SG) {
        const unsigned gm = (actmask >> g0) & ((1u << VT_SG) - 1u);
        if (gm == 0) continue;
        {   // the group's owners park their scalars where every lane can read them, together with everything of the
            // fit that depends on the column's ends only and that the 32 lanes would otherwise each work out again:
            // the knot line (Z1, ZN, H and the reciprocals of its three interval lengths), the sample line of the
            // centre time, and the window origin (INTRVL of the start depth, tension:1287-1354)
            const int s = lane - g0;
            if (s >= 0 && s < VT_SG && ((gm >> s) & 1u)) {
                double* o = od + s * VT_OWN_D; int* oq = oi + s * VT_OWN_I;
                o[0] = o_w.t; o[1] = o_w.u; o[2] = o_w.w2; o[3] = o_w.w3; o[4] = o_zb; o[5] = o_zc; o[6] = o_zf; o[7] = o_depth; o[8] = o_pzc;
                oq[0] = o_nd.x; oq[1] = o_nd.y; oq[2] = o_nd.z; oq[3] = o_nd.w; oq[4] = o_w.mode;
                ColK oc; oc.zb = o_zb; oc.zc = o_zc; oc.zf = o_zf; oc.depth = o_depth; oc.h = -1.0 * o_depth;
                double b1, c1, f1, bN, cN, fN;
                zlev3<true>(D, oc, 0, b1, c1, f1); zlev3<true>(D, oc, L - 1, bN, cN, fN);
                VbKnots K; K.p2 = P2;
                K.Z1 = lag(D.LW4, b1, c1, f1); K.ZN = lag(D.LW4, bN, cN, fN); K.H = (K.ZN - K.Z1) * rp2;
                const double rH = qrcp(K.H), dE1 = K.x(2) - K.x(1), dEN = K.x(P2) - K.x(P2 - 1), hsc = (cN - c1) * rp2;
                o[9] = K.Z1; o[10] = K.ZN; o[11] = K.H; o[12] = rH; o[13] = qrcp(dE1); o[14] = qrcp(dEN); o[15] = dE1; o[16] = dEN;
                o[17] = c1; o[18] = hsc; o[19] = qrcp(hsc);
                int I0;
                if (o_pzc < K.Z1) I0 = 1; else if (o_pzc > K.ZN) I0 = P2 - 1;
                else {
                    I0 = (int)floor((o_pzc - K.Z1) * rH + 0.5); I0 = max(1, min(P2 - 1, I0));
                    while (I0 > 1 && o_pzc < K.x(I0)) --I0;
                    while (I0 < P2 - 1 && !(o_pzc < K.x(I0 + 1))) ++I0;
                }
                oq[5] = max(1, min(I0 - VW / 2 + 1, P2 - VW + 1));             // knots [ka, ka + VW - 1]
            }
        }
        __syncwarp();
        for (int pp = 0; pp < VT_SG; ++pp) {
            const int p = g0 + pp;
            if (!((actmask >> p) & 1u)) continue;
            // ---- 0. the owner's scalars ---------------------------------------------------
            Stencil s0; s0.q = nullptr; s0.xp = 0.0; s0.yp = 0.0;
            ColK col;
            const double* o = od + pp * VT_OWN_D;
            {
                const int* oq = oi + pp * VT_OWN_I;
                s0.nd = make_int4(oq[0], oq[1], oq[2], oq[3]); s0.w.mode = oq[4];
                s0.w.t = o[0]; s0.w.u = o[1]; s0.w.w2 = o[2]; s0.w.w3 = o[3];
                col.zb = o[4]; col.zc = o[5]; col.zf = o[6]; col.depth = o[7]; col.h = -1.0 * col.depth;
            }
            // ---- 1. KH and depth of level l at the three hydro times (ver_turb:102-108) ----
            //         plus, per level, what the single-profile test below needs: the time-interpolated KH at the
            //         three internal times (>= 0 ?) and K4 = sum_t LW4(t) KH_t, parked in the slope table's second row
            bool lev_ok = true;
            for (int l = lane; l < L; l += 32) {
                double kb, kcc, kf, zb, zcc, zf;
                gather_bcf_inl<T, PH>(D, fk, L, l, s0, G_RHO, s0.nd, kb, kcc, kf);
                zlev3<true>(D, col, l, zb, zcc, zf);
                kc[l] = kb; kc[L + l] = kcc; kc[2 * L + l] = kf;
                zc[l] = zb; zc[L + l] = zcc; zc[2 * L + l] = zf;
                const double db = kb - kcc, df = kf - kcc;
                lev_ok = lev_ok && kcc + (w0b * db + w2b * df) >= 0.0 && kcc + (w0c * db + w2c * df) >= 0.0 && kcc + (w0f * db + w2f * df) >= 0.0;
                sl[L + l] = lag(D.LW4, kb, kcc, kf);
            }
            __syncwarp();
            double z1[3], hs[3], k1[3], kN[3];
#pragma unroll
            for (int t = 0; t < 3; ++t) {
                z1[t] = zc[t * L]; const double zN = zc[t * L + L - 1];
                hs[t] = (zN - z1[t]) * rp2;
                k1[t] = kc[t * L]; kN[t] = kc[t * L + L - 1];
            }
            VbKnots K; K.p2 = P2; K.Z1 = o[9]; K.ZN = o[10]; K.H = o[11];
            // ---- 2-4, common case: ONE profile instead of three.  Without the clamps of ver_turb:264-268 the
            //         chain resample -> 8-point mean -> time polynomial -> (b + 4c + f)/6 is linear in the KH
            //         values, and with the levels at the same relative depths at the three hydro times
            //         (shared_geom) the resampling weights do not depend on the time either, so
            //         ifity(k) = [resample + mean] of K4(level) = sum_t LW4(t) KH_t(level), on the geometry of
            //         the centre time.  The clamps cannot bind when the time-interpolated KH of EVERY level is
            //         >= 0 at the three internal times (every knot value is an average of convex combinations
            //         of those): checked per column; otherwise the three-profile path below runs.  The two
            //         agree to rounding (sums of the same terms in another order).
            const bool single = __all_sync(VT_FULL, lev_ok) && shared_geom && k1[0] >= 0.0;
            if (single) {
                const double* k4 = sl + L; const double* zt = zc + L;          // centre-time geometry
                const double z1c = o[17], hsc = o[18], rhsc = o[19];
                // first sample at or above each level (the samples newx(j) = z1 + (j - 4) hs are an arithmetic
                // progression: a quotient, checked against the sample's own formula); level 0 owns sample 5
                for (int l = lane; l < L; l += 32) {
                    const double zlo = zt[l];
                    int g = P2 + 4;
                    if (hsc > 0.0 && l < L - 1) {
                        g = 4 + (int)ceil((zlo - z1c) * rhsc);
                        g = max(5, min(P2 + 4, g));
                        while (g <= P2 + 3 && fma((double)(g - 4), hsc, z1c) < zlo) ++g;
                        while (g > 5 && !(fma((double)(g - 5), hsc, z1c) < zlo)) --g;
                    }
                    cm[l] = l == 0 ? 5 : g;
                }
                const double k4top = k4[L - 1], k4bot = k4[0];
                __syncwarp();
                // one lane per segment [level l, level l + 1): slope / intercept (ver_turb:126-133) stay in
                // registers, the segment's own samples cm[l] .. cm[l + 1] - 1 are evaluated in a row (the marks are
                // monotone, so this is the reference's walking `jlo`, ver_turb:135-177)
                for (int l = lane; l < L - 1; l += 32) {
                    const double zlo = zt[l], klo = k4[l];
                    const double sseg = qdiv(klo - k4[l + 1], zlo - zt[l + 1]), bseg = klo - sseg * zlo;
                    const int ja = cm[l], jb = cm[l + 1];
                    double dj = (double)(ja - 4);
                    for (int j = ja; j < jb; ++j, dj += 1.0) ys[j] = fma(sseg, fma(dj, hsc, z1c), bseg);
                }
                if (lane < 4) { ys[1 + lane] = k1[0]; ys[P2 + 4 + lane] = k4top; }          // pads: KHb(1) below, the top value above
                __syncwarp();
                for (int k = 1 + lane; k <= P2; k += 32) {
                    const double* y = ys + min(k, P2 - 1);
                    const double avg = (y[0] + y[1] + y[2] + y[3] + y[4] + y[5] + y[6] + y[7]) * 0.125;
                    fy[k] = k == 1 ? k4bot : (k == P2 ? k4top : avg);
                }
            } else {
                // ---- 2. per level: segment slope / intercept (ver_turb:126-133) and the first sample at or
                //         above the level.  The samples newx(j) = z1 + (j - 4) hs are an arithmetic progression,
                //         so that index is a quotient, checked against the sample's own formula.
                for (int l = lane; l < L - 1; l += 32) {
    #pragma unroll
                    for (int t = 0; t < 3; ++t) {
                        const double zlo = zc[t * L + l], zhi = zc[t * L + l + 1], klo = kc[t * L + l], khi = kc[t * L + l + 1];
                        const double s = qdiv(klo - khi, zlo - zhi);
                        sl[t * L + l] = s; ic[t * L + l] = klo - s * zlo;
                        if (t == 1 || !shared_geom) {
                            int g = P2 + 4;
                            if (hs[t] > 0.0) {
                                g = 4 + (int)ceil((zlo - z1[t]) * qrcp(hs[t]));
                                g = max(5, min(P2 + 4, g));
                                while (g <= P2 + 3 && fma((double)(g - 4), hs[t], z1[t]) < zlo) ++g;
                                while (g > 5 && !(fma((double)(g - 5), hs[t], z1[t]) < zlo)) --g;
                            }
                            cm[t * L + l] = g;                                     // level 0 is never looked up
                        }
                    }
                }
                __syncwarp();
                // ---- 3. resample (ver_turb:135-177): sample j lies in the segment above the highest interior
                //         level at or below it = the reference's walking `jlo`.  Each lane takes RS consecutive
                //         samples: one bisection over the level marks, then a cursor.
                {
                    const int ja = 5 + lane * RS, jb = min(ja + RS, P2 + 4);
                    if (shared_geom) {
                        if (ja < jb) {
                            const int* c = cm + L;
                            int lo = 0, hi = L - 2;                                // segment = #{m in 1 .. L-2 : c[m] <= j}
                            while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (c[mid] <= ja) lo = mid; else hi = mid - 1; }
                            int nxt = lo < L - 2 ? c[lo + 1] : 0x7fffffff;
                            double s0_ = sl[lo], b0_ = ic[lo], s1_ = sl[L + lo], b1_ = ic[L + lo], s2_ = sl[2 * L + lo], b2_ = ic[2 * L + lo];
                            for (int j = ja; j < jb; ++j) {
                                while (nxt <= j) {
                                    ++lo; nxt = lo < L - 2 ? c[lo + 1] : 0x7fffffff;
                                    s0_ = sl[lo]; b0_ = ic[lo]; s1_ = sl[L + lo]; b1_ = ic[L + lo]; s2_ = sl[2 * L + lo]; b2_ = ic[2 * L + lo];
                                }
                                const double dj = (double)(j - 4);
                                ys[j] = fma(s0_, fma(dj, hs[0], z1[0]), b0_);
                                ys[NSP + j] = fma(s1_, fma(dj, hs[1], z1[1]), b1_);
                                ys[2 * NSP + j] = fma(s2_, fma(dj, hs[2], z1[2]), b2_);
                            }
                        }
                    } else {
    #pragma unroll
                        for (int t = 0; t < 3; ++t) {
                            if (ja < jb) {
                                const int* c = cm + t * L;
                                int lo = 0, hi = L - 2;
                                while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (c[mid] <= ja) lo = mid; else hi = mid - 1; }
                                int nxt = lo < L - 2 ? c[lo + 1] : 0x7fffffff;
                                double s = sl[t * L + lo], b = ic[t * L + lo];
                                for (int j = ja; j < jb; ++j) {
                                    while (nxt <= j) { ++lo; nxt = lo < L - 2 ? c[lo + 1] : 0x7fffffff; s = sl[t * L + lo]; b = ic[t * L + lo]; }
                                    ys[t * NSP + j] = fma(s, fma((double)(j - 4), hs[t], z1[t]), b);
                                }
                            }
                        }
                    }
                    if (lane < 12) {                                               // the pads, ver_turb:169-177
                        const int t = lane >> 2, r = lane & 3;
                        ys[t * NSP + 1 + r] = k1[0];                               // ledger 11: KHb(1) for all three times
                        ys[t * NSP + P2 + 4 + r] = t == 0 ? kN[0] : t == 1 ? kN[1] : kN[2];
                    }
                }
                __syncwarp();
                // ---- 4. 8-point moving average, time polynomial, clamp, (b + 4c + f)/6 (ver_turb:184-275)
                //         fy overlays the column, which step 3 was the last to read
                for (int k = 1 + lane; k <= P2; k += 32) {
                    double my[3];
    #pragma unroll
                    for (int t = 0; t < 3; ++t) {
                        const double* y = ys + t * NSP + min(k, P2 - 1);
                        const double avg = (y[0] + y[1] + y[2] + y[3] + y[4] + y[5] + y[6] + y[7]) * 0.125;
                        my[t] = k == 1 ? k1[t] : (k == P2 ? kN[t] : avg);          // ends: the data values (ver_turb:197-210)
                    }
                    const double db = my[0] - my[1], df = my[2] - my[1];       // lag(): centre + weighted differences
                    double fb = my[1] + (w0b * db + w2b * df), fc = my[1] + (w0c * db + w2c * df), ff = my[1] + (w0f * db + w2f * df);
                    fb = fb < 0.0 ? 0.0 : fb; fc = fc < 0.0 ? 0.0 : fc; ff = ff < 0.0 ? 0.0 : ff;
                    {   // (ifityb + 4 ifityc + ifityf)/6: quotient by the constant from its reciprocal + one correction (<= 1 ulp)
                        const double a6 = fb + 4.0 * fc + ff, q6 = a6 * 0.16666666666666666;
                        fy[k] = fma(fma(-6.0, q6, a6), 0.16666666666666666, q6);
                    }
                }
            }
            __syncwarp();                                                      // knot values complete; samples and column dead
            // ---- 5. chord slope of every interval, then the YPC1 knot slopes (tension:852-978).  Interior
            //         intervals have length H: their chord slope is a product and the three-point formula
            //         (DXIM1 SI + DXI SIM1)/(DXIM1 + DXI) the mean of the two chord slopes.
            const double rH = o[12], rE1 = o[13], rEN = o[14], dE1 = o[15], dEN = o[16];
            for (int k = 1 + lane; k <= P2 - 1; k += 32) {
                const double dy = fy[k + 1] - fy[k];
                sk[k] = dy * (k == 1 ? rE1 : (k == P2 - 1 ? rEN : rH));
            }
            __syncwarp();
            for (int k = 3 + lane; k <= P2 - 2; k += 32) {
                // the clamp interval is [0, m3] or [-m3, 0] by the sign of the chord slope of larger magnitude
                // (tension:935-945), which on equal spacing is the sign of t itself (t = 0 when they cancel):
                // clamp(t) = sign(t) min(|t|, m3), zeros included
                const double s1 = sk[k - 1], s2 = sk[k], a1 = fabs(s1), a2 = fabs(s2), m3 = 3.0 * (a1 < a2 ? a1 : a2), t = 0.5 * (s1 + s2);
                const double at = fabs(t);
                yp[k] = copysign(at < m3 ? at : m3, t);
            }
            if (lane < 4) {                                                    // knots 1, 2, P2 - 1, P2: unequal spacing
                const bool top = lane >= 2, end = lane == 0 || lane == 3;
                const double sa = sk[top ? P2 - 2 : 1], sb = sk[top ? P2 - 1 : 2];         // left, right chord slope
                const double da = top ? K.H : dE1, db_ = top ? dEN : K.H;
                // YP(1) = clamp(SI + DXI (SI - S2)/(DXI + DX2)), YP(N) likewise from the other side (tension:905-915,
                // 962-972); knots 2 and P2 - 1 the three-point formula (:930-945): one quotient serves the four lanes
                const double si = top ? sb : sa, so = top ? sa : sb, di = top ? db_ : da;
                const double q = qdiv(end ? di * (si - so) : da * sb + db_ * sa, da + db_);
                const double v = end ? ypc1_end(si, si + q) : ypc1_clamp(q, sa, sb);
                yp[lane == 0 ? 1 : lane == 1 ? 2 : lane == 2 ? P2 - 1 : P2] = v;
            }
            __syncwarp();
            // ---- 6. window origin; SIGS on every interval (tension:314-782) -------------------
            const int ka = oi[pp * VT_OWN_I + 5];                               // window origin: knots [ka, ka + VW - 1]
            double* st = stage + pp * VT_STAGE_LD;
            // SIGS' classification of every interval (tension:480-527, 638-660).  Two tiers: a screen without
            // divisions that is NECESSARY for an interval to need a solve -- convexity case with T > 2 (inside the
            // window, where the tension factor is used) or with T within a whisker of LT_BAND (outside, where only
            // the verdict of the Newton loop matters and it can fail only for T in LT_BAND, lt_step2.cuh);
            // monotonicity case with S (3 S - S1 - S2) < 0 -- and, only in passes where some lane passes it, the
            // exact tests and the queueing of the solves (convexity Newton loop, monotonicity secant loop), which
            // run 32 at a time.
            for (int k0 = 1; k0 <= P2 - 1; k0 += 32) {
                const int k = k0 + lane, kk = min(k, P2 - 1);
                const bool live = k <= P2 - 1, inwin = live && k >= ka && k <= ka + VW - 2;
                const double S = sk[kk], S1 = yp[kk], S2 = yp[kk + 1];
                const double D1 = S - S1, D2 = S2 - S, D1D2 = D1 * D2;
                // SIGMA = SBIG: D1 D2 = 0 with S1 /= S2, or S = 0 with S1 S2 > 0 (then D1 D2 = -S1 S2 exactly)
                const bool big = (D1D2 == 0.0 && S1 != S2) || (S == 0.0 && D1D2 < 0.0);
                const double a = fabs(D1), b = fabs(D2), T0 = 3.0 * S - S1 - S2;
                const double flo = inwin ? 2.0 : 2.0199, fhi = inwin ? 1.0e300 : 2.0601;       // MAX(D1/D2, D2/D1) in (flo, fhi)
                const bool use = live && (inwin || !window_only);
                const bool cand_c = D1D2 > 0.0 && ((a > flo * b && a < fhi * b) || (b > flo * a && b < fhi * a));
                const bool cand_m = D1D2 < 0.0 && S * T0 < 0.0;
                int pend = 0;
                if (__any_sync(VT_FULL, use && (cand_c || cand_m))) {
                    const double hi = a > b ? a : b, lo = a > b ? b : a;
                    const double Tq = qdiv(hi, lo);                            // = MAX(D1/D2, D2/D1) when D1 D2 > 0
                    const double D0 = T0 * T0 - S1 * S2;
                    const bool conv = cand_c && Tq > 2.0 && (inwin || (Tq > LT_BAND_LO && Tq < LT_BAND_HI));
                    const bool mono = cand_m && !(S1 * S < 0.0 || S2 * S < 0.0) && !(D0 <= 0.0);
                    pend = use ? (conv ? 1 : (mono ? 2 : 0)) : 0;
                    const unsigned pm = __ballot_sync(VT_FULL, pend != 0);
                    if (pend) {
                        const int pos = qlen + __popc(pm & ((1u << lane) - 1u));
                        qa[pos] = conv ? Tq : S; qb[pos] = S1; qc[pos] = S2;
                        qtag[pos] = ((pend - 1) << 24) | (pp << 16) | (inwin ? (k - ka) : 0xffff);
                    }
                    qlen += __popc(pm);
                    __syncwarp();
                    if (qlen >= 32) { vb_drain(qa, qb, qc, qtag, qlen - 32, 32, stage, errm); qlen -= 32; }
                }
                if (inwin && !pend) st[2 * VW + (k - ka)] = big ? 85.0 : 0.0;
            }
            for (int q = lane; q < VW; q += 32) {
                const bool in = ka + q <= P2;
                st[q] = in ? fy[ka + q] : 0.0; st[VW + q] = in ? yp[ka + q] : 0.0;
                if (ka + q >= P2 || q == VW - 1) st[2 * VW + q] = 0.0;        // no interval starts at the last knot
            }
            if (lane == p) {                                                   // the owner keeps its knot line
                D.vz1[ci] = K.Z1; D.vzn[ci] = K.ZN;
            }
            if (lane == 0) kas[pp] = ka;
            __syncwarp();                                                      // A and B are overwritten by the next column
        }
        // ---- the group's remaining solves, verdicts, windows -> scratch ---------------------------
        if (qlen > 0) { vb_drain(qa, qb, qc, qtag, 0, qlen, stage, errm); qlen = 0; }
        {
            const int s = lane - g0;
            if (s >= 0 && s < VT_SG && ((gm >> s) & 1u)) D.vka[ci] = kas[s] | (((*errm >> s) & 1) << 16);
        }
        // row-major scratch: VT_SG consecutive particles per row
        for (int e = lane; e < VT_SG * 3 * VW; e += 32) {
            const int s = e % VT_SG, row = e / VT_SG;
            if ((gm >> s) & 1u) D.vw[(size_t)row * D.vw_stride + (size_t)(warp_first + g0 + s - base)] = stage[s * VT_STAGE_LD + row];
        }
        __syncwarp();
        if (lane == 0) *errm = 0;
        __syncwarp();
    }
}

// -------------------------------------------------------------------- the walk ----
// A particle that walks out of its window: refit a window around interval I with the per-thread
// routine (gather + VtCtx::build) and park it in this particle's column of the scratch.
template <class T, int PH>
LT_DEVN void vwalk_refit(const LtDev& D, int n, int i, int I, int& ka, int& ia, int& ib, bool& sigerr)
{
    const double Xpar = D.x[n], Ypar = D.y[n];
    const int re = D.r_ele[n];
    Stencil s0; s0.q = D.R.ele + (size_t)(re - 1) * 8; s0.nd = __ldg(D.R.node + (re - 1));
    s0.xp = Xpar; s0.yp = Ypar; s0.w = make_weights(s0.q, Xpar, Ypar, true);
    ColK col; col.zb = D.s_zeb[n]; col.zc = D.s_zec[n]; col.zf = D.s_zef[n]; col.depth = D.s_depth[n]; col.h = -1.0 * col.depth;
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
    V.build(max(1, min(I - VW / 2 + 1, V.p2 - VW + 1)));
    double* W = D.vw + i;
    for (int q = 0; q < VW; ++q) {
        W[(size_t)q * D.vw_stride] = V.fy[q]; W[(size_t)(VW + q) * D.vw_stride] = V.yp[q]; W[(size_t)(2 * VW + q) * D.vw_stride] = V.sg[q];
    }
    ka = V.ka; ia = V.ia; ib = V.ib; sigerr = sigerr || V.sigerr;
}

// The interval the particle is in, with everything HVAL / HPVAL need that does not depend on the
// evaluation point (tension:1043-1117, 1190-1249): end knots, chord slope, D1, D2, the tension
// regime and its constants.  Recomputed only when the particle changes interval.
struct IvC {
    double X1, X2, DX, rDX, Y1, S1, S, D1, D2, SIG, EMS, rE, rSE;
    int I, reg;                         // reg 0: |sigma| < 1e-9 (cubic), 1: <= .5 (SNHCSH), 2: exponential
};
LT_DEV void ivc_setup(IvC& c, double Y2, double P2_, double sigma)
{
    c.DX = c.X2 - c.X1; c.rDX = qrcp(c.DX);
    c.S = (Y2 - c.Y1) * c.rDX; c.D1 = c.S - c.S1; c.D2 = P2_ - c.S;
    c.SIG = fabs(sigma);
    c.reg = c.SIG < 1.e-9 ? 0 : (c.SIG <= .5 ? 1 : 2);
    if (c.reg == 2) {
        c.EMS = exp_neg(-c.SIG);                                       // E1 E2 of the reference (= exp(-SIG) to rounding)
        const double TM = 1.0 - c.EMS, E = TM * (c.SIG * (1.0 + c.EMS) - TM - TM);
        c.rE = qrcp(E); c.rSE = qrcp(c.SIG * E);
    }
}
// HPVAL (tension:1190-1249) at T in the cached interval; regime 1 is evaluated by the caller
LT_DEV double hpval_c(const IvC& c, double T)
{
    const double B1 = (c.X2 - T) * c.rDX, B2 = 1.0 - B1, D1 = c.D1, D2 = c.D2;
    if (c.reg == 0) return c.S1 + B2 * (D1 + D2 - 3.0 * B1 * (D2 - D1));
    const double SIG = c.SIG, SB1 = SIG * B1, SB2 = SIG - SB1;
    if (-SB1 > 85.0 || -SB2 > 85.0) return c.S;
    const double EMS = c.EMS, TM = 1.0 - EMS, E1 = exp_neg(-SB1), E2 = EMS * qrcp(E1);   // E2 = exp(-SB2) = EMS / E1
    return c.S + (TM * ((E2 - E1) * (D1 + D2) + TM * (D1 - D2)) + SIG * ((E1 * EMS - E2) * D1 + (E1 - E2 * EMS) * D2)) * c.rE;
}
// HVAL (tension:1043-1117)
LT_DEV double hval_c(const IvC& c, double T)
{
    const double U = T - c.X1, B2 = U * c.rDX, B1 = 1.0 - B2, D1 = c.D1, D2 = c.D2;
    if (c.reg == 0) return c.Y1 + U * (c.S1 + B2 * (D1 + B1 * (D1 - D2)));
    const double SIG = c.SIG, SB1 = SIG * B1, SB2 = SIG - SB1;
    if (-SB1 > 85.0 || -SB2 > 85.0) return c.Y1 + c.S * U;
    const double EMS = c.EMS, TM = 1.0 - EMS, TS = TM * TM, TP = 1.0 + EMS, E1 = exp_neg(-SB1), E2 = EMS * qrcp(E1);
    return c.Y1 + c.S * U + c.DX * (TM * (TP - E1 - E2) * (D1 + D2) +
           SIG * ((E2 + EMS * (E1 - 2.0) - B1 * TS) * D1 + (E1 + EMS * (E2 - 2.0) - B2 * TS) * D2)) * c.rSE;
}

template <class T, int PH>
LT_DEV void vwalk_particle(const LtDev& D, int n, int i)
{   // n = slot, i = index within the chunk (scratch column)
    if (!D.s_act[n]) return;
    const double background = (double)1.0E-6f;                          // ledger 2
    const double P_zc = D.s_pzc[n], P_depth = D.s_depth[n], P_zetac = D.s_zec[n];
    const int p2 = 4 * D.P.ws;
    VbKnots K; K.p2 = p2; K.Z1 = D.vz1[i]; K.ZN = D.vzn[i]; K.H = (K.ZN - K.Z1) * (1.0 / (double)p2);
    const double rH = qrcp(K.H);
    const int kw = D.vka[i];
    int ka = kw & 0xffff, ia = ka, ib = min(ka + VW - 2, p2 - 1);
    bool sigerr = (kw >> 16) != 0;
    const double* W = D.vw + i;
    const size_t ld = D.vw_stride;
    const Rng g = make_rng(D, n);
    const double deltat = 2.0;
    const int loop = D.P.idt / 2;                                       // :282-283
    double ParZc = P_zc;
    uint4 rnd = make_uint4(0, 0, 0, 0);
    IvC c; c.I = -1; c.X1 = 0.0; c.X2 = 0.0; c.reg = 0;
    auto load_iv = [&](double zq) {
        if (c.I >= 0 && zq >= c.X1 && zq < c.X2) return;               // still inside [X(I), X(I+1)): INTRVL gives I
        int I;                                                          // INTRVL (tension:1287-1354)
        if (zq < K.Z1) { I = 1; c.X1 = K.x(1); c.X2 = K.x(2); }
        else if (zq > K.ZN) { I = p2 - 1; c.X1 = K.x(I); c.X2 = K.x(I + 1); }
        else {
            I = (int)floor((zq - K.Z1) * rH + 0.5); I = max(1, min(p2 - 1, I));
            c.X1 = K.x(I); c.X2 = K.x(I + 1);
            while (I > 1 && zq < c.X1) { --I; c.X2 = c.X1; c.X1 = K.x(I); }
            while (I < p2 - 1 && !(zq < c.X2)) { ++I; c.X1 = c.X2; c.X2 = K.x(I + 1); }
        }
        if (I < ia || I > ib) vwalk_refit<T, PH>(D, n, i, I, ka, ia, ib, sigerr);
        const int q = I - ka;
        c.I = I;
        c.Y1 = W[(size_t)q * ld]; c.S1 = W[(size_t)(VW + q) * ld];
        ivc_setup(c, W[(size_t)(q + 1) * ld], W[(size_t)(VW + q + 1) * ld], W[(size_t)(2 * VW + q) * ld]);
    };
    // regime 1 (0 < |sigma| <= .5: T in (2, 2.025], rare) keeps the out-of-line routines on the scratch values
    auto snh = [&](double zq, bool deriv) {
        const int q = c.I - ka;
        const double Y2 = W[(size_t)(q + 1) * ld], P2_ = W[(size_t)(VW + q + 1) * ld], sg = W[(size_t)(2 * VW + q) * ld];
        return deriv ? hpval_interval(zq, c.X1, c.X2, c.Y1, Y2, c.S1, P2_, sg) : hval_interval(zq, c.X1, c.X2, c.Y1, Y2, c.S1, P2_, sg);
    };
#pragma unroll 1
    for (int it = 0; it < loop; ++it) {                                 // :291-337
        double Kprimec = 0.0;
        if (!(ParZc < P_depth || ParZc > P_zetac)) {
            load_iv(ParZc);
            if (sigerr) Kprimec = c.S;                                  // linint slope (same knots, same interval)
            else Kprimec = c.reg == 1 ? snh(ParZc, true) : hpval_c(c, ParZc);
        }
        const double KprimeZc = -1.0 * Kprimec * deltat;
        const double Z3rdc = ParZc + 0.5 * KprimeZc;
        double KH3rdc;
        if (Z3rdc < P_depth || Z3rdc > P_zetac) KH3rdc = background;
        else {
            load_iv(Z3rdc);
            if (sigerr) KH3rdc = c.S * Z3rdc + (c.Y1 - c.S * c.X1);
            else KH3rdc = c.reg == 1 ? snh(Z3rdc, false) : hval_c(c, Z3rdc);
            if (KH3rdc < background) KH3rdc = background;
        }
        if ((it & 1) == 0) rnd = philox(g, 1u + (unsigned)(it >> 1));
        // DEV (2/r KH3rd deltat)**0.5 with DEV = sqrt(-2 ln u1) cos(2 PI u2) (norm_module.f90:35-37): the two
        // square roots as one, sqrt(a) sqrt(b) = sqrt(a b) to rounding
        const unsigned w1 = (it & 1) ? rnd.z : rnd.x, w2 = (it & 1) ? rnd.w : rnd.y;
        const double rad = sqrt((-2.0 * log(u_real3(w1))) * (2.0 * KH3rdc * deltat));
        ParZc = ParZc + KprimeZc + cos(2.0 * D.P.PI * u_real3(w2)) * rad;
    }
    if (sigerr) D.nsig[n] += 1;
    D.s_turbv[n] = P_zc - ParZc;                                        // :342
}

// ------------------------------------------------- opt-in single-precision walk ----
// ltgpu_params.vturb_fp32_walk = 1 (default 0).  BASELINE.json's north star accepts statistical parity
// for turbulent runs ("the dispersion statistics are compared within a stated tolerance"): this walk
// evaluates HVAL / HPVAL and the Box-Muller deviates of the 60 sub-steps in FP32 (knot values, slopes,
// tension factors and the Philox words are the FP64 path's; the position itself stays FP64 and enters
// the spline as a single-precision offset from the interval's left knot).  It is NOT the headline:
// bench.py reports it as a secondary entry, gated by tests/test_parity_gpu.py::test_fp32_walk_*.
struct IvF {
    double X1d, X2d;                    // interval ends, FP64 (containment test, offset origin)
    float DX, rDX, Y1, S1, S, D1, D2, SIG, EMS, rE, rSE;
    int I, reg;
};
LT_DEV void ivf_setup(IvF& c, double Y1, double Y2, double P1, double P2_, double sigma)
{
    const double DX = c.X2d - c.X1d, S = (Y2 - Y1) / DX;
    c.DX = (float)DX; c.rDX = 1.0f / c.DX; c.Y1 = (float)Y1; c.S1 = (float)P1; c.S = (float)S;
    c.D1 = (float)(S - P1); c.D2 = (float)(P2_ - S);                  // differences formed in FP64: no cancellation in FP32
    c.SIG = fabsf((float)sigma);
    c.reg = c.SIG < 1.e-4f ? 0 : 2;                                    // below 1e-4 the exponential form cancels in FP32: cubic
    if (c.reg == 2) {
        const double sg = fabs(sigma), EMS = exp(-sg), TM = 1.0 - EMS, E = TM * (sg * (1.0 + EMS) - TM - TM);
        c.EMS = (float)EMS; c.rE = (float)(1.0 / E); c.rSE = (float)(1.0 / (sg * E));   // once per interval, in FP64
    }
}
LT_DEV float hpval_f(const IvF& c, float U)
{   // U = T - X1
    const float B2 = U * c.rDX, B1 = 1.0f - B2, D1 = c.D1, D2 = c.D2;
    if (c.reg == 0) return c.S1 + B2 * (D1 + D2 - 3.0f * B1 * (D2 - D1));
    const float SIG = c.SIG, SB1 = SIG * B1, SB2 = SIG - SB1;
    if (-SB1 > 85.0f || -SB2 > 85.0f) return c.S;
    const float EMS = c.EMS, TM = 1.0f - EMS, E1 = __expf(-SB1), E2 = __expf(-SB2);
    return c.S + (TM * ((E2 - E1) * (D1 + D2) + TM * (D1 - D2)) + SIG * ((E1 * EMS - E2) * D1 + (E1 - E2 * EMS) * D2)) * c.rE;
}
LT_DEV float hval_f(const IvF& c, float U)
{
    const float B2 = U * c.rDX, B1 = 1.0f - B2, D1 = c.D1, D2 = c.D2;
    if (c.reg == 0) return c.Y1 + U * (c.S1 + B2 * (D1 + B1 * (D1 - D2)));
    const float SIG = c.SIG, SB1 = SIG * B1, SB2 = SIG - SB1;
    if (-SB1 > 85.0f || -SB2 > 85.0f) return c.Y1 + c.S * U;
    const float EMS = c.EMS, TM = 1.0f - EMS, TS = TM * TM, TP = 1.0f + EMS, E1 = __expf(-SB1), E2 = __expf(-SB2);
    return c.Y1 + c.S * U + c.DX * (TM * (TP - E1 - E2) * (D1 + D2) +
           SIG * ((E2 + EMS * (E1 - 2.0f) - B1 * TS) * D1 + (E1 + EMS * (E2 - 2.0f) - B2 * TS) * D2)) * c.rSE;
}

template <class T, int PH>
LT_DEV void vwalk_particle_f32(const LtDev& D, int n, int i)
{
    if (!D.s_act[n]) return;
    const float background = 1.0E-6f;                                   // ledger 2
    const double P_zc = D.s_pzc[n], P_depth = D.s_depth[n], P_zetac = D.s_zec[n];
    const int p2 = 4 * D.P.ws;
    VbKnots K; K.p2 = p2; K.Z1 = D.vz1[i]; K.ZN = D.vzn[i]; K.H = (K.ZN - K.Z1) * (1.0 / (double)p2);
    const double rH = qrcp(K.H);
    const int kw = D.vka[i];
    int ka = kw & 0xffff, ia = ka, ib = min(ka + VW - 2, p2 - 1);
    bool sigerr = (kw >> 16) != 0;
    const double* W = D.vw + i;
    const size_t ld = D.vw_stride;
    const Rng g = make_rng(D, n);
    const float deltat = 2.0f, twopi = (float)(2.0 * D.P.PI);
    const int loop = D.P.idt / 2;
    double ParZc = P_zc;
    uint4 rnd = make_uint4(0, 0, 0, 0);
    IvF c; c.I = -1; c.X1d = 0.0; c.X2d = 0.0; c.reg = 0;
    auto load_iv = [&](double zq) {
        if (c.I >= 0 && zq >= c.X1d && zq < c.X2d) return;
        int I;
        if (zq < K.Z1) { I = 1; c.X1d = K.x(1); c.X2d = K.x(2); }
        else if (zq > K.ZN) { I = p2 - 1; c.X1d = K.x(I); c.X2d = K.x(I + 1); }
        else {
            I = (int)floor((zq - K.Z1) * rH + 0.5); I = max(1, min(p2 - 1, I));
            c.X1d = K.x(I); c.X2d = K.x(I + 1);
            while (I > 1 && zq < c.X1d) { --I; c.X2d = c.X1d; c.X1d = K.x(I); }
            while (I < p2 - 1 && !(zq < c.X2d)) { ++I; c.X1d = c.X2d; c.X2d = K.x(I + 1); }
        }
        if (I < ia || I > ib) vwalk_refit<T, PH>(D, n, i, I, ka, ia, ib, sigerr);
        const int q = I - ka;
        c.I = I;
        ivf_setup(c, W[(size_t)q * ld], W[(size_t)(q + 1) * ld], W[(size_t)(VW + q) * ld], W[(size_t)(VW + q + 1) * ld], W[(size_t)(2 * VW + q) * ld]);
    };
#pragma unroll 1
    for (int it = 0; it < loop; ++it) {
        float Kprimec = 0.0f;
        if (!(ParZc < P_depth || ParZc > P_zetac)) {
            load_iv(ParZc);
            Kprimec = sigerr ? c.S : hpval_f(c, (float)(ParZc - c.X1d));
        }
        const float KprimeZc = -1.0f * Kprimec * deltat;
        const double Z3rdc = ParZc + (double)(0.5f * KprimeZc);
        float KH3rdc;
        if (Z3rdc < P_depth || Z3rdc > P_zetac) KH3rdc = background;
        else {
            load_iv(Z3rdc);
            const float U = (float)(Z3rdc - c.X1d);
            KH3rdc = sigerr ? c.Y1 + c.S * U : hval_f(c, U);
            if (KH3rdc < background) KH3rdc = background;
        }
        if ((it & 1) == 0) rnd = philox(g, 1u + (unsigned)(it >> 1));
        const unsigned w1 = (it & 1) ? rnd.z : rnd.x, w2 = (it & 1) ? rnd.w : rnd.y;
        const float u1 = ((float)(w1 >> 8) + 0.5f) * (1.0f / 16777216.0f), u2 = ((float)(w2 >> 8) + 0.5f) * (1.0f / 16777216.0f);
        const float DEV = sqrtf(-2.0f * __logf(u1)) * __cosf(twopi * u2);
        ParZc = ParZc + (double)(KprimeZc + DEV * sqrtf(2.0f * KH3rdc * deltat));
    }
    if (sigerr) D.nsig[n] += 1;
    D.s_turbv[n] = P_zc - ParZc;
}
