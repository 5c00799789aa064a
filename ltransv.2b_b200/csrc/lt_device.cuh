// lt_device.cuh -- device functions of the particle step (sm_100a).
//
// Each function names the reference routine whose RESULT it reproduces
// (file:line relative to the reference's Model/ directory).  The structure is not
// the reference's: interpolation weights are computed once per (grid, RK stage)
// instead of once per value, the three hydro time levels come from one vector
// load, s-level depths are evaluated on demand instead of being tabulated, and a
// 4-knot tension spline only solves the interval it is evaluated in.
#pragma once
#include <math.h>
#include "lt_types.h"

#ifdef LT_DEBUG_TRACE
__device__ unsigned long long g_dbgcnt[8];          // debug builds only: solver-cap census
__device__ unsigned long long g_dbgcnt2[8];         // VTurb build census
__device__ unsigned long long g_dbghist[64];        // VTurb walk: intervals above / below the first one
__device__ double g_dbgcase[8];                      // inputs of the last secant-cap case
#define LT_DBG_COUNT(k) atomicAdd(&g_dbgcnt[k], 1ull)
#define LT_DBG_MAX(k, v) atomicMax(&g_dbgcnt[k], (unsigned long long)(v))
#define LT_ASSERT(c) do { if (!(c)) atomicMax(&g_dbgcnt[7], (unsigned long long)__LINE__); } while (0)   /* index checks */
#else
#define LT_DBG_COUNT(k)
#define LT_DBG_MAX(k, v)
#define LT_ASSERT(c)
#endif

#define LT_DEV __device__ __forceinline__
#define LT_DEVN __device__ __noinline__

// Fast FP64 reciprocal / quotient for the SMOOTH numerics only (splines, s-levels):
// MUFU.RCP64H seed + Newton steps, <= 1 ulp, ~8 instructions instead of ~40 with a
// slow-path branch.  Geometry predicates that the reference decides by exact equality
// (gridcell, inpoly, intersect_reflect) keep IEEE division.  -DLT_IEEE_DIV restores
// IEEE division everywhere (used to check bit-level agreement with the oracle).
LT_DEV double qrcp(double a)
{
#ifdef LT_IEEE_DIV
    return 1.0 / a;
#else
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    r = fma(fma(-a, r, 1.0), r, r);
    r = fma(fma(-a, r, 1.0), r, r);
    return r;
#endif
}
LT_DEV double qdiv(double a, double b)
{
#ifdef LT_IEEE_DIV
    return a / b;
#else
    double r = qrcp(b), q = a * r;
    return fma(fma(-b, q, a), r, q);
#endif
}

// exp(x) for x in [-746, 0] (every use in TSPACK is exp(-sigma), exp(-sigma b)): Cody-Waite
// reduction x = k ln2 + r, |r| <= ln2/2, degree-13 Taylor polynomial (truncation
// 0.347^14/14! = 4e-18 < 2^-53), scaling by 2^k through the exponent field.  < 1 ulp; about
// half the instructions of the general-purpose exp() (no overflow / NaN / positive paths).
LT_DEV double exp_neg(double x)
{
#ifdef LT_IEEE_DIV
    return exp(x);
#else
    if (x < -708.0) return 0.0;
    const double t = fma(x, 1.4426950408889634, 6755399441055744.0);     // round to nearest via the magic constant
    const int k = __double2loint(t);
    const double kd = t - 6755399441055744.0;
    double r = fma(kd, -6.93147180369123816490e-01, x);
    r = fma(kd, -1.90821492927058770002e-10, r);
    double p = 1.6059043836821613e-10;                                    // 1/13!
    p = fma(p, r, 2.08767569878681e-09);  p = fma(p, r, 2.505210838544172e-08); p = fma(p, r, 2.755731922398589e-07);
    p = fma(p, r, 2.7557319223985893e-06); p = fma(p, r, 2.48015873015873e-05); p = fma(p, r, 1.984126984126984e-04);
    p = fma(p, r, 1.3888888888888889e-03); p = fma(p, r, 8.333333333333333e-03); p = fma(p, r, 4.1666666666666664e-02);
    p = fma(p, r, 1.6666666666666666e-01); p = fma(p, r, 0.5); p = fma(p, r, 1.0); p = fma(p, r, 1.0);
    return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));   // k >= -1022 here: no denormals
#endif
}

#define kF32_1em3 0.001f        // DBLE(0.001)    is a float32 literal widened (LTRANS.f90:901)
#define kF32_1em6 0.000001f     // DBLE(0.000001) likewise                     (LTRANS.f90:1002)

// ---------------------------------------------------------------- fields ----
// One (node, level) of a hydro field = 4 ring slots in one 16-byte (f32) / 32-byte (f64)
// chunk.  The back / centre / forward records sit in slots PH, PH+1, PH+2 (mod 4), the
// fourth being refilled; PH is a template parameter so the component pick costs nothing
// (a run-time pick cost 13% of k_advect's instructions, profiles/r01_notes.md).
template <int PH> LT_DEV void pick3(double a0, double a1, double a2, double a3, double& b, double& c, double& f)
{
    if (PH == 0) { b = a0; c = a1; f = a2; }
    else if (PH == 1) { b = a1; c = a2; f = a3; }
    else if (PH == 2) { b = a2; c = a3; f = a0; }
    else { b = a3; c = a0; f = a1; }
}
// How the 4-level field profiles of k_advect are read.  At Gulf scale (3.3 GB of fields) a field line is
// touched by one or two particles per step and then never again, while the kernels' thread-local scratch
// is re-read within microseconds: LT_FIELD_LD = __ldcs marks those lines evict-first so that they do not
// push the scratch out of L2 (k_advect 31.3 -> 29.6 ms per step at 12.5 M particles; no effect when the
// whole field set fits L2).  The single-level / whole-column reads (LoadBCF::get) stay __ldg.
#ifndef LT_FIELD_LD
#define LT_FIELD_LD __ldcs
#endif
template <class T, int PH> struct LoadBCF;
template <int PH> struct LoadBCF<float, PH> {
    typedef float4 Raw;
    static LT_DEV Raw load(const float* base, size_t idx) { return LT_FIELD_LD(reinterpret_cast<const float4*>(base) + idx); }
    static LT_DEV void pick(const Raw& q, double& b, double& c, double& f)
    {
        float fb, fc, ff;
        if (PH == 0) { fb = q.x; fc = q.y; ff = q.z; } else if (PH == 1) { fb = q.y; fc = q.z; ff = q.w; }
        else if (PH == 2) { fb = q.z; fc = q.w; ff = q.x; } else { fb = q.w; fc = q.x; ff = q.y; }
        b = (double)fb; c = (double)fc; f = (double)ff;
    }
    static LT_DEV void get(const float* base, size_t idx, double& b, double& c, double& f)
    {
        float4 q = __ldg(reinterpret_cast<const float4*>(base) + idx);
        float fb, fc, ff;
        if (PH == 0) { fb = q.x; fc = q.y; ff = q.z; } else if (PH == 1) { fb = q.y; fc = q.z; ff = q.w; }
        else if (PH == 2) { fb = q.z; fc = q.w; ff = q.x; } else { fb = q.w; fc = q.x; ff = q.y; }
        b = (double)fb; c = (double)fc; f = (double)ff;
    }
};
template <int PH> struct LoadBCF<double, PH> {
    struct Raw { double2 lo, hi; };
    static LT_DEV Raw load(const double* base, size_t idx)
    {
        const double2* p = reinterpret_cast<const double2*>(base) + 2 * idx;
        Raw r; r.lo = LT_FIELD_LD(p); r.hi = LT_FIELD_LD(p + 1); return r;
    }
    static LT_DEV void pick(const Raw& q, double& b, double& c, double& f) { pick3<PH>(q.lo.x, q.lo.y, q.hi.x, q.hi.y, b, c, f); }
    static LT_DEV void get(const double* base, size_t idx, double& b, double& c, double& f)
    {
        const double2* p = reinterpret_cast<const double2*>(base) + 2 * idx;
        double2 lo = __ldg(p), hi = __ldg(p + 1);
        pick3<PH>(lo.x, lo.y, hi.x, hi.y, b, c, f);
    }
};

// ------------------------------------------------------------- gridcell -----
// gridcell_module.f90:26-257, single element.  q = x0..x3,y0..y3.
LT_DEVN bool gridcell(const double* __restrict__ q, double X, double Y)
{
    double x0 = q[0], x1 = q[1], x2 = q[2], x3 = q[3], y0 = q[4], y1 = q[5], y2 = q[6], y3 = q[7];
    if ((Y < y0 && Y < y1 && Y < y2 && Y < y3) || (Y > y0 && Y > y1 && Y > y2 && Y > y3)) return false;
    if ((X < x0 && X < x1 && X < x2 && X < x3) || (X > x0 && X > x1 && X > x2 && X > x3)) return false;
    if ((X == x0 && Y == y0) || (X == x1 && Y == y1) || (X == x2 && Y == y2) || (X == x3 && Y == y3)) return true;
    const double ex[4] = {x0, x1, x2, x3}, ey[4] = {y0, y1, y2, y3};
    // horizontal pairs in the reference's order 12,13,14,23,24,34
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = a + 1; b < 4; ++b)
            if (ey[a] == ey[b] && Y == ey[a])
                return (ex[a] > ex[b] && X > ex[b] && X < ex[a]) || (ex[b] > ex[a] && X > ex[a] && X < ex[b]);
    if (Y == y0 || Y == y1 || Y == y2 || Y == y3) {
        double hi = fmax(fmax(y0, y1), fmax(y2, y3)), lo = fmin(fmin(y0, y1), fmin(y2, y3));
        if (Y == hi || Y == lo) return false;
    }
    int total = 0;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        double bx1 = ex[p], by1 = ey[p], bx2 = ex[(p + 1) & 3], by2 = ey[(p + 1) & 3];
        if (X <= bx1 || X <= bx2) {
            if ((by1 > by2 && Y >= by2 && Y <= by1) || (by2 > by1 && Y >= by1 && Y <= by2)) {
                if (bx1 == bx2) {
                    if (X == bx1) return true;
                    if (Y != by2) total++;
                } else {
                    double slope = (by1 - by2) / (bx1 - bx2);
                    double xi = (Y - by1 + (slope * bx1)) / slope;
                    if (xi == X) return true;
                    if (xi > X && Y != by2) total++;
                }
            }
        }
    }
    return (total & 1) != 0;
}

// Quick decision for a CONVEX element (LtGridTab::convex, checked on the host): the four edge
// cross products all carry the sign of the element's orientation <=> the point is inside; one
// of them clearly carries the opposite sign <=> it is outside.  "Clearly" = by more than 1e-9 of
// the element's area scale, about 1e7 times the rounding error of the products, so the verdict
// equals gridcell's crossing count; points within that sliver of an edge or a vertex (where the
// reference's on-edge / on-vertex rules decide) return -1 and take the literal routine.
LT_DEV int gridcell_quick(const double* __restrict__ q, double X, double Y)
{
    const double x0 = q[0], x1 = q[1], x2 = q[2], x3 = q[3], y0 = q[4], y1 = q[5], y2 = q[6], y3 = q[7];
    const double A = (x2 - x0) * (y3 - y1) - (x3 - x1) * (y2 - y0);          // twice the signed area
    const double sg = A < 0.0 ? -1.0 : 1.0;
    const double c0 = sg * ((x1 - x0) * (Y - y0) - (y1 - y0) * (X - x0)), c1 = sg * ((x2 - x1) * (Y - y1) - (y2 - y1) * (X - x1));
    const double c2 = sg * ((x3 - x2) * (Y - y2) - (y3 - y2) * (X - x2)), c3 = sg * ((x0 - x3) * (Y - y3) - (y0 - y3) * (X - x3));
    const double tol = 1e-9 * fabs(A);
    const double lo = fmin(fmin(c0, c1), fmin(c2, c3));
    if (lo > tol) return 1;
    if (lo < -tol) return 0;
    return -1;
}
LT_DEV bool gridcell_any(const LtGridTab& G, int e0, double X, double Y)
{   // e0 = 0-based element
    const double* q = G.ele + (size_t)e0 * 8;
    const int r = G.convex ? gridcell_quick(q, X, Y) : -1;
    return r < 0 ? gridcell(q, X, Y) : r != 0;
}

// setEle (hydro:1414-1532), neighbour-search form.  Returns false for "jumped over
// an element" (a 0 entry reached, ledger 17).  If all 10 entries are non-zero and none
// matches, the reference raises nothing and keeps the old element (hydro:1464-1476).
#ifndef LT_FE_ATTR
#define LT_FE_ATTR LT_DEVN
#endif
LT_FE_ATTR bool find_element(const LtGridTab& G, double X, double Y, int& ele)
{
    // ele == 0: never located (start-up screen, ltgpu_screen_initial).  The reference would index
    // its adjacency table out of bounds here; reported as "not in element" instead of faulting.
    if (ele < 1 || ele > G.nE) return false;
    const int* row = G.adj + (size_t)(ele - 1) * 10;
    for (int i = 0; i < 10; ++i) {
        int check = __ldg(row + i);
        LT_ASSERT(check >= 0 && check <= G.nE);
        if (check == 0) return false;
        if (gridcell_any(G, check - 1, X, Y)) { ele = check; return true; }
    }
    return true;
}

// ------------------------------------------------- interpolation weights ----
// mode 1,2: barycentric (t,u) in triangle (1,2,3) / (3,4,1) (hydro:1706-1714); mode 3: inverse
// distance, the four weights are t,u,w2,w3 (hydro:1726-1735); mode 4..7: on node 1..4.
struct Wt { int mode; double t, u, w2, w3; };

// setInterp (hydro:1680-1740) when `setinterp_quirk` (an on-node point outside both
// triangles keeps tOK = 2), interp (hydro:2533-2565) otherwise.  The weights are a function
// of (element, point) only: the reference recomputes them for every interpolated value
// (36 times per RK stage); here once per (grid, stage).
LT_DEVN Wt make_weights(const double* __restrict__ q, double xp, double yp, bool setinterp_quirk)
{
    double x1 = q[0], x2 = q[1], x3 = q[2], x4 = q[3], y1 = q[4], y2 = q[5], y3 = q[6], y4 = q[7];
    Wt w; w.w2 = 0.0; w.w3 = 0.0;
    w.t = qdiv((xp - x1) * (y3 - y1) + (y1 - yp) * (x3 - x1), (x2 - x1) * (y3 - y1) - (y2 - y1) * (x3 - x1));
    w.u = qdiv((xp - x1) * (y2 - y1) + (y1 - yp) * (x2 - x1), (x3 - x1) * (y2 - y1) - (y3 - y1) * (x2 - x1));
    w.mode = 1;
    if (w.t < 0. || w.u < 0. || (w.t + w.u) > 1.0) {
        w.t = qdiv((xp - x3) * (y1 - y3) + (y3 - yp) * (x1 - x3), (x4 - x3) * (y1 - y3) - (y4 - y3) * (x1 - x3));
        w.u = qdiv((xp - x3) * (y4 - y3) + (y3 - yp) * (x4 - x3), (x1 - x3) * (y4 - y3) - (y1 - y3) * (x4 - x3));
        w.mode = 2;
        if (w.t < 0. || w.u < 0. || (w.t + w.u) > 1.0) {
            bool n1 = xp == x1 && yp == y1, n2 = xp == x2 && yp == y2, n3 = xp == x3 && yp == y3, n4 = xp == x4 && yp == y4;
            if (n1 || n2 || n3 || n4) { if (!setinterp_quirk) w.mode = n4 ? 7 : n3 ? 6 : n2 ? 5 : 4; }
            else {
                double D1 = rsqrt((x1 - xp) * (x1 - xp) + (y1 - yp) * (y1 - yp));
                double D2 = rsqrt((x2 - xp) * (x2 - xp) + (y2 - yp) * (y2 - yp));
                double D3 = rsqrt((x3 - xp) * (x3 - xp) + (y3 - yp) * (y3 - yp));
                double D4 = rsqrt((x4 - xp) * (x4 - xp) + (y4 - yp) * (y4 - yp));
                double rT = qrcp(D1 + D2 + D3 + D4);
                w.t = D1 * rT; w.u = D2 * rT; w.w2 = D3 * rT; w.w3 = D4 * rT;
                w.mode = 3;
            }
        }
    }
    return w;
}
LT_DEV double combine(const Wt& w, double v1, double v2, double v3, double v4)
{
    if (w.mode <= 2) {      // v1 + (v2-v1) t + (v3-v1) u   or   v3 + (v4-v3) t + (v1-v3) u, without diverging
        const bool m2 = w.mode == 2;
        const double o = m2 ? v3 : v1, a = m2 ? v4 : v2, b = m2 ? v1 : v3;
        return o + (a - o) * w.t + (b - o) * w.u;
    }
    if (w.mode == 3) return w.t * v1 + w.u * v2 + w.w2 * v3 + w.w3 * v4;
    return w.mode == 4 ? v1 : w.mode == 5 ? v2 : w.mode == 6 ? v3 : v4;
}

// free-slip corner substitution (hydro:1936-1994, 2330-2519), incl. ledger 16
LT_DEVN void freeslip(double* v, const int* m, int one_land_sum, const int* md)
{
    int sum = m[0] + m[1] + m[2] + m[3];
    if (sum >= 4) return;
    if (sum == one_land_sum) {
        if (m[0] == 0) v[0] = 0.5 * (v[1] + v[3]);
        else if (m[1] == 0) v[1] = 0.5 * (v[0] + v[2]);
        else if (m[2] == 0) v[2] = 0.5 * (v[1] + v[3]);
        else if (m[3] == 0) v[3] = 0.5 * (v[0] + v[2]);
    } else if (sum == 2) {
        if (m[0] == 0 && m[1] == 0) { v[0] = v[3]; v[1] = v[2]; }
        else if (m[1] == 0 && m[2] == 0) { v[1] = v[0]; v[2] = v[3]; }
        else if (m[2] == 0 && m[3] == 0) { v[2] = v[1]; v[3] = v[0]; }
        else if (m[3] == 0 && m[0] == 0) { v[3] = v[2]; v[0] = v[1]; }
        else if (md[0] == 0 && md[2] == 0) { v[0] = v[3]; v[2] = v[1]; }
        else if (md[3] == 0 && md[1] == 0) { v[3] = v[0]; v[1] = v[2]; }
    } else if (sum == 1) {
        if (m[0] == 1) { v[1] = v[0]; v[2] = v[0]; v[3] = v[0]; }
        else if (m[1] == 1) { v[0] = v[1]; v[2] = v[1]; v[3] = v[1]; }
        else if (m[2] == 1) { v[0] = v[2]; v[1] = v[2]; v[3] = v[2]; }
        else if (m[3] == 1) { v[0] = v[3]; v[1] = v[3]; v[2] = v[3]; }
    }
}

// which grid a stencil sits on
enum { G_RHO = 0, G_U = 1, G_V = 2 };
struct Stencil {            // element corner nodes + coordinates + weights at one point
    int4 nd; const double* q; Wt w; double xp, yp;
};

// getInterp (hydro:1743-2005) / interp (hydro:2008-2569) with precomputed weights.
// free-slip substitution on the three time levels of one stencil (kept out of line: it is off
// in the shipped configuration and would otherwise be inlined into every gather)
LT_DEVN void gather_freeslip(const LtDev& D, double* b, double* c, double* f, int4 nd, int grid, int4 und)
{
    const uint8_t* mk = grid == G_RHO ? D.R.mask : grid == G_U ? D.U.mask : D.V.mask;
    int m[4] = { mk[nd.x], mk[nd.y], mk[nd.z], mk[nd.w] }, md[4];
    md[0] = m[0]; md[1] = m[1]; md[2] = m[2]; md[3] = m[3];
    if (grid == G_V) {   // v_mask(unode*) in the diagonal test (hydro:2500-2503); out of range -> water
        int nv = D.V.nodes;
        md[0] = und.x < nv ? D.V.mask[und.x] : 1; md[1] = und.y < nv ? D.V.mask[und.y] : 1;
        md[2] = und.z < nv ? D.V.mask[und.z] : 1; md[3] = und.w < nv ? D.V.mask[und.w] : 1;
    }
    int one = grid == G_U ? 1 : 3;                       // hydro:2408
    freeslip(b, m, one, md); freeslip(c, m, one, md); freeslip(f, m, one, md);
}

// value of one (field, level) at the three hydro times: the 4-corner gather of
// getInterp (hydro:1743-2005) / interp (hydro:2008-2569) with precomputed weights.
// Out of line on purpose: ~16 call sites per kernel; inlined they made k_advect 15 k SASS
// instructions and instruction fetch its top stall (profiles/r01_notes.md).
// FreeSlip works on COPIES: taking the address of b/c/f for the (normally dead) call would
// force the twelve values of every gather into local memory.
#define LT_FREESLIP4(b0, b1, b2, b3, c0, c1, c2, c3, f0, f1, f2, f3) do { if (D.P.FreeSlip) { \
        double bb_[4] = {b0, b1, b2, b3}, cc_[4] = {c0, c1, c2, c3}, ff_[4] = {f0, f1, f2, f3}; \
        gather_freeslip(D, bb_, cc_, ff_, s.nd, grid, und); \
        b0 = bb_[0]; b1 = bb_[1]; b2 = bb_[2]; b3 = bb_[3]; c0 = cc_[0]; c1 = cc_[1]; c2 = cc_[2]; c3 = cc_[3]; \
        f0 = ff_[0]; f1 = ff_[1]; f2 = ff_[2]; f3 = ff_[3]; } } while (0)

template <class T, int PH>
LT_DEV void gather_bcf_inl(const LtDev& D, const T* fld, int L, int lev0, const Stencil& s, int grid, int4 und,
                           double& rb, double& rc, double& rf)
{
    const Wt w = s.w;
    double b0, b1, b2, b3, c0, c1, c2, c3, f0, f1, f2, f3;
    LoadBCF<T, PH>::get(fld, (size_t)s.nd.x * L + lev0, b0, c0, f0);
    LoadBCF<T, PH>::get(fld, (size_t)s.nd.y * L + lev0, b1, c1, f1);
    LoadBCF<T, PH>::get(fld, (size_t)s.nd.z * L + lev0, b2, c2, f2);
    LoadBCF<T, PH>::get(fld, (size_t)s.nd.w * L + lev0, b3, c3, f3);
    LT_FREESLIP4(b0, b1, b2, b3, c0, c1, c2, c3, f0, f1, f2, f3);
    rb = combine(w, b0, b1, b2, b3);
    rc = combine(w, c0, c1, c2, c3);
    rf = combine(w, f0, f1, f2, f3);
}
template <class T, int PH>
LT_DEVN void gather_bcf(const LtDev& D, const T* fld, int L, int lev0, const Stencil& s, int grid, int4 und,
                        double& rb, double& rc, double& rf)
{
    gather_bcf_inl<T, PH>(D, fld, L, lev0, s, grid, und, rb, rc, rf);
}

// Four consecutive levels of one field at the three hydro times (the 4-knot profile of
// WCTS_ITPI, hydro:2603-2610).  With the [node][level][slot] layout the four levels of a
// corner are 64 contiguous bytes (f32), and all 16 vector loads are issued before the first
// use, so one memory latency is exposed per profile instead of one per level.
template <class T, int PH>
LT_DEVN void gather4_bcf(const LtDev& D, const T* fld, int L, int lev0, const Stencil& s, int grid, int4 und,
                         double* __restrict__ vb, double* __restrict__ vc, double* __restrict__ vf)
{
    // all 16 vector loads are issued first and held RAW (float4: 64 registers for f32; the
    // converted doubles would need 96 and spill), then converted and combined level by level
    typedef typename LoadBCF<T, PH>::Raw Raw;
    const size_t n0 = (size_t)s.nd.x * L + lev0, n1 = (size_t)s.nd.y * L + lev0, n2 = (size_t)s.nd.z * L + lev0, n3 = (size_t)s.nd.w * L + lev0;
    Raw r[4][4];                                 // [level][corner]
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        r[i][0] = LoadBCF<T, PH>::load(fld, n0 + i); r[i][1] = LoadBCF<T, PH>::load(fld, n1 + i);
        r[i][2] = LoadBCF<T, PH>::load(fld, n2 + i); r[i][3] = LoadBCF<T, PH>::load(fld, n3 + i);
    }
    const Wt w = s.w;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        double b0, b1, b2, b3, c0, c1, c2, c3, f0, f1, f2, f3;
        LoadBCF<T, PH>::pick(r[i][0], b0, c0, f0); LoadBCF<T, PH>::pick(r[i][1], b1, c1, f1);
        LoadBCF<T, PH>::pick(r[i][2], b2, c2, f2); LoadBCF<T, PH>::pick(r[i][3], b3, c3, f3);
        LT_FREESLIP4(b0, b1, b2, b3, c0, c1, c2, c3, f0, f1, f2, f3);
        vb[i] = combine(w, b0, b1, b2, b3);
        vc[i] = combine(w, c0, c1, c2, c3);
        vf[i] = combine(w, f0, f1, f2, f3);
    }
}

LT_DEV double gather_static(const LtDev& D, const double* arr, const Stencil& s)
{
    double v[4] = { __ldg(arr + s.nd.x), __ldg(arr + s.nd.y), __ldg(arr + s.nd.z), __ldg(arr + s.nd.w) };
    if (D.P.FreeSlip) {
        int m[4] = { D.R.mask[s.nd.x], D.R.mask[s.nd.y], D.R.mask[s.nd.z], D.R.mask[s.nd.w] };
        freeslip(v, m, 3, m);
    }
    return combine(s.w, v[0], v[1], v[2], v[3]);
}

// ---------------------------------------------------------------- TSPACK ----
// SNHCSH tension_module.f90:784-850
LT_DEV void snhcsh(double X, double& SINHM, double& COSHM, double& COSHMM)
{
    const double P1 = -3.51754964808151394800e5, P2 = -1.15614435765005216044e4,
                 P3 = -1.63725857525983828727e2, P4 = -7.89474443963537015605e-1,
                 Q1 = -2.11052978884890840399e6, Q2 = 3.61578279834431989373e4,
                 Q3 = -2.77711081420602794433e2, Q4 = 1.0;
    double AX = fabs(X), XS = AX * AX;
    if (AX <= .5) {
        double XC = X * XS;
        double P = ((P4 * XS + P3) * XS + P2) * XS + P1;
        double Q = ((Q4 * XS + Q3) * XS + Q2) * XS + Q1;
        SINHM = XC * qdiv(P, Q);
        double XSD4 = .25 * XS, XSD2 = XSD4 + XSD4;
        P = ((P4 * XSD4 + P3) * XSD4 + P2) * XSD4 + P1;
        Q = ((Q4 * XSD4 + Q3) * XSD4 + Q2) * XSD4 + Q1;
        double F = XSD4 * qdiv(P, Q);
        COSHMM = XSD2 * F * (F + 2.0);
        COSHM = COSHMM + XSD2;
    } else {
        double EXPX = exp(AX);
        double RE_ = qrcp(EXPX);
        SINHM = -(((RE_ + AX) + AX) - EXPX) / 2.0;
        if (X < 0.0) SINHM = -SINHM;
        COSHM = ((RE_ - 2.0) + EXPX) / 2.0;
        COSHMM = COSHM - XS / 2.0;
    }
}

// SIGS (tension_module.f90:314-782) for ONE interval: minimum tension factor on
// [x_i, x_i+1] given the end slopes.  Intervals are independent in SIGS (TOL = 0,
// SIGMA zeroed on entry), so solving only the interval that is evaluated gives the
// value the reference's full sweep stores for it.  err = 1 <=> NIT > 10000 (SigErr).
#define LT_RTOL (200.0 * 1.1102230246251565e-16)      /* 200 * 2^-53, the :433-439 loop */
// Initial guess of the convexity equation's root.  The root of  s (cosh s - 1)/(sinh s - s) = T + 1
// is a function of T alone; g_sigtab holds root(x)/x on x = sqrt(10 T - 20) in [0, 32] (x is the
// reference's own starting value, exact to leading order as T -> 2), filled by ltgpu_create.
// Starting Newton ~1e-7 from the root instead of 5-50 % away cuts the iterations from 5-7 to
// 2-3; the stopping rule, the NIT cap and therefore the converged value are the reference's.
#define LT_SIGTAB_N 8192
#define LT_SIGTAB_INV_H 256.0
__device__ double g_sigtab[LT_SIGTAB_N + 1];
LT_DEV double sig_guess(double T)
{
    double x = sqrt(10.0 * T - 20.0);
#ifdef LT_IEEE_DIV
    return x;
#else
    double u = x * LT_SIGTAB_INV_H;
    if (!(u < (double)LT_SIGTAB_N)) return T + 1.0;
    int i = (int)u; double f = u - (double)i;
    double c0 = g_sigtab[i], c1 = g_sigtab[i + 1];
    return x * (c0 + f * (c1 - c0));
#endif
}

// One Newton iteration of the convexity equation  SIG * T1(SIG) = TP1  (tension:528-579).
// Returns true when the loop of the reference would exit; `out` is then the tension factor
// (0 with err = 1 when the reference would raise SigErr).  State: SIG, NIT, chk, chk_at.
struct NewtonState { double SIG, TP1, chk; int NIT, chk_at; };
LT_DEV void newton_start(NewtonState& q, double TP1, double SIG0) { q.SIG = SIG0; q.TP1 = TP1; q.chk = SIG0; q.NIT = 0; q.chk_at = 1; }
LT_DEV bool newton_step(NewtonState& q, double& out, int& err)
{
    const double SBIG = 85.0, RTOL = LT_RTOL, FTOL = 0.0;
    double T1, FP, SIG = q.SIG;
    if (SIG <= .5) {
        double SINHM, COSHM, COSHMM;
        snhcsh(SIG, SINHM, COSHM, COSHMM);
        double RS_ = qrcp(SINHM);
        T1 = qdiv(COSHM, SINHM);
        FP = T1 + SIG * (SIG * RS_ - T1 * T1 + 1.0);
    } else {
        // products and sums rounded separately (no FMA contraction): the rounding noise of SSM
        // and F decides how often the loop wanders instead of converging, and the fall-back
        // RATE has to be the reference's
        double EMS = exp_neg(-SIG);
        double SSM = __dsub_rn(1.0, __dmul_rn(EMS, (EMS + SIG + SIG)));
        double OM = 1.0 - EMS;
        T1 = qdiv(__dmul_rn(OM, OM), SSM);
        double RM_ = qrcp(SSM);
        FP = T1 + SIG * (2.0 * SIG * EMS * RM_ - T1 * T1 + 1.0);
    }
    double F = __dsub_rn(__dmul_rn(SIG, T1), q.TP1);
    if (++q.NIT > 10000) { LT_DBG_COUNT(0); err = 1; out = 0.0; return true; }         // tension:556-559
    if (FP <= 0.0) { out = fmin(SIG, SBIG); return true; }
    double DSIG = -qdiv(F, FP);
    if (fabs(DSIG) <= RTOL * SIG || (F >= 0.0 && F <= FTOL) || fabs(F) <= RTOL) { LT_DBG_MAX(2, q.NIT); out = fmin(SIG, SBIG); return true; }
    SIG = SIG + DSIG;
    q.SIG = SIG;
    // Newton's map SIG -> SIG' is a pure function of SIG.  When F's rounding noise sits just
    // above RTOL the iteration falls into a short cycle and the reference spins until
    // NIT > 10000 and raises SigErr.  A repeated iterate proves the cycle, so Brent's
    // checkpointing reaches the same verdict without 10^4 iterations.
    if (SIG == q.chk) { LT_DBG_COUNT(1); LT_DBG_MAX(5, q.NIT); err = 1; out = 0.0; return true; }
    if (q.NIT == q.chk_at) { q.chk = SIG; q.chk_at <<= 1; }
    return false;
}

// The secant / bisection solve of SIGS' monotonicity branch (tension:638-760) for one interval with
// chord slope S, end slopes S1, S2 (D1 = S - S1, D2 = S2 - S, D1 D2 < 0) that has passed the
// early exits (S1 S >= 0, S2 S >= 0, D0 > 0, S T0 < 0).  Rare: data with an inflection inside the
// interval.  err = 1 when the loop cannot terminate (see below).
LT_DEVN double sigs_monotone_solve(double S, double S1, double S2, double D1, double D2, double D0, int& err)
{
    const double SBIG = 85.0, RTOL = LT_RTOL, FTOL = 0.0;
    double SGN = copysign(1.0, S);
    double SIG = SBIG;
    double FMAX = qdiv(SGN * (SIG * S - S1 - S2), SIG - 2.0);
    if (FMAX <= 0.0) return SBIG;
    double STOL = RTOL * SIG, F = FMAX, F0 = qdiv(SGN * D0, 3.0 * (D1 - D2)), FNEG = F0;
    double DSIG = SIG, DMAX = SIG, D1PD2 = D1 + D2, A = 0.0, E = 0.0;
    bool CONT = true;                                  // ledger 18
    int NIT = 0;
    // The reference's loop has no iteration cap (tension:664-754).  In the SIG <= .5 branch
    // A*(C2+C1) is not guarded, so SQRT can return NaN; once F is NaN and the exit test of
    // that pass fails every later pass is NaN too and the reference spins forever.  That is
    // reported as SigErr (linint fallback in WCTS_ITPI) at once; the 10^5-pass backstop the
    // oracle shares reaches the same verdict ~50 ms later.
    for (;;) {
        DSIG = qdiv(-F * DSIG, F - F0);
        if (fabs(DSIG) > fabs(DMAX) || DSIG * DMAX > 0.0) { DSIG = DMAX; F0 = FNEG; if (++NIT > 100000) { LT_DBG_COUNT(3); err = 1; return 0.0; } continue; }
        if (fabs(DSIG) < STOL / 2.0) DSIG = -copysign(STOL / 2.0, DMAX);
        SIG = SIG + DSIG;
        F0 = F;
        double C1, C2;
        if (SIG <= .5) {
            double SINHM, COSHM, COSHMM;
            snhcsh(SIG, SINHM, COSHM, COSHMM);
            C1 = SIG * COSHM * D2 - SINHM * D1PD2;
            C2 = SIG * (SINHM + SIG) * D2 - COSHM * D1PD2;
            A = C2 - C1;
            E = SIG * SINHM - COSHMM - COSHMM;
        } else {
            double EMS = exp_neg(-SIG), EMS2 = EMS + EMS, TM = 1.0 - EMS;
            double SSINH = TM * (1.0 + EMS), SSM = SSINH - SIG * EMS2, SCM = TM * TM;
            C1 = SIG * SCM * D2 - SSM * D1PD2;
            C2 = SIG * SSINH * D2 - SCM * D1PD2;
            F = FMAX;
            CONT = true;
            if (C1 * (SIG * SCM * D1 - SSM * D1PD2) >= 0.0) CONT = false;
            if (CONT) A = EMS2 * (SIG * TM * D2 + (TM - SIG) * D1PD2);
            if (A * (C2 + C1) < 0.0) CONT = false;
            if (CONT) E = SIG * SSINH - SCM - SCM;
        }
        if (CONT) F = qdiv(SGN * (E * S2 - C2) + sqrt(A * (C2 + C1)), E);
        if (++NIT > 100000) {
            LT_DBG_COUNT(3);
#ifdef LT_DEBUG_TRACE
            g_dbgcase[1] = S; g_dbgcase[3] = S1; g_dbgcase[4] = S2; g_dbgcase[5] = SIG; g_dbgcase[6] = DSIG; g_dbgcase[7] = F;
#endif
            err = 1; return 0.0;
        }
        STOL = RTOL * SIG;
        if (fabs(DMAX) <= STOL || (F >= 0.0 && F <= FTOL) || fabs(F) <= RTOL) break;
        if (!(F == F)) { LT_DBG_COUNT(6); err = 1; return 0.0; }
        DMAX = DMAX + DSIG;
        if (F0 * F > 0.0 && fabs(F) >= fabs(F0)) { DSIG = DMAX; F0 = FNEG; continue; }
        if (F0 * F <= 0.0) {
            double T1 = DMAX, T2 = FNEG;
            DMAX = DSIG; FNEG = F0;
            if (fabs(DSIG) > fabs(T1) && fabs(F) < fabs(T2)) { DSIG = T1; F0 = T2; }
        }
    }
    LT_DBG_MAX(4, NIT);
    return fmin(SIG, SBIG);
}

// SIGS classification of one interval (tension:314-527, 638-760) from its chord slope S and end
// slopes S1, S2.  Returns true when the tension factor is known (`sigma`); false when the
// convexity Newton solve is needed (T filled: TP1 = T + 1, start value sig_guess(T)).
LT_DEV bool sigs_classify_s(double S, double S1, double S2, double& sigma, double& T, int& err)
{
    const double SBIG = 85.0;
    const double D1 = S - S1, D2 = S2 - S, D1D2 = D1 * D2;
#ifdef LT_DEBUG_TRACE
    atomicAdd(&g_dbgcnt2[7], 1ull);                                       // intervals classified
#endif
    if ((D1D2 == 0.0 && S1 != S2) || (S == 0.0 && S1 * S2 > 0.0)) { sigma = SBIG; return true; }
    sigma = 0.0;
    if (D1D2 >= 0.0) {
        if (D1D2 == 0.0) return true;
        // T = MAX(D1/D2, D2/D1) is the quotient with the larger magnitude on top; it can only
        // exceed 2 when that magnitude exceeds twice the other, so most intervals skip the division
        const double a = fabs(D1), b = fabs(D2), hi = a > b ? a : b, lo = a > b ? b : a;
        if (!(hi > 2.0 * lo)) return true;
        T = qdiv(hi, lo);
        if (T <= 2.0) return true;
#ifdef LT_DEBUG_TRACE
        atomicAdd(&g_dbgcnt2[4], 1ull);                                   // convexity solves (T > 2)
        if (T > 2.0245 && T < 2.047) atomicAdd(&g_dbgcnt2[5], 1ull);      // ... inside the band where the Newton loop can cycle
        if (T > 2.02 && T < 2.10) atomicAdd(&g_dbgcnt2[6], 1ull);
#endif
        return false;
    }
    // monotonicity :638-660
    if (S1 * S < 0.0 || S2 * S < 0.0) return true;
    const double T0 = 3.0 * S - S1 - S2;
    const double D0 = T0 * T0 - S1 * S2;
    if (D0 <= 0.0 || S * T0 >= 0.0) return true;
    sigma = sigs_monotone_solve(S, S1, S2, D1, D2, D0, err);
    return true;
}
LT_DEVN bool sigs_classify(double DX, double Y1, double Y2, double S1, double S2, double& sigma, double& TP1o, double& SIG0, int& err)
{
    double T = 0.0;
    if (sigs_classify_s(qdiv(Y2 - Y1, DX), S1, S2, sigma, T, err)) return true;
    TP1o = T + 1.0;
    SIG0 = sig_guess(T);                               // reference: SQRT(10 T - 20) (tension:524)
    return false;
}


// SIGS for one interval, solved on the spot (intervals are independent in SIGS: TOL = 0 and
// SIGMA zeroed on entry, so solving only the interval that is evaluated gives the value the
// reference's full sweep stores for it).  err = 1 <=> SigErr.
LT_DEV double sigs_interval(double DX, double Y1, double Y2, double S1, double S2, int& err)
{
    double sigma, TP1, SIG0;
    if (sigs_classify(DX, Y1, Y2, S1, S2, sigma, TP1, SIG0, err)) return sigma;
    NewtonState q; newton_start(q, TP1, SIG0);
    while (!newton_step(q, sigma, err)) {}
    return sigma;
}

// HVAL on one interval (tension_module.f90:1043-1117)
LT_DEVN double hval_interval(double T, double X1, double X2, double Y1, double Y2, double YP1, double YP2, double SIGMA)
{
    const double SBIG = 85.0;
    double DX = X2 - X1, U = T - X1, B2 = qdiv(U, DX), B1 = 1.0 - B2, S1 = YP1;
    double S = qdiv(Y2 - Y1, DX), D1 = S - S1, D2 = YP2 - S, SIG = fabs(SIGMA);
    if (SIG < 1.e-9) return Y1 + U * (S1 + B2 * (D1 + B1 * (D1 - D2)));
    if (SIG <= .5) {
        double SB2 = SIG * B2, SM, CM, CMM, SM2, CM2, DUMMY;
        snhcsh(SIG, SM, CM, CMM); snhcsh(SB2, SM2, CM2, DUMMY);
        double E = SIG * SM - CMM - CMM;
        return Y1 + S1 * U + DX * ((CM * SM2 - SM * CM2) * (D1 + D2) + SIG * (CM * CM2 - (SM + SIG) * SM2) * D1) * qrcp(SIG * E);
    }
    double SB1 = SIG * B1, SB2 = SIG - SB1;
    if (-SB1 > SBIG || -SB2 > SBIG) return Y1 + S * U;
    double E1 = exp_neg(-SB1), E2 = exp_neg(-SB2), EMS = E1 * E2, TM = 1.0 - EMS, TS = TM * TM, TP = 1.0 + EMS;
    double E = TM * (SIG * TP - TM - TM);
    return Y1 + S * U + DX * (TM * (TP - E1 - E2) * (D1 + D2) +
           SIG * ((E2 + EMS * (E1 - 2.0) - B1 * TS) * D1 + (E1 + EMS * (E2 - 2.0) - B2 * TS) * D2)) * qrcp(SIG * E);
}
// HPVAL on one interval (tension_module.f90:1190-1249)
LT_DEVN double hpval_interval(double T, double X1, double X2, double Y1, double Y2, double YP1, double YP2, double SIGMA)
{
    const double SBIG = 85.0;
    double DX = X2 - X1, B1 = qdiv(X2 - T, DX), B2 = 1.0 - B1, S1 = YP1;
    double S = qdiv(Y2 - Y1, DX), D1 = S - S1, D2 = YP2 - S, SIG = fabs(SIGMA);
    if (SIG < 1.e-9) return S1 + B2 * (D1 + D2 - 3.0 * B1 * (D2 - D1));
    if (SIG <= .5) {
        double SB2 = SIG * B2, SM, CM, CMM, SM2, CM2, DUMMY;
        snhcsh(SIG, SM, CM, CMM); snhcsh(SB2, SM2, CM2, DUMMY);
        double SINH2 = SM2 + SB2, E = SIG * SM - CMM - CMM;
        return S1 + ((CM * CM2 - SM * SINH2) * (D1 + D2) + SIG * (CM * SINH2 - (SM + SIG) * CM2) * D1) * qrcp(E);
    }
    double SB1 = SIG * B1, SB2 = SIG - SB1;
    if (-SB1 > SBIG || -SB2 > SBIG) return S;
    double E1 = exp_neg(-SB1), E2 = exp_neg(-SB2), EMS = E1 * E2, TM = 1.0 - EMS;
    double E = TM * (SIG * (1.0 + EMS) - TM - TM);
    return S + (TM * ((E2 - E1) * (D1 + D2) + TM * (D1 - D2)) + SIG * ((E1 * EMS - E2) * D1 + (E1 - E2 * EMS) * D2)) * qrcp(E);
}

// YPC1 interior / end formulas (tension_module.f90:852-978).  MIN(MAX(0,T), M) / MAX(MIN(0,T), -M) are
// clamps of T to [0, M] / [-M, 0] (M >= 0), written as two compares: FP64 fmin / fmax cost several
// instructions each and these clamps were 14 % of k_advect (profiles/r02_notes.md).
LT_DEV double clampd(double t, double lo, double hi) { return t < lo ? lo : (t > hi ? hi : t); }
LT_DEV double ypc1_end(double SI, double T) { const double m = 3.0 * SI; return SI >= 0.0 ? clampd(T, 0.0, m) : clampd(T, m, 0.0); }
LT_DEV double ypc1_clamp(double T, double SIM1, double SI)
{
    const double ASIM1 = fabs(SIM1), ASI = fabs(SI), m = 3.0 * (ASIM1 < ASI ? ASIM1 : ASI);
    const bool pos = !signbit(ASIM1 > ASI ? SIM1 : SI);               // SGN = SIGN(1, SI), or of SIM1 when it is the larger
    return clampd(T, pos ? 0.0 : -m, pos ? m : 0.0);
}
LT_DEV double ypc1_mid(double DXIM1, double DXI, double SIM1, double SI)
{
    return ypc1_clamp(qdiv(DXIM1 * SI + DXI * SIM1, DXIM1 + DXI), SIM1, SI);
}
LT_DEV double ypc1_mid_r(double DXIM1, double DXI, double SIM1, double SI, double rsum)
{   // ypc1_mid with 1/(DXIM1+DXI) supplied
    return ypc1_clamp((DXIM1 * SI + DXI * SIM1) * rsum, SIM1, SI);
}

// TSPSI(N=4) + HVAL, or the linint fallback: the water-column profile value of
// WCTS_ITPI (hydro:2619-2644) at T.

// ---------------------------------------------------------------- Philox ----
// Philox4x32-10, key = (seed, 0), counter = (id_lo, id_hi, step, block): see
// include/ltrans_b200.h for the stream layout.
struct Rng { unsigned id_lo, id_hi, step, seed; };
LT_DEV uint4 philox(const Rng& g, unsigned block)
{
    unsigned c0 = g.id_lo, c1 = g.id_hi, c2 = g.step, c3 = block, k0 = g.seed, k1 = 0u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        unsigned n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}
LT_DEV double u_real3(unsigned w) { return ((double)w + 0.5) / 4294967296.0; }   // random_module.f90:239-246
LT_DEV double u_real1(unsigned w) { return (double)w / 4294967295.0; }           // random_module.f90:213-220
// norm_module.f90:25-39 (cosine branch only; PI from the namelist)
LT_DEV double box_muller(const LtDev& D, unsigned w1, unsigned w2)
{
    return sqrt(-2.0 * log(u_real3(w1))) * cos(2.0 * D.P.PI * u_real3(w2));
}
LT_DEVN double box_muller_n(const LtDev& D, unsigned w1, unsigned w2) { return box_muller(D, w1, w2); }
LT_DEVN double log10_n(double x) { return log10(x); }

// -------------------------------------------------------------- boundary ----
struct Hit { double ix, iy, rx, ry; int seg; bool water; };
// inpoly (point_in_polygon_module.f90:25-167) on a double2 vertex list.
LT_DEVN bool inpoly(double x, double y, int n, const double2* __restrict__ e, bool onout)
{
    bool on = false;
    for (int i = 0; i < n; ++i) {                    // :49-57
        double2 q = __ldg(e + i);
        if (q.y == y && q.x > x) on = true;
        if (q.x == x && q.y == y) return !onout;
    }
    int crossed = 0;
    if (on) {                                        // :60-117 (vertex on the ray): literal walk
        auto hilo = [&](int i1) { double yy = __ldg(e + (i1 - 1)).y; return yy > y ? 1 : (yy < y ? -1 : 0); };
        bool first = true; int i = 1;
        for (;;) {
            if (i > n) break;
            if (hilo(i) == 0 && __ldg(e + (i - 1)).x > x) {
                if (first) { i = i + 1; continue; }
                if (hilo(i - 1) == 0) return !onout;
                int j = 1;
                for (;;) {
                    if ((i + j) == (n + 1)) j = 2 - i;
                    if (hilo(i + j) != 0) break;
                    if (__ldg(e + (i + j - 1)).x < x) return !onout;
                    j = j + 1;
                }
                if ((hilo(i - 1) + hilo(i + j)) == 0) crossed = crossed + 1;
                if (j < 0) break;
                i = i + j;
            }
            first = false;
            i = i + 1;
        }
    }
    double2 a = __ldg(e);
    for (int i = 1; i < n; ++i) {                    // :120-160
        double2 b = __ldg(e + i);
        bool skip = (a.x <= x && b.x <= x) || (a.y <= y && b.y <= y) || (a.y >= y && b.y >= y);
        if (!skip) {
            if (a.x > x && b.x > x) crossed++;
            else {
                double m = (b.y - a.y) / (b.x - a.x);
                double bb = a.y - m * a.x;
                double ix = (y - bb) / m;
                if (ix == x) return !onout;
                if (ix > x) crossed++;
            }
        }
        a = b;
    }
    return (crossed & 1) != 0;
}

// ibounds (boundary_module.f90:1548-1614)
LT_DEV bool in_any_island(const LtDev& D, double x, double y)
{
    if (D.maxisland <= 0) return false;
    int i = 1, start = 0, isle = __ldg(D.hid);
    for (;;) {
        i = i + 1;
        bool endIsle = (i == D.maxisland) || (__ldg(D.hid + i) != isle);
        if (endIsle) {
            if (inpoly(x, y, i - start, D.hxy + start, false)) return true;
            if (i == D.maxisland) break;
            start = i; isle = __ldg(D.hid + i);
        }
    }
    return false;
}

// intersect_reflect (boundary_module.f90:1620-1902): nearest intersection of the
// move (Xpos,Ypos)->(nXpos,nYpos) with any boundary segment, mirror image of the end
// point, strict `<` nearest with first index winning.

// ------------------------------------------------------------ settlement ----
// testSettlement / psettle / hsettle (settlement_module.f90:485-622).  Polygon edge
// columns 4,5 are contiguous per column (Fortran (pedges,5) column-major), so inpoly
// reads x and y from two arrays here.
LT_DEVN bool inpoly_cols(double x, double y, int n, const double* __restrict__ ex, const double* __restrict__ ey, bool onout)
{
    bool on = false;
    for (int i = 0; i < n; ++i) {
        double qx = __ldg(ex + i), qy = __ldg(ey + i);
        if (qy == y && qx > x) on = true;
        if (qx == x && qy == y) return !onout;
    }
    int crossed = 0;
    if (on) {
        auto hilo = [&](int i1) { double yy = __ldg(ey + (i1 - 1)); return yy > y ? 1 : (yy < y ? -1 : 0); };
        bool first = true; int i = 1;
        for (;;) {
            if (i > n) break;
            if (hilo(i) == 0 && __ldg(ex + (i - 1)) > x) {
                if (first) { i = i + 1; continue; }
                if (hilo(i - 1) == 0) return !onout;
                int j = 1;
                for (;;) {
                    if ((i + j) == (n + 1)) j = 2 - i;
                    if (hilo(i + j) != 0) break;
                    if (__ldg(ex + (i + j - 1)) < x) return !onout;
                    j = j + 1;
                }
                if ((hilo(i - 1) + hilo(i + j)) == 0) crossed = crossed + 1;
                if (j < 0) break;
                i = i + j;
            }
            first = false;
            i = i + 1;
        }
    }
    double ax = __ldg(ex), ay = __ldg(ey);
    for (int i = 1; i < n; ++i) {
        double bx = __ldg(ex + i), by = __ldg(ey + i);
        bool skip = (ax <= x && bx <= x) || (ay <= y && by <= y) || (ay >= y && by >= y);
        if (!skip) {
            if (ax > x && bx > x) crossed++;
            else {
                double m = (by - ay) / (bx - ax);
                double bb = ay - m * ax;
                double ixx = (y - bb) / m;
                if (ixx == x) return !onout;
                if (ixx > x) crossed++;
            }
        }
        ax = bx; ay = by;
    }
    return (crossed & 1) != 0;
}

LT_DEVN int test_settlement(const LtDev& D, double P_age, int R_ele, double Px, double Py)
{
    if (!(P_age >= D.P.pediage)) return 0;           // settletime = P_pediage (behavior:154)
    int polyin = 0, pidx = -1;
    LT_ASSERT(R_ele >= 1 && R_ele <= D.R.nE);
    for (int q = __ldg(D.elepoly_ptr + R_ele - 1); q < __ldg(D.elepoly_ptr + R_ele); ++q) {
        int pi = __ldg(D.elepoly_idx + q);
        int start = __ldg(D.poly_start + pi), size = __ldg(D.poly_size + pi);
        const double* c1 = D.polys, *c2 = D.polys + D.pedges, *c3 = c2 + D.pedges, *c4 = c3 + D.pedges, *c5 = c4 + D.pedges;
        double cx = __ldg(c2 + start - 1), cy = __ldg(c3 + start - 1);
        double dis = sqrt((Px - cx) * (Px - cx) + (Py - cy) * (Py - cy));
        if (dis > __ldg(D.poly_maxdis + pi)) continue;
        if (inpoly_cols(Px, Py, size, c4 + start - 1, c5 + start - 1, false)) {
            polyin = (int)llrint(__ldg(c1 + start - 1)); pidx = pi; break;
        }
    }
    if (polyin <= 0) return 0;
    if (D.P.holesExist) {
        for (int q = __ldg(D.polyhole_ptr + pidx); q < __ldg(D.polyhole_ptr + pidx + 1); ++q) {
            int hi = __ldg(D.polyhole_idx + q);
            int start = __ldg(D.hole_start + hi), size = __ldg(D.hole_size + hi);
            const double* c2 = D.holes + D.hedges, *c3 = c2 + D.hedges, *c4 = c3 + D.hedges, *c5 = c4 + D.hedges;
            double cx = __ldg(c2 + start - 1), cy = __ldg(c3 + start - 1);
            double dis = sqrt((Px - cx) * (Px - cx) + (Py - cy) * (Py - cy));
            if (dis > __ldg(D.hole_maxdis + hi)) continue;
            if (inpoly_cols(Px, Py, size, c4 + start - 1, c5 + start - 1, true)) return 0;   // onin = .FALSE.
        }
    }
    return polyin;
}
