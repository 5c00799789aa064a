/*
 * ora_leaf.c -- leaf numerics of the LTRANS v.2b particle step, restated in C.
 *
 * TEST INFRASTRUCTURE ONLY (see ltrans_oracle.h).  PARITY UNPINNED BY THE
 * REFERENCE except for MT19937 (published known answers).
 *
 * Every function cites the reference file:line (relative to
 * /root/reference/Model/) it follows.  Control flow is kept in the reference's
 * order on purpose (no algebraic clean-up): this file is the checker.
 * Build with -O2 -ffp-contract=off (no FMA contraction, no fast-math).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "ltrans_oracle.h"

/* ------------------------------------------------------------------------ */
/* interpolation_module.f90:70-107  polintd (n is always 3 on this path)     */
double ora_polintd(const double xa[3], const double ya[3], double x)
{
    int ns = 1;
    double dif = fabs(x - xa[0]);
    for (int i = 1; i <= 3; ++i) {                 /* :82-88 closest entry   */
        double dift = fabs(x - xa[i - 1]);
        if (dift < dif) { ns = i; dif = dift; }
    }
    double c = (xa[1] - x) * ((ya[2] - ya[1]) / (xa[1] - xa[2]));   /* :91 */
    c = c - (xa[1] - x) * ((ya[1] - ya[0]) / (xa[0] - xa[1]));      /* :92 */
    c = c / (xa[0] - xa[2]);                                        /* :93 */
    double a, b;
    if (ns == 3) { a = (ya[2] - ya[1]) / (xa[1] - xa[2]); b = xa[2] - x; }   /* :96-98 */
    else         { a = (ya[1] - ya[0]) / (xa[0] - xa[1]); b = xa[0] - x; }   /* :99-101 */
    return ya[ns - 1] + (xa[ns - 1] - x) * a + b * c;               /* :105 */
}

/* interpolation_module.f90:25-59  linint: bisection + linear                */
void ora_linint(const double* xa, const double* ya, int32_t n, double x, double* y, double* m)
{
    int jlo = 1, jhi = n;
    for (;;) {
        int k = (jhi + jlo) / 2;
        if (xa[k - 1] > x) jhi = k; else jlo = k;
        if (jhi - jlo == 1) break;
    }
    *m = (ya[jlo - 1] - ya[jhi - 1]) / (xa[jlo - 1] - xa[jhi - 1]);
    double b = ya[jlo - 1] - (*m) * xa[jlo - 1];
    *y = (*m) * x + b;
}

/* ------------------------------------------------------------------------ */
/* gridcell_module.f90:26-257, single-element form (checkele present).
 * ex/ey = ele_x(1:4,i), ele_y(1:4,i).  Returns `triangle` (1 = in element).
 * Exact floating equality is kept everywhere (ledger 22).                    */
int32_t ora_gridcell(const double ex[4], const double ey[4], double X, double Y)
{
    /* 1 (:52-56) */
    if ((Y < ey[0] && Y < ey[1] && Y < ey[2] && Y < ey[3]) ||
        (Y > ey[0] && Y > ey[1] && Y > ey[2] && Y > ey[3])) return 0;
    /* 2 (:58-62) */
    if ((X < ex[0] && X < ex[1] && X < ex[2] && X < ex[3]) ||
        (X > ex[0] && X > ex[1] && X > ex[2] && X > ex[3])) return 0;
    /* 3 (:64-72) on a node */
    if ((X == ex[0] && Y == ey[0]) || (X == ex[1] && Y == ey[1]) ||
        (X == ex[2] && Y == ey[2]) || (X == ex[3] && Y == ey[3])) return 1;
    /* 4 (:74-155) horizontal segment between ANY pair of corners, in the
     * reference's order 12,13,14,23,24,34 */
    static const int pa[6] = {0, 0, 0, 1, 1, 2};
    static const int pb[6] = {1, 2, 3, 2, 3, 3};
    int any = 0;
    for (int q = 0; q < 6; ++q) if (ey[pa[q]] == ey[pb[q]]) any = 1;
    if (any) {
        for (int q = 0; q < 6; ++q) {
            int a = pa[q], b = pb[q];
            if (ey[a] == ey[b] && Y == ey[a]) {
                if ((ex[a] > ex[b] && X > ex[b] && X < ex[a]) ||
                    (ex[b] > ex[a] && X > ex[a] && X < ex[b])) return 1;
                else return 0;
            }
        }
    }
    /* 5 (:158-190) Y equal to a corner y: reject if it is the top or bottom */
    if (Y == ey[0] || Y == ey[1] || Y == ey[2] || Y == ey[3]) {
        double bhigh = ey[0];
        if (ey[1] > bhigh) bhigh = ey[1];
        if (ey[2] > bhigh) bhigh = ey[2];
        if (ey[3] > bhigh) bhigh = ey[3];
        if (Y == bhigh) return 0;
        double blow = ey[0];
        if (ey[1] < blow) blow = ey[1];
        if (ey[2] < blow) blow = ey[2];
        if (ey[3] < blow) blow = ey[3];
        if (Y == blow) return 0;
    }
    /* 6 (:193-247) crossings.  The `exit` inside the p-loop leaves only that
     * loop but triangle is already 1, so with checkele the answer is 1.      */
    int counter[4] = {0, 0, 0, 0};
    for (int p = 0; p < 4; ++p) {
        double bx1 = ex[p], by1 = ey[p], bx2 = ex[(p + 1) & 3], by2 = ey[(p + 1) & 3];
        if (X <= bx1 || X <= bx2) {
            if ((by1 > by2 && Y >= by2 && Y <= by1) || (by2 > by1 && Y >= by1 && Y <= by2)) {
                if (bx1 == bx2) {
                    if (X == bx1) return 1;
                    counter[p] = 1;
                    if (Y == by2) counter[p] = 0;
                } else {
                    double slope = (by1 - by2) / (bx1 - bx2);
                    double xi = (Y - by1 + (slope * bx1)) / slope;
                    if (xi > X) { counter[p] = 1; if (Y == by2) counter[p] = 0; }
                    if (xi == X) return 1;
                }
            }
        }
    }
    int total = counter[0] + counter[1] + counter[2] + counter[3];
    return (total % 2) != 0;
}

/* ------------------------------------------------------------------------ */
/* point_in_polygon_module.f90:25-167  inpoly.  e(n,2) is passed as two
 * arrays.  onin < 0 means "argument absent" (onout = false).                 */
int32_t ora_inpoly(double x, double y, int32_t n, const double* e1, const double* e2, int32_t onin)
{
    int onout = (onin < 0) ? 0 : !onin;             /* :37-41 */
    int result = 1, crossed = 0, on = 0;
    int* hilo = (int*)calloc((size_t)n + 2, sizeof(int));
    for (int i = 1; i <= n; ++i) {                  /* :49-57 */
        if (e2[i - 1] > y) hilo[i] = 1;
        if (e2[i - 1] < y) hilo[i] = -1;
        if (e2[i - 1] == y && e1[i - 1] > x) on = 1;
        if (e1[i - 1] == x && e2[i - 1] == y) {
            if (onout) result = 0;
            free(hilo); return result;
        }
    }
    if (on) {                                       /* :60-117 */
        int first = 1, i = 1;
        for (;;) {
            if (i > n) break;
            if (hilo[i] == 0 && e1[i - 1] > x) {
                if (first) { i = i + 1; continue; }
                if (hilo[i - 1] == 0) {             /* :76-79 */
                    if (onout) result = 0;
                    free(hilo); return result;
                }
                int j = 1;
                for (;;) {                          /* :85-96 */
                    if ((i + j) == (n + 1)) j = 2 - i;
                    if (hilo[i + j] != 0) break;
                    if (e1[i + j - 1] < x) {
                        if (onout) result = 0;
                        free(hilo); return result;
                    }
                    j = j + 1;
                }
                if ((hilo[i - 1] + hilo[i + j]) == 0) crossed = crossed + 1;   /* :102 */
                if (j < 0) break;                   /* :107 */
                i = i + j;                          /* :111 */
            }
            first = 0;
            i = i + 1;
        }
    }
    for (int i = 1; i <= n - 1; ++i) {              /* :120-160 */
        double ax = e1[i - 1], ay = e2[i - 1], bx = e1[i], by = e2[i];
        if (ax <= x && bx <= x) continue;
        if (ay <= y && by <= y) continue;
        if (ay >= y && by >= y) continue;
        if (ax > x && bx > x) { crossed = crossed + 1; continue; }
        double m = (by - ay) / (bx - ax);
        double b = ay - m * ax;
        double ix = (y - b) / m;
        if (ix == x) {
            if (onout) result = 0;
            free(hilo); return result;
        }
        if (ix > x) crossed = crossed + 1;
    }
    if (crossed % 2 == 0) result = 0;               /* :163 */
    free(hilo);
    return result;
}

/* ------------------------------------------------------------------------ */
/* tension_module.f90:784-850  SNHCSH                                         */
void ora_snhcsh(double X, double* SINHM, double* COSHM, double* COSHMM)
{
    static const double P1 = -3.51754964808151394800e5, P2 = -1.15614435765005216044e4,
                        P3 = -1.63725857525983828727e2, P4 = -7.89474443963537015605e-1,
                        Q1 = -2.11052978884890840399e6, Q2 = 3.61578279834431989373e4,
                        Q3 = -2.77711081420602794433e2, Q4 = 1.0;
    double AX = fabs(X), XS = AX * AX;
    if (AX <= .5) {
        double XC = X * XS;
        double P = ((P4 * XS + P3) * XS + P2) * XS + P1;
        double Q = ((Q4 * XS + Q3) * XS + Q2) * XS + Q1;
        *SINHM = XC * (P / Q);
        double XSD4 = .25 * XS, XSD2 = XSD4 + XSD4;
        P = ((P4 * XSD4 + P3) * XSD4 + P2) * XSD4 + P1;
        Q = ((Q4 * XSD4 + Q3) * XSD4 + Q2) * XSD4 + Q1;
        double F = XSD4 * (P / Q);
        *COSHMM = XSD2 * F * (F + 2.0);
        *COSHM = *COSHMM + XSD2;
    } else {
        double EXPX = exp(AX);
        *SINHM = -(((1.0 / EXPX + AX) + AX) - EXPX) / 2.0;
        if (X < 0.0) *SINHM = -*SINHM;
        *COSHM = ((1.0 / EXPX - 2.0) + EXPX) / 2.0;
        *COSHMM = *COSHM - XS / 2.0;
    }
}

/* tension_module.f90:852-978  YPC1 (N >= 3 on this path; N == 2 kept)        */
static int ypc1(int N, const double* X, const double* Y, double* YP)
{
    int NM1 = N - 1;
    double DXI = X[1] - X[0];
    if (DXI <= 0.0) return 2;
    double SI = (Y[1] - Y[0]) / DXI;
    if (NM1 == 1) { YP[0] = SI; YP[1] = SI; return 0; }
    double DX2 = X[2] - X[1];
    if (DX2 <= 0.0) return 3;
    double S2 = (Y[2] - Y[1]) / DX2;
    double T = SI + DXI * (SI - S2) / (DXI + DX2);
    if (SI >= 0.0) YP[0] = fmin(fmax(0.0, T), 3.0 * SI);
    else           YP[0] = fmax(fmin(0.0, T), 3.0 * SI);
    double DXIM1 = 0.0, SIM1 = 0.0;
    for (int I = 2; I <= NM1; ++I) {
        DXIM1 = DXI;
        DXI = X[I] - X[I - 1];
        if (DXI <= 0.0) return I + 1;
        SIM1 = SI;
        SI = (Y[I] - Y[I - 1]) / DXI;
        T = (DXIM1 * SI + DXI * SIM1) / (DXIM1 + DXI);
        double ASIM1 = fabs(SIM1), ASI = fabs(SI);
        double SGN = copysign(1.0, SI);
        if (ASIM1 > ASI) SGN = copysign(1.0, SIM1);
        if (SGN > 0.0) YP[I - 1] = fmin(fmax(0.0, T), 3.0 * fmin(ASIM1, ASI));
        else           YP[I - 1] = fmax(fmin(0.0, T), -3.0 * fmin(ASIM1, ASI));
    }
    T = SI + DXI * (SI - SIM1) / (DXIM1 + DXI);
    if (SI >= 0.0) YP[N - 1] = fmin(fmax(0.0, T), 3.0 * SI);
    else           YP[N - 1] = fmax(fmin(0.0, T), 3.0 * SI);
    return 0;
}

/* tension_module.f90:314-782  SIGS as modified in LTRANS (TOL = 0, zeroed
 * SIGMA on entry, NIT cap 10000 -> SigErr).  Ledger 18: logical CONT is not
 * assigned on the SIG <= 0.5 secant branch (undefined in the reference; TSPACK's
 * original always evaluates F there); restated as .TRUE. at the start of every
 * interval and carried between that interval's iterations, which keeps intervals
 * independent of one another.                                                 */
static void sigs(int N, const double* X, const double* Y, const double* YP, double* SIGMA,
                 int* IER, int* SigErr)
{
    const double SBIG = 85.0;
    int NM1 = N - 1;
    if (NM1 < 1) { *IER = -1; return; }
    for (int I = 0; I < NM1; ++I) SIGMA[I] = 0.0;          /* :411-414 */
    const double FTOL = 0.0;
    volatile double one_plus;                               /* STORE() :1256 */
    double RTOL = 1.0;
    for (;;) { RTOL = RTOL / 2.0; one_plus = RTOL + 1.0; if (one_plus <= 1.0) break; }
    RTOL = RTOL * 200.0;                                    /* :433-439 */
    int ICNT = 0; double DSM = 0.0;
    int FLAG = 0;
#define STORE_SIG() do { SIG = fmin(SIG, SBIG); if (SIG > SIGIN) { SIGMA[I - 1] = SIG; ICNT++; \
        DSIG = SIG - SIGIN; if (SIGIN > 0.0) DSIG = DSIG / SIGIN; DSM = fmax(DSM, DSIG); } } while (0)
    for (int I = 1; I <= NM1; ++I) {
        int IP1 = I + 1;
        double DX = X[IP1 - 1] - X[I - 1];
        if (DX <= 0.0) { *IER = -IP1; (void)DSM; return; }
        double SIGIN = SIGMA[I - 1];
        if (SIGIN >= SBIG) continue;
        int CONT = 1;                                       /* ledger 18: .TRUE. per interval */
        double A = 0.0, E = 0.0;
        double S1 = YP[I - 1], S2 = YP[IP1 - 1];
        double S = (Y[IP1 - 1] - Y[I - 1]) / DX;
        double D1 = S - S1, D2 = S2 - S, D1D2 = D1 * D2;
        double SIG = SBIG, DSIG;
        if ((D1D2 == 0.0 && S1 != S2) || (S == 0.0 && S1 * S2 > 0.0)) { STORE_SIG(); continue; }
        SIG = 0.0;
        if (D1D2 >= 0.0) {                                  /* convexity */
            if (D1D2 == 0.0) { STORE_SIG(); continue; }
            double T = fmax(D1 / D2, D2 / D1);
            if (T <= 2.0) { STORE_SIG(); continue; }
            double TP1 = T + 1.0;
            SIG = sqrt(10.0 * T - 20.0);
            int NIT = 0;
            for (;;) {
                double T1, FP, F;
                if (SIG <= .5) {
                    double SINHM, COSHM, COSHMM;
                    ora_snhcsh(SIG, &SINHM, &COSHM, &COSHMM);
                    T1 = COSHM / SINHM;
                    FP = T1 + SIG * (SIG / SINHM - T1 * T1 + 1.0);
                } else {
                    double EMS = exp(-SIG);
                    double SSM = 1.0 - EMS * (EMS + SIG + SIG);
                    T1 = (1.0 - EMS) * (1.0 - EMS) / SSM;
                    FP = T1 + SIG * (2.0 * SIG * EMS / SSM - T1 * T1 + 1.0);
                }
                F = SIG * T1 - TP1;
                NIT = NIT + 1;
                if (NIT > 10000) { if (getenv("ORA_SIGDBG")) fprintf(stderr, "SIGDBG newton I=%d N=%d SIG=%.17g F=%.17g FP=%.17g T=%.17g D1=%.17g D2=%.17g\n", I, N, SIG, F, FP, T, D1, D2); *SigErr = *SigErr + 1; return; }   /* :556-559 */
                FLAG = 0;
                if (FP <= 0.0) { FLAG = 1; break; }
                DSIG = -F / FP;
                if (fabs(DSIG) <= RTOL * SIG || (F >= 0.0 && F <= FTOL) || fabs(F) <= RTOL) {
                    FLAG = 1; break;
                }
                SIG = SIG + DSIG;
            }
        }
        if (FLAG) { FLAG = 0; STORE_SIG(); continue; }
        /* monotonicity (D1D2 < 0) :638-760 */
        if (S1 * S < 0.0 || S2 * S < 0.0) { STORE_SIG(); continue; }
        double T0 = 3.0 * S - S1 - S2;
        double D0 = T0 * T0 - S1 * S2;
        if (D0 <= 0.0 || S * T0 >= 0.0) { STORE_SIG(); continue; }
        double SGN = copysign(1.0, S);
        SIG = SBIG;
        double FMAX = SGN * (SIG * S - S1 - S2) / (SIG - 2.0);
        if (FMAX <= 0.0) { STORE_SIG(); continue; }
        double STOL = RTOL * SIG;
        double F = FMAX;
        double F0 = SGN * D0 / (3.0 * (D1 - D2));
        double FNEG = F0;
        DSIG = SIG;
        double DMAX = SIG;
        double D1PD2 = D1 + D2;
        int NIT = 0;
        for (;;) {
            DSIG = -F * DSIG / (F - F0);
            if (fabs(DSIG) > fabs(DMAX) || DSIG * DMAX > 0.0) { DSIG = DMAX; F0 = FNEG; continue; }
            if (fabs(DSIG) < STOL / 2.0) DSIG = -copysign(STOL / 2.0, DMAX);
            SIG = SIG + DSIG;
            F0 = F;
            double C1, C2;
            if (SIG <= .5) {
                double SINHM, COSHM, COSHMM;
                ora_snhcsh(SIG, &SINHM, &COSHM, &COSHMM);
                C1 = SIG * COSHM * D2 - SINHM * D1PD2;
                C2 = SIG * (SINHM + SIG) * D2 - COSHM * D1PD2;
                A = C2 - C1;
                E = SIG * SINHM - COSHMM - COSHMM;
            } else {
                double EMS = exp(-SIG);
                double EMS2 = EMS + EMS;
                double TM = 1.0 - EMS;
                double SSINH = TM * (1.0 + EMS);
                double SSM = SSINH - SIG * EMS2;
                double SCM = TM * TM;
                C1 = SIG * SCM * D2 - SSM * D1PD2;
                C2 = SIG * SSINH * D2 - SCM * D1PD2;
                F = FMAX;
                CONT = 1;
                if (C1 * (SIG * SCM * D1 - SSM * D1PD2) >= 0.0) CONT = 0;
                if (CONT) A = EMS2 * (SIG * TM * D2 + (TM - SIG) * D1PD2);
                if (A * (C2 + C1) < 0.0) CONT = 0;
                if (CONT) E = SIG * SSINH - SCM - SCM;
            }
            if (CONT) F = (SGN * (E * S2 - C2) + sqrt(A * (C2 + C1))) / E;
            NIT = NIT + 1;
            if (NIT > 100000) { if (getenv("ORA_SIGDBG")) fprintf(stderr, "SIGDBG secant I=%d N=%d SIG=%.17g F=%.17g DMAX=%.17g STOL=%.17g\n", I, N, SIG, F, DMAX, STOL); *SigErr = *SigErr + 1; return; }  /* safety net, not in reference */
            STOL = RTOL * SIG;
            if (fabs(DMAX) <= STOL || (F >= 0.0 && F <= FTOL) || fabs(F) <= RTOL) break;
            DMAX = DMAX + DSIG;
            if (F0 * F > 0.0 && fabs(F) >= fabs(F0)) { DSIG = DMAX; F0 = FNEG; continue; }
            if (F0 * F <= 0.0) {
                double T1 = DMAX, T2 = FNEG;
                DMAX = DSIG;
                FNEG = F0;
                if (fabs(DSIG) > fabs(T1) && fabs(F) < fabs(T2)) { DSIG = T1; F0 = T2; }
            }
        }
        STORE_SIG();
    }
#undef STORE_SIG
    (void)ICNT;
}

/* tension_module.f90:231-312  TSPSI                                          */
void ora_tspsi(int32_t n, const double* x, const double* y, double* yp, double* sigma,
               int32_t* ier, int32_t* sigerr)
{
    *ier = 0;
    if (n < 2) { *ier = -1; return; }
    int ierr = ypc1(n, x, y, yp);
    if (ierr != 0) { *ier = -4; return; }
    int e2 = 0, se = *sigerr;
    sigs(n, x, y, yp, sigma, &e2, &se);
    *sigerr = se;
}

/* tension_module.f90:1287-1354  INTRVL (stateless: ledger 20)                */
static int intrvl(double T, int N, const double* X)
{
    int IL = 1, IH = N;
    for (;;) {
        if (IH <= IL + 1) break;
        int K = (IL + IH) / 2;
        if (T < X[K - 1]) IH = K; else IL = K;
    }
    return IL;
}

/* tension_module.f90:980-1122  HVAL                                          */
double ora_hval(double T, int32_t N, const double* X, const double* Y, const double* YP,
                const double* SIGMA, int32_t* IER)
{
    const double SBIG = 85.0;
    int I;
    if (N < 2) { *IER = -1; return 0.0; }
    if (T < X[0]) { I = 1; *IER = 1; }
    else if (T > X[N - 1]) { I = N - 1; *IER = 1; }
    else { I = intrvl(T, N, X); *IER = 0; }
    int IP1 = I + 1;
    double DX = X[IP1 - 1] - X[I - 1];
    if (DX <= 0.0) { *IER = -2; return 0.0; }
    double U = T - X[I - 1];
    double B2 = U / DX, B1 = 1.0 - B2;
    double Y1 = Y[I - 1], S1 = YP[I - 1];
    double S = (Y[IP1 - 1] - Y1) / DX;
    double D1 = S - S1, D2 = YP[IP1 - 1] - S;
    double SIG = fabs(SIGMA[I - 1]);
    if (SIG < 1.e-9) {
        return Y1 + U * (S1 + B2 * (D1 + B1 * (D1 - D2)));
    } else if (SIG <= .5) {
        double SB2 = SIG * B2, SM, CM, CMM, SM2, CM2, DUMMY;
        ora_snhcsh(SIG, &SM, &CM, &CMM);
        ora_snhcsh(SB2, &SM2, &CM2, &DUMMY);
        double E = SIG * SM - CMM - CMM;
        return Y1 + S1 * U + DX * ((CM * SM2 - SM * CM2) * (D1 + D2) +
               SIG * (CM * CM2 - (SM + SIG) * SM2) * D1) / (SIG * E);
    } else {
        double SB1 = SIG * B1, SB2 = SIG - SB1;
        if (-SB1 > SBIG || -SB2 > SBIG) return Y1 + S * U;
        double E1 = exp(-SB1), E2 = exp(-SB2);
        double EMS = E1 * E2, TM = 1.0 - EMS, TS = TM * TM, TP = 1.0 + EMS;
        double E = TM * (SIG * TP - TM - TM);
        return Y1 + S * U + DX * (TM * (TP - E1 - E2) * (D1 + D2) +
               SIG * ((E2 + EMS * (E1 - 2.0) - B1 * TS) * D1 +
                      (E1 + EMS * (E2 - 2.0) - B2 * TS) * D2)) / (SIG * E);
    }
}

/* tension_module.f90:1124-1254  HPVAL                                        */
double ora_hpval(double T, int32_t N, const double* X, const double* Y, const double* YP,
                 const double* SIGMA, int32_t* IER)
{
    const double SBIG = 85.0;
    int I;
    if (N < 2) { *IER = -1; return 0.0; }
    if (T < X[0]) { I = 1; *IER = 1; }
    else if (T > X[N - 1]) { I = N - 1; *IER = 1; }
    else { I = intrvl(T, N, X); *IER = 0; }
    int IP1 = I + 1;
    double DX = X[IP1 - 1] - X[I - 1];
    if (DX <= 0.0) { *IER = -2; return 0.0; }
    double B1 = (X[IP1 - 1] - T) / DX, B2 = 1.0 - B1;
    double S1 = YP[I - 1];
    double S = (Y[IP1 - 1] - Y[I - 1]) / DX;
    double D1 = S - S1, D2 = YP[IP1 - 1] - S;
    double SIG = fabs(SIGMA[I - 1]);
    if (SIG < 1.e-9) {
        return S1 + B2 * (D1 + D2 - 3.0 * B1 * (D2 - D1));
    } else if (SIG <= .5) {
        double SB2 = SIG * B2, SM, CM, CMM, SM2, CM2, DUMMY;
        ora_snhcsh(SIG, &SM, &CM, &CMM);
        ora_snhcsh(SB2, &SM2, &CM2, &DUMMY);
        double SINH2 = SM2 + SB2;
        double E = SIG * SM - CMM - CMM;
        return S1 + ((CM * CM2 - SM * SINH2) * (D1 + D2) +
                     SIG * (CM * SINH2 - (SM + SIG) * CM2) * D1) / E;
    } else {
        double SB1 = SIG * B1, SB2 = SIG - SB1;
        if (-SB1 > SBIG || -SB2 > SBIG) return S;
        double E1 = exp(-SB1), E2 = exp(-SB2);
        double EMS = E1 * E2, TM = 1.0 - EMS;
        double E = TM * (SIG * (1.0 + EMS) - TM - TM);
        return S + (TM * ((E2 - E1) * (D1 + D2) + TM * (D1 - D2)) +
                    SIG * ((E1 * EMS - E2) * D1 + (E1 - E2 * EMS) * D2)) / E;
    }
}

/* ------------------------------------------------------------------------ */
/* hydrodynamic_module.f90:2691-2777  getSlevel / getWlevel (same formula on
 * SC,CS or SCW,CSW).  hc is REAL(4) widened (ledger 4); depth is negative.   */
double ora_slevel(double zeta, double depth, double sc, double cs, float hc_f, int32_t vtransform)
{
    double hc = (double)hc_f;
    double h = -1.0 * depth, S;
    switch (vtransform) {
    case 1: S = hc * sc + (h - hc) * cs; return S + zeta * (1.0 + S / h);
    case 2: S = (hc * sc + h * cs) / (hc + h); return zeta + (zeta + h) * S;
    case 3: return zeta * (1.0 + sc) + hc * sc + (h - hc) * cs;
    default: return NAN;
    }
}

/* ------------------------------------------------------------------------ */
/* random_module.f90:108-246  MT19937 (KAT: init_genrand(5489) -> 3499211612;
 * init_by_array({0x123,0x234,0x345,0x456}) -> 1067595299, 955945823, ...).
 * The Fortran keeps words in signed INTEGER(4); unsigned here, identical bits. */
#define MT_N 624
#define MT_M 397
static uint32_t mt[MT_N];
static int mti = MT_N + 1;
static int mt_initialized = 0;

void ora_mt_init_genrand(uint32_t s)
{
    mt[0] = s;
    for (mti = 1; mti < MT_N; ++mti)
        mt[mti] = 1812433253u * (mt[mti - 1] ^ (mt[mti - 1] >> 30)) + (uint32_t)mti;
    mt_initialized = 1;
}

void ora_mt_init_by_array(const uint32_t* key, int32_t len)
{
    ora_mt_init_genrand(19650218u);
    int i = 1, j = 0;
    int k = MT_N > len ? MT_N : len;
    for (; k; --k) {
        mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
        i++; j++;
        if (i >= MT_N) { mt[0] = mt[MT_N - 1]; i = 1; }
        if (j >= len) j = 0;
    }
    for (k = MT_N - 1; k; --k) {
        mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
        i++;
        if (i >= MT_N) { mt[0] = mt[MT_N - 1]; i = 1; }
    }
    mt[0] = 0x80000000u;
}

uint32_t ora_mt_int32(void)
{
    static const uint32_t mag01[2] = {0u, 0x9908b0dfu};
    uint32_t y;
    if (!mt_initialized) ora_mt_init_genrand(21641u);      /* :171-173 */
    if (mti >= MT_N) {
        int kk;
        for (kk = 0; kk < MT_N - MT_M; ++kk) {
            y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
            mt[kk] = mt[kk + MT_M] ^ (y >> 1) ^ mag01[y & 1u];
        }
        for (; kk < MT_N - 1; ++kk) {
            y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
            mt[kk] = mt[kk + (MT_M - MT_N)] ^ (y >> 1) ^ mag01[y & 1u];
        }
        y = (mt[MT_N - 1] & 0x80000000u) | (mt[0] & 0x7fffffffu);
        mt[MT_N - 1] = mt[MT_M - 1] ^ (y >> 1) ^ mag01[y & 1u];
        mti = 0;
    }
    y = mt[mti++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}
/* :213-220 genrand_real1 in [0,1];  :239-246 genrand_real3 in (0,1) */
double ora_mt_real1(void) { return (double)ora_mt_int32() / 4294967295.0; }
double ora_mt_real3(void) { return ((double)ora_mt_int32() + 0.5) / 4294967296.0; }

/* Philox4x32-10 (Salmon et al., SC'11), the device's counter-based stream.   */
void ora_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
