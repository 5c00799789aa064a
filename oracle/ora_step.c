/*
 * ora_step.c -- CPU restatement of update_particles / find_currents and the
 * hydrodynamic, boundary, turbulence, behaviour and settlement routines they call.
 *
 * TEST INFRASTRUCTURE ONLY (see ltrans_oracle.h).  PARITY UNPINNED BY THE
 * REFERENCE.  Citations are file:line relative to /root/reference/Model/.
 * The particle loop body is deliberately a literal walk through
 * LTRANS.f90:778-1400 in the reference's order, recomputing what it recomputes.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "ltrans_oracle.h"

/* float32 literals widened to double (ledger 1-3) */
#define F32(x) ((double)(float)(x))

enum { FLD_U = 0, FLD_V, FLD_W, FLD_KH, FLD_SALT, FLD_TEMP };

struct ora_ctx {
    ltgpu_params prm;
    int rng_mode, nthreads;
    /* grid (hydrodynamic_module.f90:23-69) */
    int vi, uj, ui, vj, rho_nodes, u_nodes, v_nodes, nRE, nUE, nVE;
    double *rx, *ry, *ux, *uy, *vx, *vy, *depth, *angle;
    int32_t *rho_mask, *u_mask, *v_mask;
    double *SC, *CS, *SCW, *CSW;
    int32_t *RE, *UE, *VE;            /* (4,n) 1-based node ids */
    int32_t *rAdj, *uAdj, *vAdj;      /* (n,10) column-major, 1-based, 0 = none */
    double *r_ele_x, *r_ele_y, *u_ele_x, *u_ele_y, *v_ele_x, *v_ele_y;   /* (4,n) */
    /* bounds (boundary_module.f90:39-50) */
    int nbounds, maxbound, maxisland;
    double *bnd_x, *bnd_y; int32_t *land;
    double *bx, *by, *hx, *hy; int32_t *hid;
    /* habitat (settlement_module.f90:29-33) */
    int pedges, hedges, npoly, nhole;
    double *polys, *holes;
    int32_t *poly_id, *poly_start, *poly_size, *hole_id, *hole_start, *hole_size;
    double *poly_maxdis, *hole_maxdis;
    int32_t *elepoly_ptr, *elepoly_idx, *polyhole_ptr, *polyhole_idx;
    /* fields t_*(3,node,level): time slot fastest (hydro:43-45,682-689) */
    double *t_zeta, *t_salt, *t_temp, *t_Wvel, *t_KH, *t_Uvel, *t_Vvel;
    int t_b, t_c, t_f, npushed, have_salt;
    double *pend_zeta, *pend_salt, *pend_temp, *pend_W, *pend_KH, *pend_U, *pend_V;
    int pending;
    /* particles (LTRANS.f90:104-122, behavior_module.f90:61-86, settlement:30-32) */
    int n; int64_t first_id;
    double *X, *Y, *Z, *nX, *nY, *nZ, *DOB, *Age, *Lifespan, *P_Salt, *P_Temp;
    int32_t *startpoly, *endpoly, *hitBottom, *hitLand, *r_ele, *u_ele, *v_ele;
    int32_t *nsig;                    /* diagnostic: SigErr fall-backs per particle */
    double *timer, *P_Sprev, *P_zprev, *P_swim3;
    int32_t *P_behave; uint8_t *bottom, *dead, *oob, *settle;
    /* events / stop */
    ltgpu_event* ev; int nev, capev; int bad_particle;
    double last_ix3;
    /* test hooks, NULL / 0 in every run (ora_behave_case): scripted random words, salinity at the particle */
    const uint32_t* script; int script_pos; const double* script_PS;
};

typedef struct {            /* module globals set by setEle / setInterp */
    int rnode[4], unode[4], vnode[4];
    double t, u, Wgt[4]; int tOK;
} elestate;

typedef struct { ora_ctx* c; uint32_t id_lo, id_hi, step; } prng;

/* ---- helpers ------------------------------------------------------------- */
static void* dup_(const void* p, size_t bytes)
{
    if (!p || !bytes) return NULL;
    void* q = malloc(bytes); memcpy(q, p, bytes); return q;
}
#define DUP(dst, src, count) do { free(dst); (dst) = dup_((src), sizeof(*(src)) * (size_t)(count)); } while (0)

int32_t ora_create(const ltgpu_params* prm, ora_ctx** out)
{
    if (!prm || !out) return LTGPU_E_ARG;
    ora_ctx* c = (ora_ctx*)calloc(1, sizeof(ora_ctx));
    c->prm = *prm; c->rng_mode = ORA_RNG_PHILOX; c->nthreads = 1;
    c->t_b = 1; c->t_c = 2; c->t_f = 3;
    *out = c; return LTGPU_OK;
}
int32_t ora_set_rng(ora_ctx* c, int32_t mode)
{
    c->rng_mode = mode;
    if (mode == ORA_RNG_MT) ora_mt_init_genrand((uint32_t)c->prm.seed);   /* LTRANS.f90:241 */
    return LTGPU_OK;
}
int32_t ora_set_threads(ora_ctx* c, int32_t nt) { c->nthreads = nt < 1 ? 1 : nt; return LTGPU_OK; }

int32_t ora_destroy(ora_ctx* c)
{
    if (!c) return LTGPU_OK;
    void* ptrs[] = { c->rx, c->ry, c->ux, c->uy, c->vx, c->vy, c->depth, c->angle, c->rho_mask, c->u_mask,
        c->v_mask, c->SC, c->CS, c->SCW, c->CSW, c->RE, c->UE, c->VE, c->rAdj, c->uAdj, c->vAdj,
        c->r_ele_x, c->r_ele_y, c->u_ele_x, c->u_ele_y, c->v_ele_x, c->v_ele_y, c->bnd_x, c->bnd_y,
        c->land, c->bx, c->by, c->hx, c->hy, c->hid, c->polys, c->holes, c->poly_id, c->poly_start,
        c->poly_size, c->hole_id, c->hole_start, c->hole_size, c->poly_maxdis, c->hole_maxdis,
        c->elepoly_ptr, c->elepoly_idx, c->polyhole_ptr, c->polyhole_idx, c->t_zeta, c->t_salt,
        c->t_temp, c->t_Wvel, c->t_KH, c->t_Uvel, c->t_Vvel, c->pend_zeta, c->pend_salt, c->pend_temp,
        c->pend_W, c->pend_KH, c->pend_U, c->pend_V, c->X, c->Y, c->Z, c->nX, c->nY, c->nZ, c->DOB,
        c->Age, c->Lifespan, c->P_Salt, c->P_Temp, c->startpoly, c->endpoly, c->hitBottom, c->hitLand,
        c->r_ele, c->u_ele, c->v_ele, c->timer, c->P_Sprev, c->P_zprev, c->P_swim3, c->P_behave,
        c->bottom, c->dead, c->oob, c->settle, c->ev };
    for (size_t i = 0; i < sizeof(ptrs) / sizeof(ptrs[0]); ++i) free(ptrs[i]);
    free(c); return LTGPU_OK;
}

static void ele_xy(int n, const int32_t* E, const double* nx, const double* ny, double** ox, double** oy)
{   /* hydro:561-580 */
    free(*ox); free(*oy);
    *ox = (double*)malloc(sizeof(double) * 4 * (size_t)n);
    *oy = (double*)malloc(sizeof(double) * 4 * (size_t)n);
    for (int j = 0; j < n; ++j) for (int i = 0; i < 4; ++i) {
        (*ox)[4 * j + i] = nx[E[4 * j + i] - 1];
        (*oy)[4 * j + i] = ny[E[4 * j + i] - 1];
    }
}

int32_t ora_set_grid(ora_ctx* c, int32_t vi, int32_t uj, int32_t ui, int32_t vj,
    const double* rx, const double* ry, const double* ux, const double* uy,
    const double* vx, const double* vy, const double* depth, const double* angle,
    const int32_t* rho_mask, const int32_t* u_mask, const int32_t* v_mask,
    const double* SC, const double* CS, const double* SCW, const double* CSW,
    const int32_t* RE, const int32_t* UE, const int32_t* VE,
    int32_t nRE, int32_t nUE, int32_t nVE,
    const int32_t* rAdj, const int32_t* uAdj, const int32_t* vAdj)
{
    c->vi = vi; c->uj = uj; c->ui = ui; c->vj = vj;
    c->rho_nodes = vi * uj; c->u_nodes = ui * uj; c->v_nodes = vi * vj;
    c->nRE = nRE; c->nUE = nUE; c->nVE = nVE;
    int us = c->prm.us, ws = c->prm.ws;
    DUP(c->rx, rx, c->rho_nodes); DUP(c->ry, ry, c->rho_nodes);
    DUP(c->ux, ux, c->u_nodes);   DUP(c->uy, uy, c->u_nodes);
    DUP(c->vx, vx, c->v_nodes);   DUP(c->vy, vy, c->v_nodes);
    DUP(c->depth, depth, c->rho_nodes); DUP(c->angle, angle, c->rho_nodes);
    DUP(c->rho_mask, rho_mask, c->rho_nodes); DUP(c->u_mask, u_mask, c->u_nodes);
    DUP(c->v_mask, v_mask, c->v_nodes);
    DUP(c->SC, SC, us); DUP(c->CS, CS, us); DUP(c->SCW, SCW, ws); DUP(c->CSW, CSW, ws);
    DUP(c->RE, RE, 4 * nRE); DUP(c->UE, UE, 4 * nUE); DUP(c->VE, VE, 4 * nVE);
    DUP(c->rAdj, rAdj, 10 * nRE); DUP(c->uAdj, uAdj, 10 * nUE); DUP(c->vAdj, vAdj, 10 * nVE);
    ele_xy(nRE, c->RE, c->rx, c->ry, &c->r_ele_x, &c->r_ele_y);
    ele_xy(nUE, c->UE, c->ux, c->uy, &c->u_ele_x, &c->u_ele_y);
    ele_xy(nVE, c->VE, c->vx, c->vy, &c->v_ele_x, &c->v_ele_y);
    size_t r3 = 3 * (size_t)c->rho_nodes;
    free(c->t_zeta); free(c->t_salt); free(c->t_temp); free(c->t_Wvel); free(c->t_KH);
    free(c->t_Uvel); free(c->t_Vvel);
    c->t_zeta = (double*)calloc(r3, 8);
    c->t_salt = (double*)calloc(r3 * us, 8); c->t_temp = (double*)calloc(r3 * us, 8);
    c->t_Wvel = (double*)calloc(r3 * ws, 8); c->t_KH = (double*)calloc(r3 * ws, 8);
    c->t_Uvel = (double*)calloc(3 * (size_t)c->u_nodes * us, 8);
    c->t_Vvel = (double*)calloc(3 * (size_t)c->v_nodes * us, 8);
    c->pend_zeta = (double*)calloc(c->rho_nodes, 8);
    c->pend_salt = (double*)calloc((size_t)c->rho_nodes * us, 8);
    c->pend_temp = (double*)calloc((size_t)c->rho_nodes * us, 8);
    c->pend_W = (double*)calloc((size_t)c->rho_nodes * ws, 8);
    c->pend_KH = (double*)calloc((size_t)c->rho_nodes * ws, 8);
    c->pend_U = (double*)calloc((size_t)c->u_nodes * us, 8);
    c->pend_V = (double*)calloc((size_t)c->v_nodes * us, 8);
    return LTGPU_OK;
}

int32_t ora_set_bounds(ora_ctx* c, int32_t nbounds, const double* bnd_x, const double* bnd_y,
    const int32_t* land, int32_t maxbound, const double* bx, const double* by,
    int32_t maxisland, const double* hx, const double* hy, const int32_t* hid)
{
    c->nbounds = nbounds; c->maxbound = maxbound; c->maxisland = maxisland;
    DUP(c->bnd_x, bnd_x, 2 * nbounds); DUP(c->bnd_y, bnd_y, 2 * nbounds); DUP(c->land, land, nbounds);
    DUP(c->bx, bx, maxbound); DUP(c->by, by, maxbound);
    DUP(c->hx, hx, maxisland); DUP(c->hy, hy, maxisland); DUP(c->hid, hid, maxisland);
    return LTGPU_OK;
}

int32_t ora_set_habitat(ora_ctx* c, int32_t pedges, const double* polys, int32_t hedges,
    const double* holes, int32_t npoly, const int32_t* poly_id, const int32_t* poly_start,
    const int32_t* poly_size, const double* poly_maxdis, int32_t nhole, const int32_t* hole_id,
    const int32_t* hole_start, const int32_t* hole_size, const double* hole_maxdis,
    const int32_t* elepoly_ptr, const int32_t* elepoly_idx,
    const int32_t* polyhole_ptr, const int32_t* polyhole_idx)
{
    c->pedges = pedges; c->hedges = hedges; c->npoly = npoly; c->nhole = nhole;
    DUP(c->polys, polys, 5 * pedges); DUP(c->holes, holes, 6 * hedges);
    DUP(c->poly_id, poly_id, npoly); DUP(c->poly_start, poly_start, npoly);
    DUP(c->poly_size, poly_size, npoly); DUP(c->poly_maxdis, poly_maxdis, npoly);
    DUP(c->hole_id, hole_id, nhole); DUP(c->hole_start, hole_start, nhole);
    DUP(c->hole_size, hole_size, nhole); DUP(c->hole_maxdis, hole_maxdis, nhole);
    DUP(c->elepoly_ptr, elepoly_ptr, c->nRE + 1);
    DUP(c->elepoly_idx, elepoly_idx, elepoly_ptr[c->nRE]);
    DUP(c->polyhole_ptr, polyhole_ptr, npoly + 1);
    DUP(c->polyhole_idx, polyhole_idx, polyhole_ptr[npoly]);
    return LTGPU_OK;
}

/* whole-grid form of gridcell (gridcell_module.f90:26-257 without checkele):
 * scan stops at the first definite hit; an on-edge hit found in step 6 whose
 * crossing total is even sets P_ele but keeps scanning (the inner `exit`).  */
static int gridcell_scan(int n, const double* ex, const double* ey, double X, double Y, int* P_ele)
{
    int triangle = 0;
    for (int i = 0; i < n; ++i) {
        const double* qx = ex + 4 * i; const double* qy = ey + 4 * i;
        if (!ora_gridcell(qx, qy, X, Y)) continue;
        triangle = 1; *P_ele = i + 1;
        /* decide whether this was a definite hit (loop exit) or a soft one */
        int soft = 0;
        /* steps 3,4 return definite; step 6: soft iff on an edge and total even */
        int onnode = (X == qx[0] && Y == qy[0]) || (X == qx[1] && Y == qy[1]) ||
                     (X == qx[2] && Y == qy[2]) || (X == qx[3] && Y == qy[3]);
        if (!onnode) {
            int horiz = 0;
            for (int a = 0; a < 4; ++a) for (int b = a + 1; b < 4; ++b)
                if (qy[a] == qy[b] && Y == qy[a]) horiz = 1;
            if (!horiz) {
                int counter[4] = {0, 0, 0, 0}, onedge = 0;
                for (int p = 0; p < 4 && !onedge; ++p) {
                    double bx1 = qx[p], by1 = qy[p], bx2 = qx[(p + 1) & 3], by2 = qy[(p + 1) & 3];
                    if (X <= bx1 || X <= bx2) {
                        if ((by1 > by2 && Y >= by2 && Y <= by1) || (by2 > by1 && Y >= by1 && Y <= by2)) {
                            if (bx1 == bx2) {
                                if (X == bx1) { onedge = 1; break; }
                                counter[p] = 1; if (Y == by2) counter[p] = 0;
                            } else {
                                double slope = (by1 - by2) / (bx1 - bx2);
                                double xi = (Y - by1 + (slope * bx1)) / slope;
                                if (xi > X) { counter[p] = 1; if (Y == by2) counter[p] = 0; }
                                if (xi == X) { onedge = 1; break; }
                            }
                        }
                    }
                }
                int total = counter[0] + counter[1] + counter[2] + counter[3];
                if (onedge && (total % 2) == 0) soft = 1;
            }
        }
        if (!soft) break;
    }
    return triangle;
}

int32_t ora_set_particles(ora_ctx* c, int32_t n, int64_t first_id,
    const double* x, const double* y, const double* z, const double* dob,
    const int32_t* startpoly, const int32_t* r_ele, const int32_t* u_ele, const int32_t* v_ele)
{
    c->n = n; c->first_id = first_id;
    DUP(c->X, x, n); DUP(c->Y, y, n); DUP(c->Z, z, n);
    DUP(c->nX, x, n); DUP(c->nY, y, n); DUP(c->nZ, z, n);      /* LTRANS.f90:262-264 */
    DUP(c->DOB, dob, n);
#define ZALLOC(p, T) do { free(p); (p) = (T*)calloc((size_t)n, sizeof(T)); } while (0)
    ZALLOC(c->Age, double); ZALLOC(c->Lifespan, double); ZALLOC(c->P_Salt, double); ZALLOC(c->P_Temp, double);
    ZALLOC(c->startpoly, int32_t); ZALLOC(c->endpoly, int32_t); ZALLOC(c->hitBottom, int32_t);
    ZALLOC(c->nsig, int32_t);
    ZALLOC(c->hitLand, int32_t); ZALLOC(c->r_ele, int32_t); ZALLOC(c->u_ele, int32_t); ZALLOC(c->v_ele, int32_t);
    ZALLOC(c->timer, double); ZALLOC(c->P_Sprev, double); ZALLOC(c->P_zprev, double); ZALLOC(c->P_swim3, double);
    ZALLOC(c->P_behave, int32_t); ZALLOC(c->bottom, uint8_t); ZALLOC(c->dead, uint8_t);
    ZALLOC(c->oob, uint8_t); ZALLOC(c->settle, uint8_t);
#undef ZALLOC
    if (startpoly) memcpy(c->startpoly, startpoly, sizeof(int32_t) * (size_t)n);
    for (int i = 0; i < n; ++i) {
        c->P_behave[i] = c->prm.Behavior;                       /* behavior:118 */
        c->bottom[i] = 1;                                       /* behavior:113 */
        if (r_ele && u_ele && v_ele) {
            c->r_ele[i] = r_ele[i]; c->u_ele[i] = u_ele[i]; c->v_ele[i] = v_ele[i];
        } else {                                                /* hydro:1436-1457 */
            int e = 0;
            gridcell_scan(c->nRE, c->r_ele_x, c->r_ele_y, x[i], y[i], &e); c->r_ele[i] = e; e = 0;
            gridcell_scan(c->nUE, c->u_ele_x, c->u_ele_y, x[i], y[i], &e); c->u_ele[i] = e; e = 0;
            gridcell_scan(c->nVE, c->v_ele_x, c->v_ele_y, x[i], y[i], &e); c->v_ele[i] = e;
        }
    }
    return LTGPU_OK;
}

/* ---- hydro records (hydro:991-1043 initial, :1371-1403 update) ----------- */
static double rd(const void* p, int dtype, size_t i)
{
    return dtype == LTGPU_F32 ? (double)((const float*)p)[i] : ((const double*)p)[i];
}
static void put_slot(ora_ctx* c, int slot, int dtype, const void* zeta, const void* u, const void* v,
                     const void* w, const void* aks, const void* salt, const void* temp)
{
    int us = c->prm.us, ws = c->prm.ws;
    size_t rn = (size_t)c->rho_nodes, un = (size_t)c->u_nodes, vn = (size_t)c->v_nodes;
    int s = slot - 1;
    for (size_t nd = 0; nd < rn; ++nd) {
        double m = (double)c->rho_mask[nd];
        c->t_zeta[3 * nd + s] = rd(zeta, dtype, nd) * m;
        for (int k = 0; k < ws; ++k) {
            c->t_Wvel[3 * (nd + rn * k) + s] = rd(w, dtype, nd + rn * k) * m;
            c->t_KH[3 * (nd + rn * k) + s] = rd(aks, dtype, nd + rn * k) * m;
        }
        if (salt && temp) for (int k = 0; k < us; ++k) {
            c->t_salt[3 * (nd + rn * k) + s] = rd(salt, dtype, nd + rn * k) * m;
            c->t_temp[3 * (nd + rn * k) + s] = rd(temp, dtype, nd + rn * k) * m;
        }
    }
    for (size_t nd = 0; nd < un; ++nd) for (int k = 0; k < us; ++k)
        c->t_Uvel[3 * (nd + un * k) + s] = rd(u, dtype, nd + un * k) * (double)c->u_mask[nd];
    for (size_t nd = 0; nd < vn; ++nd) for (int k = 0; k < us; ++k)
        c->t_Vvel[3 * (nd + vn * k) + s] = rd(v, dtype, nd + vn * k) * (double)c->v_mask[nd];
}

int32_t ora_push_hydro(ora_ctx* c, int32_t dtype, const void* zeta, const void* u, const void* v,
                       const void* w, const void* aks, const void* salt, const void* temp)
{
    if (!c->t_zeta) return LTGPU_E_ARG;
    if (c->npushed < 3) {
        put_slot(c, c->npushed + 1, dtype, zeta, u, v, w, aks, salt, temp);
        c->have_salt = (salt && temp);
        c->npushed++;
        return LTGPU_OK;
    }
    if (c->pending) return LTGPU_E_ARG;
    int us = c->prm.us, ws = c->prm.ws;
    size_t rn = (size_t)c->rho_nodes, un = (size_t)c->u_nodes, vn = (size_t)c->v_nodes;
    for (size_t i = 0; i < rn; ++i) c->pend_zeta[i] = rd(zeta, dtype, i);
    for (size_t i = 0; i < rn * ws; ++i) { c->pend_W[i] = rd(w, dtype, i); c->pend_KH[i] = rd(aks, dtype, i); }
    if (salt && temp) for (size_t i = 0; i < rn * us; ++i) {
        c->pend_salt[i] = rd(salt, dtype, i); c->pend_temp[i] = rd(temp, dtype, i); }
    for (size_t i = 0; i < un * us; ++i) c->pend_U[i] = rd(u, dtype, i);
    for (size_t i = 0; i < vn * us; ++i) c->pend_V[i] = rd(v, dtype, i);
    c->pending = 1; c->npushed++;
    return LTGPU_OK;
}

int32_t ora_rotate_hydro(ora_ctx* c)
{
    if (!c->pending) return LTGPU_E_ARG;
    c->t_b = c->t_b % 3 + 1; c->t_c = c->t_c % 3 + 1; c->t_f = c->t_f % 3 + 1;   /* hydro:1080-1082 */
    put_slot(c, c->t_f, LTGPU_F64, c->pend_zeta, c->pend_U, c->pend_V, c->pend_W, c->pend_KH,
             c->have_salt ? c->pend_salt : NULL, c->have_salt ? c->pend_temp : NULL);
    c->pending = 0;
    return LTGPU_OK;
}

/* ---- RNG ------------------------------------------------------------------ */
/* Stream layout shared with the device (include/ltrans_b200.h):
 *   counter = (id_lo, id_hi, global internal step, block), key = (seed, 0)
 *   block 0                : HTurb, words 0,1 -> devX ; 2,3 -> devY
 *   block 1 + (i >> 1)     : VTurb normal i (0-based), words 2(i&1), 2(i&1)+1
 *   block 0x80000000       : behave, words 0,1,2 in draw order               */
static uint32_t draw_word(prng* g, uint32_t block, int word)
{
    if (g->c->rng_mode == ORA_RNG_MT) return ora_mt_int32();
    if (g->c->script) return g->c->script[g->c->script_pos++];
    uint32_t ctr[4] = { g->id_lo, g->id_hi, g->step, block };
    uint32_t key[2] = { (uint32_t)g->c->prm.seed, 0u }, out[4];
    ora_philox4x32_10(ctr, key, out);
    return out[word];
}
static double real3(prng* g, uint32_t b, int w) { return ((double)draw_word(g, b, w) + 0.5) / 4294967296.0; }
static double real1(prng* g, uint32_t b, int w) { return (double)draw_word(g, b, w) / 4294967295.0; }
/* norm_module.f90:25-39 */
static double norm_(prng* g, uint32_t b, int w)
{
    double dev1 = real3(g, b, w), dev2 = real3(g, b, w + 1);
    return sqrt(-2.0 * log(dev1)) * cos(2.0 * g->c->prm.PI * dev2);
}

/* ---- events --------------------------------------------------------------- */
static void add_event(ora_ctx* c, int n1, int code, double t)
{
#pragma omp critical(ora_events)
    {
        if (c->nev == c->capev) {
            c->capev = c->capev ? 2 * c->capev : 64;
            c->ev = (ltgpu_event*)realloc(c->ev, sizeof(ltgpu_event) * (size_t)c->capev);
        }
        c->ev[c->nev].particle = n1; c->ev[c->nev].code = code; c->ev[c->nev].time = t; c->nev++;
    }
}

/* ---- setEle (hydro:1414-1532), not-first form ------------------------------ */
static int find_adj(int nE, const int32_t* Adj, const double* ex, const double* ey,
                    double X, double Y, int32_t* P_element, int errcode, int* error)
{
    int oP = *P_element;
    for (int i = 0; i < 10; ++i) {
        int check = Adj[(size_t)i * nE + (oP - 1)];
        if (check == 0) { *error = errcode; break; }     /* ledger 17: stop at first 0 */
        if (ora_gridcell(ex + 4 * (size_t)(check - 1), ey + 4 * (size_t)(check - 1), X, Y)) {
            *P_element = check; return 1;
        }
    }
    return 0;
}
static int setEle(ora_ctx* c, double X, double Y, int n, elestate* es)
{
    int error = 0;
    find_adj(c->nRE, c->rAdj, c->r_ele_x, c->r_ele_y, X, Y, &c->r_ele[n], 4, &error);
    find_adj(c->nUE, c->uAdj, c->u_ele_x, c->u_ele_y, X, Y, &c->u_ele[n], 5, &error);
    find_adj(c->nVE, c->vAdj, c->v_ele_x, c->v_ele_y, X, Y, &c->v_ele[n], 6, &error);
    for (int i = 0; i < 4; ++i) {                          /* :1515-1528 */
        es->rnode[i] = c->RE[4 * (size_t)(c->r_ele[n] - 1) + i];
        es->unode[i] = c->UE[4 * (size_t)(c->u_ele[n] - 1) + i];
        es->vnode[i] = c->VE[4 * (size_t)(c->v_ele[n] - 1) + i];
    }
    return error;
}

/* ---- free-slip corner substitution (hydro:1936-1994, 2330-2519) ------------ */
static void freeslip(double v[4], const int m[4], int one_land_sum, const int md[4])
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

/* ---- setInterp (hydro:1680-1740) ------------------------------------------ */
static void setInterp(ora_ctx* c, elestate* es, double xp, double yp)
{
    double x1 = c->rx[es->rnode[0] - 1], x2 = c->rx[es->rnode[1] - 1],
           x3 = c->rx[es->rnode[2] - 1], x4 = c->rx[es->rnode[3] - 1];
    double y1 = c->ry[es->rnode[0] - 1], y2 = c->ry[es->rnode[1] - 1],
           y3 = c->ry[es->rnode[2] - 1], y4 = c->ry[es->rnode[3] - 1];
    es->Wgt[0] = es->Wgt[1] = es->Wgt[2] = es->Wgt[3] = 0;
    double t = ((xp - x1) * (y3 - y1) + (y1 - yp) * (x3 - x1)) / ((x2 - x1) * (y3 - y1) - (y2 - y1) * (x3 - x1));
    double u = ((xp - x1) * (y2 - y1) + (y1 - yp) * (x2 - x1)) / ((x3 - x1) * (y2 - y1) - (y3 - y1) * (x2 - x1));
    int tOK = 1;
    if (t < 0. || u < 0. || (t + u) > 1.0) {
        t = ((xp - x3) * (y1 - y3) + (y3 - yp) * (x1 - x3)) / ((x4 - x3) * (y1 - y3) - (y4 - y3) * (x1 - x3));
        u = ((xp - x3) * (y4 - y3) + (y3 - yp) * (x4 - x3)) / ((x1 - x3) * (y4 - y3) - (y1 - y3) * (x4 - x3));
        tOK = 2;
        if (t < 0. || u < 0. || (t + u) > 1.0) {
            if ((xp == x1 && yp == y1) || (xp == x2 && yp == y2) || (xp == x3 && yp == y3) || (xp == x4 && yp == y4)) {
                if (xp == x1 && yp == y1) es->Wgt[0] = 1.0;     /* tOK stays 2 (:1719-1724) */
                if (xp == x2 && yp == y2) es->Wgt[1] = 1.0;
                if (xp == x3 && yp == y3) es->Wgt[2] = 1.0;
                if (xp == x4 && yp == y4) es->Wgt[3] = 1.0;
            } else {
                double D1 = 1. / sqrt((x1 - xp) * (x1 - xp) + (y1 - yp) * (y1 - yp));
                double D2 = 1. / sqrt((x2 - xp) * (x2 - xp) + (y2 - yp) * (y2 - yp));
                double D3 = 1. / sqrt((x3 - xp) * (x3 - xp) + (y3 - yp) * (y3 - yp));
                double D4 = 1. / sqrt((x4 - xp) * (x4 - xp) + (y4 - yp) * (y4 - yp));
                double TD = D1 + D2 + D3 + D4;
                es->Wgt[0] = D1 / TD; es->Wgt[1] = D2 / TD; es->Wgt[2] = D3 / TD; es->Wgt[3] = D4 / TD;
                tOK = 3;
            }
        }
    }
    es->t = t; es->u = u; es->tOK = tOK;
}

/* ---- getInterp (hydro:1743-2005): rho-grid value with setInterp's weights --- */
static double combine(const elestate* es, const double v[4])
{
    if (es->tOK == 1) return v[0] + (v[1] - v[0]) * es->t + (v[2] - v[0]) * es->u;
    else if (es->tOK == 2) return v[2] + (v[3] - v[2]) * es->t + (v[0] - v[2]) * es->u;
    return es->Wgt[0] * v[0] + es->Wgt[1] * v[1] + es->Wgt[2] * v[2] + es->Wgt[3] * v[3];
}
static void rho_freeslip(ora_ctx* c, const elestate* es, double v[4])
{
    if (!c->prm.FreeSlip) return;
    int m[4]; for (int i = 0; i < 4; ++i) m[i] = c->rho_mask[es->rnode[i] - 1];
    freeslip(v, m, 3, m);
}
static double getInterp_static(ora_ctx* c, const elestate* es, const double* arr)
{
    double v[4]; for (int i = 0; i < 4; ++i) v[i] = arr[es->rnode[i] - 1];
    rho_freeslip(c, es, v);
    return combine(es, v);
}
static double getInterp_zeta(ora_ctx* c, const elestate* es, int slot)
{
    double v[4]; for (int i = 0; i < 4; ++i) v[i] = c->t_zeta[3 * (size_t)(es->rnode[i] - 1) + (slot - 1)];
    rho_freeslip(c, es, v);
    return combine(es, v);
}
static double getInterp_kh(ora_ctx* c, const elestate* es, int slot, int lev)
{
    double v[4];
    for (int i = 0; i < 4; ++i)
        v[i] = c->t_KH[3 * ((size_t)(es->rnode[i] - 1) + (size_t)c->rho_nodes * (lev - 1)) + (slot - 1)];
    rho_freeslip(c, es, v);
    return combine(es, v);
}

/* ---- interp (hydro:2008-2569): value at (xp,yp) in the element nodes cached
 *      by the last setEle, recomputing the weights every call --------------- */
static double interp(ora_ctx* c, const elestate* es, double xp, double yp, int fld, int slot, int lev)
{
    double v[4], x1, x2, x3, x4, y1, y2, y3, y4;
    int s = slot - 1; size_t k = (size_t)(lev - 1);
    if (fld == FLD_U) {
        const int* nd = es->unode; size_t un = (size_t)c->u_nodes;
        for (int i = 0; i < 4; ++i) v[i] = c->t_Uvel[3 * ((size_t)(nd[i] - 1) + un * k) + s];
        x1 = c->ux[nd[0] - 1]; x2 = c->ux[nd[1] - 1]; x3 = c->ux[nd[2] - 1]; x4 = c->ux[nd[3] - 1];
        y1 = c->uy[nd[0] - 1]; y2 = c->uy[nd[1] - 1]; y3 = c->uy[nd[2] - 1]; y4 = c->uy[nd[3] - 1];
        if (c->prm.FreeSlip) {                 /* ledger 16: one-land test is sum == 1 (hydro:2408) */
            int m[4]; for (int i = 0; i < 4; ++i) m[i] = c->u_mask[nd[i] - 1];
            freeslip(v, m, 1, m);
        }
    } else if (fld == FLD_V) {
        const int* nd = es->vnode; size_t vn = (size_t)c->v_nodes;
        for (int i = 0; i < 4; ++i) v[i] = c->t_Vvel[3 * ((size_t)(nd[i] - 1) + vn * k) + s];
        x1 = c->vx[nd[0] - 1]; x2 = c->vx[nd[1] - 1]; x3 = c->vx[nd[2] - 1]; x4 = c->vx[nd[3] - 1];
        y1 = c->vy[nd[0] - 1]; y2 = c->vy[nd[1] - 1]; y3 = c->vy[nd[2] - 1]; y4 = c->vy[nd[3] - 1];
        if (c->prm.FreeSlip) {                 /* ledger 16: diagonal test reads v_mask(unode*) (hydro:2500-2503) */
            int m[4], md[4];
            for (int i = 0; i < 4; ++i) {
                m[i] = c->v_mask[nd[i] - 1];
                int q = es->unode[i];          /* out-of-range in the reference is undefined: treat as water */
                md[i] = (q >= 1 && q <= c->v_nodes) ? c->v_mask[q - 1] : 1;
            }
            freeslip(v, m, 3, md);
        }
    } else {
        const int* nd = es->rnode; size_t rn = (size_t)c->rho_nodes;
        const double* arr = fld == FLD_W ? c->t_Wvel : fld == FLD_KH ? c->t_KH : fld == FLD_SALT ? c->t_salt : c->t_temp;
        for (int i = 0; i < 4; ++i) v[i] = arr[3 * ((size_t)(nd[i] - 1) + rn * k) + s];
        x1 = c->rx[nd[0] - 1]; x2 = c->rx[nd[1] - 1]; x3 = c->rx[nd[2] - 1]; x4 = c->rx[nd[3] - 1];
        y1 = c->ry[nd[0] - 1]; y2 = c->ry[nd[1] - 1]; y3 = c->ry[nd[2] - 1]; y4 = c->ry[nd[3] - 1];
        rho_freeslip(c, es, v);
    }
    double v1 = v[0], v2 = v[1], v3 = v[2], v4 = v[3];
    double tt = ((xp - x1) * (y3 - y1) + (y1 - yp) * (x3 - x1)) / ((x2 - x1) * (y3 - y1) - (y2 - y1) * (x3 - x1));
    double uu = ((xp - x1) * (y2 - y1) + (y1 - yp) * (x2 - x1)) / ((x3 - x1) * (y2 - y1) - (y3 - y1) * (x2 - x1));
    double vp = v1 + (v2 - v1) * tt + (v3 - v1) * uu;
    if (tt < 0. || uu < 0. || (tt + uu) > 1.0) {
        tt = ((xp - x3) * (y1 - y3) + (y3 - yp) * (x1 - x3)) / ((x4 - x3) * (y1 - y3) - (y4 - y3) * (x1 - x3));
        uu = ((xp - x3) * (y4 - y3) + (y3 - yp) * (x4 - x3)) / ((x1 - x3) * (y4 - y3) - (y1 - y3) * (x4 - x3));
        vp = v3 + (v4 - v3) * tt + (v1 - v3) * uu;
        if (tt < 0. || uu < 0. || (tt + uu) > 1.0) {
            if ((xp == x1 && yp == y1) || (xp == x2 && yp == y2) || (xp == x3 && yp == y3) || (xp == x4 && yp == y4)) {
                if (xp == x1 && yp == y1) vp = v1;
                if (xp == x2 && yp == y2) vp = v2;
                if (xp == x3 && yp == y3) vp = v3;
                if (xp == x4 && yp == y4) vp = v4;
            } else {
                double D1 = 1. / sqrt((x1 - xp) * (x1 - xp) + (y1 - yp) * (y1 - yp));
                double D2 = 1. / sqrt((x2 - xp) * (x2 - xp) + (y2 - yp) * (y2 - yp));
                double D3 = 1. / sqrt((x3 - xp) * (x3 - xp) + (y3 - yp) * (y3 - yp));
                double D4 = 1. / sqrt((x4 - xp) * (x4 - xp) + (y4 - yp) * (y4 - yp));
                double TD = D1 + D2 + D3 + D4;
                vp = (D1 / TD) * v1 + (D2 / TD) * v2 + (D3 / TD) * v3 + (D4 / TD) * v4;
            }
        }
    }
    return vp;
}

/* The two horizontal interpolators on a bare quadrilateral (tests/test_oracle_differential.py): corner coordinates
 * x[4], y[4] and corner values v[4] in the element's node order.  which = 0: setInterp + getInterp (hydro:1680-1740,
 * 1996-2004: weights stored once, an on-node point outside both triangles keeps tOK = 2); 1: interp (hydro:2533-2565:
 * weights per call, an on-node point takes the node value).  Built on the same routines the step uses. */
double ora_interp_quad(const double x[4], const double y[4], const double v[4], double xp, double yp, int32_t which)
{
    ora_ctx c; memset(&c, 0, sizeof c);
    elestate es; memset(&es, 0, sizeof es);
    double rx[4] = {x[0], x[1], x[2], x[3]}, ry[4] = {y[0], y[1], y[2], y[3]};
    double fld[12];
    for (int i = 0; i < 4; ++i) { es.rnode[i] = i + 1; fld[3 * i] = fld[3 * i + 1] = fld[3 * i + 2] = v[i]; }
    c.rx = rx; c.ry = ry; c.rho_nodes = 4; c.t_Wvel = fld;
    if (which == 0) { setInterp(&c, &es, xp, yp); return combine(&es, v); }
    return interp(&c, &es, xp, yp, FLD_W, 1, 1);
}

/* diagnostic only: SigErr (linint) fall-backs that reached a result during the current
 * particle-step; summed per particle for ora_fetch_sigerr */
static __thread int tl_nsig;

/* ---- WCTS_ITPI (hydro:2577-2689) ------------------------------------------ */
static double WCTS_core(const double* abb_zb, const double* abb_zc, const double* abb_zf,
    const double* abb_vb, const double* abb_vc, const double* abb_vf,
    double P_zb, double P_zc, double P_zf, const double ex[3], const double ix[3], int p, int v, int* nfall);
/* Where find_currents / WCTS_ITPI get a field value at the particle: in the step, interp() at (Xpos, Ypos) in the
 * elements of the last setEle; in ora_find_currents_column, bare profiles handed in by the test. when = 0, 1, 2: the
 * back, centre, forward hydro record; lev 1-based. */
typedef double (*valfn)(const void* h, int fld, int when, int lev);
typedef struct { ora_ctx* c; const elestate* es; double X, Y; } stepvals;
static double stepvals_get(const void* h, int fld, int when, int lev)
{
    const stepvals* s = (const stepvals*)h;
    return interp(s->c, s->es, s->X, s->Y, fld, when == 0 ? s->c->t_b : when == 1 ? s->c->t_c : s->c->t_f, lev);
}
static double WCTS_get(valfn get, const void* h, int fld, int deplvl,
    const double* Pwc_zb, const double* Pwc_zc, const double* Pwc_zf,
    double P_zb, double P_zc, double P_zf, const double ex[3], const double ix[3], int p, int v)
{
    enum { nN = 4 };
    double abb_zb[nN], abb_zc[nN], abb_zf[nN], abb_vb[nN], abb_vc[nN], abb_vf[nN];
    for (int i = 1; i <= nN; ++i) {
        abb_zb[i - 1] = Pwc_zb[i + deplvl - 2];
        abb_zc[i - 1] = Pwc_zc[i + deplvl - 2];
        abb_zf[i - 1] = Pwc_zf[i + deplvl - 2];
        abb_vb[i - 1] = get(h, fld, 0, i + deplvl - 1);
        abb_vc[i - 1] = get(h, fld, 1, i + deplvl - 1);
        abb_vf[i - 1] = get(h, fld, 2, i + deplvl - 1);
    }
    int nfall = 0;
    double r = WCTS_core(abb_zb, abb_zc, abb_zf, abb_vb, abb_vc, abb_vf, P_zb, P_zc, P_zf, ex, ix, p, v, &nfall);
    tl_nsig += nfall;
    return r;
}
static double WCTS_ITPI(ora_ctx* c, const elestate* es, int fld, double Xpos, double Ypos, int deplvl,
    const double* Pwc_zb, const double* Pwc_zc, const double* Pwc_zf,
    double P_zb, double P_zc, double P_zf, const double ex[3], const double ix[3], int p, int v)
{
    stepvals sv = { c, es, Xpos, Ypos };
    return WCTS_get(stepvals_get, &sv, fld, deplvl, Pwc_zb, Pwc_zc, Pwc_zf, P_zb, P_zc, P_zf, ex, ix, p, v);
}

/* WCTS_ITPI after the gather of the 4-level profiles (hydro:2619-2689); *nfall = SigErr fall-backs that count */
static double WCTS_core(const double* abb_zb, const double* abb_zc, const double* abb_zf,
    const double* abb_vb, const double* abb_vc, const double* abb_vf,
    double P_zb, double P_zc, double P_zf, const double ex[3], const double ix[3], int p, int v, int* nfall)
{
    enum { nN = 4 };
    double YP[nN], SIGM[nN], slope, P_vb = 0.0, P_vc = 0.0, P_vf = 0.0;
    int IER, SigErr;
    SigErr = 0; ora_tspsi(nN, abb_zb, abb_vb, YP, SIGM, &IER, &SigErr);
    if (SigErr == 0) P_vb = ora_hval(P_zb, nN, abb_zb, abb_vb, YP, SIGM, &IER);
    else { ora_linint(abb_zb, abb_vb, nN, P_zb, &P_vb, &slope); (*nfall)++; }
    SigErr = 0; ora_tspsi(nN, abb_zc, abb_vc, YP, SIGM, &IER, &SigErr);
    if (SigErr == 0) P_vc = ora_hval(P_zc, nN, abb_zc, abb_vc, YP, SIGM, &IER);
    else { ora_linint(abb_zc, abb_vc, nN, P_zc, &P_vc, &slope); (*nfall)++; }
    SigErr = 0; ora_tspsi(nN, abb_zf, abb_vf, YP, SIGM, &IER, &SigErr);
    if (SigErr == 0) P_vf = ora_hval(P_zf, nN, abb_zf, abb_vf, YP, SIGM, &IER);
    else { ora_linint(abb_zf, abb_vf, nN, P_zf, &P_vf, &slope); if (p != 1) (*nfall)++; }   /* unused when p == 1 */
    double ey[3];
    if (p == 1) { ey[0] = P_vb; ey[1] = P_vb; ey[2] = P_vc; }      /* ledger 8 */
    else        { ey[0] = P_vb; ey[1] = P_vc; ey[2] = P_vf; }
    double vb = ora_polintd(ex, ey, ix[0]);
    double vc = ora_polintd(ex, ey, ix[1]);
    double vf = ora_polintd(ex, ey, ix[2]);
    double P_V = (vb + vc * 4 + vf) / 6.0;
    switch (v) { case 1: return vb; case 2: return vc; case 3: return vf; default: return P_V; }
}

/* WCTS_ITPI on bare 4-level profiles (tests/test_oracle_differential.py) */
double ora_wcts_profile(const double* zb, const double* zc, const double* zf, const double* vb, const double* vc, const double* vf,
    double P_zb, double P_zc, double P_zf, const double ex[3], const double ix[3], int32_t p, int32_t v, int32_t* nfall)
{
    int nf = 0;
    double r = WCTS_core(zb, zc, zf, vb, vc, vf, P_zb, P_zc, P_zf, ex, ix, p, v, &nf);
    *nfall = nf;
    return r;
}

/* ---- find_currents (LTRANS.f90:1422-1614) --------------------------------- */
static void find_currents_core(int us, int ws, double z0, valfn get, const void* h, double Zpar,
    const double* Pwc_zb, const double* Pwc_zc, const double* Pwc_zf,
    const double* Pwc_wzb, const double* Pwc_wzc, const double* Pwc_wzf,
    double P_zb, double P_zc, double P_zf, const double ex[3], const double ix[3], int p, int version,
    double* Uad, double* Vad, double* Wad)
{
    int i;
    for (i = 3; i <= us - 2; ++i)
        if (Zpar < Pwc_zb[i - 1] || Zpar < Pwc_zc[i - 1] || Zpar < Pwc_zf[i - 1]) break;
    int ii = i - 2;
    for (i = 3; i <= ws - 2; ++i)
        if (Zpar < Pwc_wzb[i - 1] || Zpar < Pwc_wzc[i - 1] || Zpar < Pwc_wzf[i - 1]) break;
    int iii = i - 2;
    if (Zpar < Pwc_wzb[0] + z0 || Zpar < Pwc_wzc[0] + z0 || Zpar < Pwc_wzf[0] + z0) {
        *Uad = 0.0; *Vad = 0.0; *Wad = 0.0;
    } else if (Zpar < Pwc_zb[0] || Zpar < Pwc_zc[0] || Zpar < Pwc_zf[0]) {
        double Ub = get(h, FLD_U, 0, 1), Uc = get(h, FLD_U, 1, 1), Uf = get(h, FLD_U, 2, 1);
        double Vb = get(h, FLD_V, 0, 1), Vc = get(h, FLD_V, 1, 1), Vf = get(h, FLD_V, 2, 1);
        double Wb = get(h, FLD_W, 0, 2), Wc = get(h, FLD_W, 1, 2), Wf = get(h, FLD_W, 2, 2);
        /* :1512 `stop 'dividing by 0'` when z0 == 0: caller guarantees z0 != 0 (checked at create) */
        double P_Ub = Ub * log10((Zpar - Pwc_wzb[0]) / z0) / log10((Pwc_zb[0] - Pwc_wzb[0]) / z0);
        double P_Uc = Uc * log10((Zpar - Pwc_wzb[0]) / z0) / log10((Pwc_zc[0] - Pwc_wzb[0]) / z0);
        double P_Uf = Uf * log10((Zpar - Pwc_wzb[0]) / z0) / log10((Pwc_zf[0] - Pwc_wzb[0]) / z0);
        double P_Vb = Vb * log10((Zpar - Pwc_wzb[0]) / z0) / log10((Pwc_zb[0] - Pwc_wzb[0]) / z0);
        double P_Vc = Vc * log10((Zpar - Pwc_wzb[0]) / z0) / log10((Pwc_zc[0] - Pwc_wzb[0]) / z0);
        double P_Vf = Vf * log10((Zpar - Pwc_wzb[0]) / z0) / log10((Pwc_zf[0] - Pwc_wzb[0]) / z0);
        double P_Wb = Wb * log10((Zpar - Pwc_wzb[0]) / z0) / log10((Pwc_wzb[1] - Pwc_wzb[0]) / z0);
        double P_Wc = Wc * log10((Zpar - Pwc_wzb[0]) / z0) / log10((Pwc_wzc[1] - Pwc_wzb[0]) / z0);
        double P_Wf = Wf * log10((Zpar - Pwc_wzb[0]) / z0) / log10((Pwc_wzf[1] - Pwc_wzb[0]) / z0);
        double ey[3], xt = ix[version - 1];
        if (p == 1) { ey[0] = P_Ub; ey[1] = P_Ub; ey[2] = P_Uc; } else { ey[0] = P_Ub; ey[1] = P_Uc; ey[2] = P_Uf; }
        *Uad = ora_polintd(ex, ey, xt);
        if (p == 1) { ey[0] = P_Vb; ey[1] = P_Vb; ey[2] = P_Vc; } else { ey[0] = P_Vb; ey[1] = P_Vc; ey[2] = P_Vf; }
        *Vad = ora_polintd(ex, ey, xt);
        if (p == 1) { ey[0] = P_Wb; ey[1] = P_Wb; ey[2] = P_Wc; } else { ey[0] = P_Wb; ey[1] = P_Wc; ey[2] = P_Wf; }
        *Wad = ora_polintd(ex, ey, xt);
    } else {
        *Uad = WCTS_get(get, h, FLD_U, ii, Pwc_zb, Pwc_zc, Pwc_zf, P_zb, P_zc, P_zf, ex, ix, p, version);
        *Vad = WCTS_get(get, h, FLD_V, ii, Pwc_zb, Pwc_zc, Pwc_zf, P_zb, P_zc, P_zf, ex, ix, p, version);
        *Wad = WCTS_get(get, h, FLD_W, iii, Pwc_wzb, Pwc_wzc, Pwc_wzf, P_zb, P_zc, P_zf, ex, ix, p, version);
    }
}
static void find_currents(ora_ctx* c, const elestate* es, double Xpar, double Ypar, double Zpar,
    const double* Pwc_zb, const double* Pwc_zc, const double* Pwc_zf,
    const double* Pwc_wzb, const double* Pwc_wzc, const double* Pwc_wzf,
    double P_zb, double P_zc, double P_zf, const double ex[3], const double ix[3], int p, int version,
    double* Uad, double* Vad, double* Wad)
{
    stepvals sv = { c, es, Xpar, Ypar };
    find_currents_core(c->prm.us, c->prm.ws, c->prm.z0, stepvals_get, &sv, Zpar, Pwc_zb, Pwc_zc, Pwc_zf, Pwc_wzb, Pwc_wzc, Pwc_wzf,
                       P_zb, P_zc, P_zf, ex, ix, p, version, Uad, Vad, Wad);
}

/* find_currents (LTRANS.f90:1422-1614) on a bare water column: rho- and w-level depths z[3][us], wz[3][ws] and the
 * u, v (us levels) and w (ws levels) profiles at the particle, each [3 records][levels]; *nfall = SigErr fall-backs */
typedef struct { const double *u, *v, *w; int us, ws; } colvals;
static double colvals_get(const void* h, int fld, int when, int lev)
{
    const colvals* q = (const colvals*)h;
    if (fld == FLD_U) return q->u[when * q->us + lev - 1];
    if (fld == FLD_V) return q->v[when * q->us + lev - 1];
    return q->w[when * q->ws + lev - 1];
}
void ora_find_currents_column(int32_t us, int32_t ws, double z0, double Zpar, const double* z, const double* wz,
    const double* u, const double* v, const double* w, double P_zb, double P_zc, double P_zf,
    const double ex[3], const double ix[3], int32_t p, int32_t version, double out[3], int32_t* nfall)
{
    colvals q = { u, v, w, us, ws };
    int before = tl_nsig;
    find_currents_core(us, ws, z0, colvals_get, &q, Zpar, z, z + us, z + 2 * us, wz, wz + ws, wz + 2 * ws,
                       P_zb, P_zc, P_zf, ex, ix, p, version, &out[0], &out[1], &out[2]);
    *nfall = tl_nsig - before; tl_nsig = before;
}

/* ---- HTurb (hor_turb_module.f90:29-50) ------------------------------------- */
static void HTurb(ora_ctx* c, prng* g, double* TurbHx, double* TurbHy)
{
    double r = 1.0, KM = c->prm.ConstantHTurb;
    double devX = norm_(g, 0u, 0), devY = norm_(g, 0u, 2);
    *TurbHx = devX * pow(2.0 / r * KM * c->prm.idt, 0.5);     /* ledger 12: (...)**0.5 */
    *TurbHy = devY * pow(2.0 / r * KM * c->prm.idt, 0.5);
}

/* ---- VTurb (ver_turb_module.f90:30-380) ------------------------------------ */
/* Everything of VTurb after step i. (the KH gather) and with the `loop` normal deviates of step ix.c handed in, so that
 * it can be driven on a bare water column (ora_vturb_column, tests/test_oracle_differential.py).  trace_id: ORA_TRACE_ID. */
static double VTurb_core(int ws, int idt, int p, const double ex[3], const double ix[3],
    const double* KHb, const double* KHc, const double* KHf,
    const double* Pwc_wzb, const double* Pwc_wzc, const double* Pwc_wzf,
    double P_zc, double P_depth, double P_zetac, const double* dev, int* sigerr_out, long long trace_id)
{
    const double background = F32(1.0E-6);                  /* ledger 2 */
    int p2 = ws * 4;
    size_t nb = sizeof(double) * (size_t)(p2 + 8);
    double *slb = malloc(nb), *slc = malloc(nb), *slf = malloc(nb), *icb = malloc(nb), *icc = malloc(nb), *icf = malloc(nb);
    double *mxb = malloc(nb), *myb = malloc(nb), *mxc = malloc(nb), *myc = malloc(nb), *mxf = malloc(nb), *myf = malloc(nb);
    double *fxb = malloc(nb), *fyb = malloc(nb), *fxc = malloc(nb), *fyc = malloc(nb), *fxf = malloc(nb), *fyf = malloc(nb);
    double *fx = malloc(nb), *fy = malloc(nb);
    double *nxb = calloc(p2 + 8, 8), *nyb = calloc(p2 + 8, 8), *nxc = calloc(p2 + 8, 8), *nyc = calloc(p2 + 8, 8),
           *nxf = calloc(p2 + 8, 8), *nyf = calloc(p2 + 8, 8);
    double *YPK = malloc(nb), *SIGK = malloc(nb);
    /* ii.a :120-124 (arrays are 1-based in the comments, 0-based here) */
    for (int j = 1; j <= p2 + 7; ++j) {
        nxb[j - 1] = Pwc_wzb[0] + ((double)(float)(j - 4)) * (Pwc_wzb[ws - 1] - Pwc_wzb[0]) / (double)p2;
        nxc[j - 1] = Pwc_wzc[0] + ((double)(float)(j - 4)) * (Pwc_wzc[ws - 1] - Pwc_wzc[0]) / (double)p2;
        nxf[j - 1] = Pwc_wzf[0] + ((double)(float)(j - 4)) * (Pwc_wzf[ws - 1] - Pwc_wzf[0]) / (double)p2;
    }
    for (int i = 1; i <= ws - 1; ++i) {                      /* :126-133 */
        slb[i - 1] = (KHb[i - 1] - KHb[i]) / (Pwc_wzb[i - 1] - Pwc_wzb[i]);
        icb[i - 1] = KHb[i - 1] - slb[i - 1] * Pwc_wzb[i - 1];
        slc[i - 1] = (KHc[i - 1] - KHc[i]) / (Pwc_wzc[i - 1] - Pwc_wzc[i]);
        icc[i - 1] = KHc[i - 1] - slc[i - 1] * Pwc_wzc[i - 1];
        slf[i - 1] = (KHf[i - 1] - KHf[i]) / (Pwc_wzf[i - 1] - Pwc_wzf[i]);
        icf[i - 1] = KHf[i - 1] - slf[i - 1] * Pwc_wzf[i - 1];
    }
    int jlo = 1;                                             /* :135-166 */
    for (int j = 5; j <= p2 + 3; ++j) { while (!(Pwc_wzb[jlo] > nxb[j - 1])) jlo++; nyb[j - 1] = slb[jlo - 1] * nxb[j - 1] + icb[jlo - 1]; }
    jlo = 1;
    for (int j = 5; j <= p2 + 3; ++j) { while (!(Pwc_wzc[jlo] > nxc[j - 1])) jlo++; nyc[j - 1] = slc[jlo - 1] * nxc[j - 1] + icc[jlo - 1]; }
    jlo = 1;
    for (int j = 5; j <= p2 + 3; ++j) { while (!(Pwc_wzf[jlo] > nxf[j - 1])) jlo++; nyf[j - 1] = slf[jlo - 1] * nxf[j - 1] + icf[jlo - 1]; }
    for (int i = 1; i <= 4; ++i) {                           /* :169-177 (ledger 11: KHb(1) for all) */
        nyb[i - 1] = KHb[0]; nyc[i - 1] = KHb[0]; nyf[i - 1] = KHb[0];
        nyb[i + p2 + 2] = KHb[ws - 1]; nyc[i + p2 + 2] = KHc[ws - 1]; nyf[i + p2 + 2] = KHf[ws - 1];
    }
    for (int i = 2; i <= p2 - 1; ++i) {                      /* :184-194 */
        myb[i - 1] = (nyb[i - 1] + nyb[i] + nyb[i + 1] + nyb[i + 2] + nyb[i + 3] + nyb[i + 4] + nyb[i + 5] + nyb[i + 6]) / 8.0;
        mxb[i - 1] = nxb[i - 1] + (nxb[i + 6] - nxb[i - 1]) / 2.0;
        myc[i - 1] = (nyc[i - 1] + nyc[i] + nyc[i + 1] + nyc[i + 2] + nyc[i + 3] + nyc[i + 4] + nyc[i + 5] + nyc[i + 6]) / 8.0;
        mxc[i - 1] = nxc[i - 1] + (nxc[i + 6] - nxc[i - 1]) / 2.0;
        myf[i - 1] = (nyf[i - 1] + nyf[i] + nyf[i + 1] + nyf[i + 2] + nyf[i + 3] + nyf[i + 4] + nyf[i + 5] + nyf[i + 6]) / 8.0;
        mxf[i - 1] = nxf[i - 1] + (nxf[i + 6] - nxf[i - 1]) / 2.0;
    }
    mxb[0] = Pwc_wzb[0]; myb[0] = KHb[0]; mxb[p2 - 1] = Pwc_wzb[ws - 1]; myb[p2 - 1] = KHb[ws - 1];   /* :197-210 */
    mxc[0] = Pwc_wzc[0]; myc[0] = KHc[0]; mxc[p2 - 1] = Pwc_wzc[ws - 1]; myc[p2 - 1] = KHc[ws - 1];
    mxf[0] = Pwc_wzf[0]; myf[0] = KHf[0]; mxf[p2 - 1] = Pwc_wzf[ws - 1]; myf[p2 - 1] = KHf[ws - 1];
    for (int k = 0; k < p2; ++k) {                           /* :220-262 */
        double ey[3];
        if (p == 1) { ey[0] = mxb[k]; ey[1] = mxb[k]; ey[2] = mxc[k]; } else { ey[0] = mxb[k]; ey[1] = mxc[k]; ey[2] = mxf[k]; }
        fxb[k] = ora_polintd(ex, ey, ix[0]); fxc[k] = ora_polintd(ex, ey, ix[1]); fxf[k] = ora_polintd(ex, ey, ix[2]);
        if (p == 1) { ey[0] = myb[k]; ey[1] = myb[k]; ey[2] = myc[k]; } else { ey[0] = myb[k]; ey[1] = myc[k]; ey[2] = myf[k]; }
        fyb[k] = ora_polintd(ex, ey, ix[0]); fyc[k] = ora_polintd(ex, ey, ix[1]); fyf[k] = ora_polintd(ex, ey, ix[2]);
    }
    for (int k = 0; k < p2; ++k) {                           /* :264-275 */
        if (fyb[k] < 0.0) fyb[k] = 0.0;
        if (fyc[k] < 0.0) fyc[k] = 0.0;
        if (fyf[k] < 0.0) fyf[k] = 0.0;
        fy[k] = (fyb[k] + 4.0 * fyc[k] + fyf[k]) / 6.0;
        fx[k] = (fxb[k] + 4.0 * fxc[k] + fxf[k]) / 6.0;
    }
    int IER, SigErr = 0;
    ora_tspsi(p2, fx, fy, YPK, SIGK, &IER, &SigErr);         /* :278-279 */
    *sigerr_out = SigErr;
    double deltat = 2.0;
    int loop = idt / (int)deltat;                            /* :283 */
    double ParZc = P_zc;
    for (int i = 1; i <= loop; ++i) {                        /* :291-337 */
        double Kprimec = 0.0, thisyc, slopem;
        if (ParZc < P_depth || ParZc > P_zetac) Kprimec = 0.0;
        else if (SigErr == 0) Kprimec = ora_hpval(ParZc, p2, fx, fy, YPK, SIGK, &IER);
        else ora_linint(fx, fy, p2, ParZc, &thisyc, &Kprimec);
        double KprimeZc = -1.0 * Kprimec * deltat;
        double Z3rdc = ParZc + 0.5 * KprimeZc;
        double KH3rdc = 0.0;
        if (Z3rdc < P_depth || Z3rdc > P_zetac) KH3rdc = background;
        else {
            if (SigErr == 0) KH3rdc = ora_hval(Z3rdc, p2, fx, fy, YPK, SIGK, &IER);
            else ora_linint(fx, fy, p2, Z3rdc, &KH3rdc, &slopem);
            if (KH3rdc < background) KH3rdc = background;
        }
        double DEV = dev[i - 1];
        double r = 1.;
        ParZc = ParZc + KprimeZc + DEV * pow(2.0 / r * KH3rdc * deltat, 0.5);
        if (trace_id >= 0 && getenv("ORA_TRACE_ID") && atoll(getenv("ORA_TRACE_ID")) == trace_id)
            fprintf(stderr, "ORATRACE %d %.17g %.17g %.17g %.17g\n", i - 1, ParZc, Kprimec, KH3rdc, DEV);
    }
    double TurbV = P_zc - ParZc;                             /* :342 (ledger 11) */
    free(slb); free(slc); free(slf); free(icb); free(icc); free(icf);
    free(mxb); free(myb); free(mxc); free(myc); free(mxf); free(myf);
    free(fxb); free(fyb); free(fxc); free(fyc); free(fxf); free(fyf); free(fx); free(fy);
    free(nxb); free(nyb); free(nxc); free(nyc); free(nxf); free(nyf); free(YPK); free(SIGK);
    return TurbV;
}

static double VTurb(ora_ctx* c, const elestate* es, prng* g, double P_zc, double P_depth, double P_zetac,
    int p, const double ex[3], const double ix[3],
    const double* Pwc_wzb, const double* Pwc_wzc, const double* Pwc_wzf)
{
    int ws = c->prm.ws, loop = c->prm.idt / 2;
    double *KHb = malloc(sizeof(double) * (size_t)ws), *KHc = malloc(sizeof(double) * (size_t)ws), *KHf = malloc(sizeof(double) * (size_t)ws);
    double* dev = malloc(sizeof(double) * (size_t)(loop > 0 ? loop : 1));
    for (int i = 1; i <= ws; ++i) {                          /* i. :102-108 */
        KHb[i - 1] = getInterp_kh(c, es, c->t_b, i);
        KHc[i - 1] = getInterp_kh(c, es, c->t_c, i);
        KHf[i - 1] = getInterp_kh(c, es, c->t_f, i);
    }
    /* the deviates of step ix.c in draw order (nothing else draws inside the loop) */
    for (int q = 0; q < loop; ++q) dev[q] = norm_(g, 1u + (uint32_t)(q >> 1), 2 * (q & 1));
    int SigErr = 0;
    double TurbV = VTurb_core(ws, c->prm.idt, p, ex, ix, KHb, KHc, KHf, Pwc_wzb, Pwc_wzc, Pwc_wzf, P_zc, P_depth, P_zetac, dev, &SigErr,
                              (long long)(((uint64_t)g->id_hi << 32) | g->id_lo));
    if (SigErr != 0) tl_nsig++;
    free(KHb); free(KHc); free(KHf); free(dev);
    return TurbV;
}

/* VTurb on a bare column: KH and w-level depths at the three hydro times, `idt / 2` normal deviates. */
double ora_vturb_column(int32_t ws, int32_t idt, int32_t p, const double ex[3], const double ix[3],
    const double* KHb, const double* KHc, const double* KHf, const double* wzb, const double* wzc, const double* wzf,
    double P_zc, double P_depth, double P_zetac, const double* dev, int32_t* sigerr)
{
    int se = 0;
    double t = VTurb_core(ws, idt, p, ex, ix, KHb, KHc, KHf, wzb, wzc, wzf, P_zc, P_depth, P_zetac, dev, &se, -1);
    *sigerr = se;
    return t;
}

/* ---- behave (behavior_module.f90:181-551) ---------------------------------- */
static void behave(ora_ctx* c, const elestate* es, prng* g, double Xpar, double Ypar, double Zpar,
    const double* Pwc_zb, const double* Pwc_zc, const double* Pwc_zf, double P_zb, double P_zc, double P_zf,
    double P_zetac, double P_age, double P_depth, double P_U, double P_V, double P_angle, int n, int it,
    const double ex[3], const double ix[3], double daytime, int p,
    int* bott, double* XBehav, double* YBehav, double* ZBehav)
{
    const ltgpu_params* P = &c->prm;
    int us = P->us, w = 0;                       /* w = next behave random word */
    const uint32_t BB = 0x80000000u;
    double negpos, dev1, devB, sw, switchslope, P_S = 0.0, parBehav, Sslope, deltaS, deltaz;
    int btest;
    *XBehav = 0.0; *YBehav = 0.0; *ZBehav = 0.0;
    /* per-particle constants of initBehave (:118-131): uniform in v.2b */
    double pediage_n = P->pediage, deadage_n = P->deadage;
    double swim1 = (P->swimfast - P->swimslow) / (pediage_n - P->swimstart);
    double swim2 = P->swimfast - swim1 * pediage_n;
    if (P_age >= P->swimstart) c->P_swim3[n] = swim1 * P_age + swim2;      /* :214-215 */
    if (P_age >= pediage_n) c->P_swim3[n] = P->swimfast;
    double swim3 = c->P_swim3[n];
    if (c->P_behave[n] == 4 || c->P_behave[n] == 5) {                      /* :220-229 */
        if (P_age >= pediage_n && P_age < deadage_n) c->P_behave[n] = 2;
        c->timer[n] = fmax(0.0, c->timer[n] - (double)P->dt);              /* ledger 15 */
    }
    if (c->P_behave[n] == 4 || (c->P_behave[n] == 5 && c->timer[n] == 0.0) || c->P_behave[n] == 7) {
        if (c->script_PS) P_S = *c->script_PS;
        else {
            int i;
            for (i = 3; i <= us - 2; ++i)
                if (Zpar < Pwc_zb[i - 1] || Zpar < Pwc_zc[i - 1] || Zpar < Pwc_zf[i - 1]) break;
            int deplvl = i - 2;
            P_S = WCTS_ITPI(c, es, FLD_SALT, Xpar, Ypar, deplvl, Pwc_zb, Pwc_zc, Pwc_zf, P_zb, P_zc, P_zf, ex, ix, p, 4);
        }
    }
    parBehav = 0.0;
    if (c->P_behave[n] == 1) {                                             /* :258-282 */
        btest = 0;
        if (P_zc < (P_zetac - 1.0)) {
            negpos = 1.0; dev1 = real1(g, BB, w++); sw = F32(0.80);
            if (dev1 > sw) negpos = -1.0;
            devB = real1(g, BB, w++); parBehav = negpos * devB * swim3; btest = 1;
        }
        if (btest == 0) {
            negpos = 1.0; dev1 = real1(g, BB, w++); sw = 0.5;
            if (dev1 > sw) negpos = -1.0;
            devB = real1(g, BB, w++); parBehav = negpos * devB * swim3;
        }
    }
    if (c->P_behave[n] == 2 || (c->P_behave[n] == 5 && c->timer[n] > 0.0)) {   /* :286-311 */
        btest = 0;
        if (P_zc > (P_depth + 1.0)) {
            negpos = 1.0; dev1 = real1(g, BB, w++); sw = F32(0.20);
            if (dev1 > sw) negpos = -1.0;
            devB = real1(g, BB, w++); parBehav = negpos * devB * swim3; btest = 1;
        }
        if (btest == 0) {
            negpos = 1.0; dev1 = real1(g, BB, w++); sw = 0.5;
            if (dev1 > sw) negpos = -1.0;
            devB = real1(g, BB, w++); parBehav = negpos * devB * swim3;
        }
    }
    if (c->P_behave[n] == 3) {                                             /* :314-353 */
        double dtime = (daytime - trunc(daytime)) * 24.0;
        double tst, E0;
        if (dtime > P->twistart && dtime < P->twiend) {
            tst = (dtime - P->twistart) * 3600.0;
            E0 = P->Em * sin(P->PI * tst / (P->daylength * 3600.0)) * sin(P->PI * tst / (P->daylength * 3600.0));
        } else E0 = 0.0;
        double P_light = E0 * exp(P->Kd * P_zc);
        if (P_light < P->thresh) {
            negpos = 1.0; dev1 = real1(g, BB, w++); sw = 0.5;
            if (dev1 > sw) negpos = -1.0;
            devB = real1(g, BB, w++); parBehav = negpos * devB * swim3;
        }
        if (P_light > P->thresh) {
            negpos = 1.0; dev1 = real1(g, BB, w++); sw = F32(0.20);
            if (dev1 > sw) negpos = -1.0;
            devB = real1(g, BB, w++); parBehav = negpos * devB * swim3;
        }
    }
    if (c->P_behave[n] == 4) {                                             /* :357-406 */
        if (it == 1) { c->P_Sprev[n] = P_S; c->P_zprev[n] = P_zc; }
        btest = 0; Sslope = 0.0;
        deltaS = c->P_Sprev[n] - P_S; deltaz = c->P_zprev[n] - P_zc;
        if (it > 1) Sslope = deltaS / deltaz;
        if (fabs(Sslope) > P->Sgradient) {
            negpos = 1.0; dev1 = real1(g, BB, w++); sw = F32(0.80);
            if (dev1 > sw) negpos = -1.0;
            parBehav = negpos * swim3; btest = 1;
        }
        if (btest == 0) {
            negpos = 1.0; dev1 = real1(g, BB, w++);
            if (P_age < 1.5 * 24. * 3600.) sw = F32(0.1);
            else if (P_age < 5. * 24. * 3600.) sw = F32(0.49);
            else if (P_age < 8. * 24. * 3600.) sw = F32(0.50);
            else {
                switchslope = (F32(0.50) - F32(0.517)) / (8.0 * 24.0 * 3600.0 - pediage_n);
                sw = switchslope * P_age + F32(0.50) - switchslope * 8.0 * 24.0 * 3600.0;
                if (P_zc < P_depth + 1.) sw = 0.5;
            }
            if (dev1 > (1 - sw)) negpos = -1.0;
            devB = real1(g, BB, w++); parBehav = negpos * devB * swim3;
        }
        c->P_Sprev[n] = P_S; c->P_zprev[n] = P_zc;
    }
    if (c->P_behave[n] == 5 && c->timer[n] == 0.0) {                       /* :410-463 */
        if (it == 1) { c->P_Sprev[n] = P_S; c->P_zprev[n] = P_zc; }
        btest = 0; Sslope = 0.0;
        deltaS = c->P_Sprev[n] - P_S; deltaz = c->P_zprev[n] - P_zc;
        if (it > 1) Sslope = deltaS / deltaz;
        if (fabs(Sslope) > P->Sgradient) {
            negpos = 1.0; dev1 = real1(g, BB, w++); sw = F32(0.20); btest = 1;
            c->timer[n] = 2.0 * 3600.0;
            if (dev1 > sw) negpos = -1.0;
            parBehav = negpos * swim3;
            if (P_age < 3.5 * 24. * 3600.) { btest = 0; c->timer[n] = 0.; }
        }
        if (btest == 0) {
            negpos = 1.0; dev1 = real1(g, BB, w++); sw = F32(0.495);
            if (P_age < 1.5 * 24. * 3600.) sw = F32(0.9);
            if (P_age > 2.0 * 24. * 3600. && P_age < 3.5 * 24. * 3600.) {
                switchslope = (F32(0.3) - F32(0.495)) / (2.0 * 24.0 * 3600.0 - 3.5 * 24.0 * 3600.0);
                sw = switchslope * P_age + F32(0.3) - switchslope * 2.0 * 24.0 * 3600.0;
            }
            if (dev1 > sw) negpos = -1.0;
            devB = real1(g, BB, w++); parBehav = negpos * devB * swim3;
        }
        c->P_Sprev[n] = P_S; c->P_zprev[n] = P_zc;
    }
    if (c->P_behave[n] == 6) {                                             /* :466-471 */
        if (P_age >= P->swimstart) parBehav = P->sink; else parBehav = swim3;
    }
    *ZBehav = parBehav * P->idt;                                           /* :491 */
    if (c->P_behave[n] == 7) {                                             /* :495-548 */
        if (it == 1) c->P_Sprev[n] = P_S;
        double ca = cos(P_angle), sa = sin(P_angle);
        double currentspeed = sqrt((P_U * ca - P_V * sa) * (P_U * ca - P_V * sa) + (P_U * sa + P_V * ca) * (P_U * sa + P_V * ca));
        if (c->bottom[n]) {
            if (c->P_Sprev[n] < P_S) { c->bottom[n] = 0; *ZBehav = P_depth + P->Swimdepth; }
            else *ZBehav = -9999;
        } else {
            if (currentspeed > F32(0.05)) {
                double Hdistance = P->Hswimspeed * P->idt;
                double theta = atan((P_U * sa + P_V * ca) / (P_U * ca - P_V * sa));
                double X = (P_U * ca - P_V * sa), Y = (P_U * sa + P_V * ca);
                if (X > 0.0) { *XBehav = Hdistance * cos(theta); *YBehav = Hdistance * sin(theta); }
                if (X < 0.0) { *XBehav = -1.0 * Hdistance * cos(theta); *YBehav = -1.0 * Hdistance * sin(theta); }
                if (X == 0 && Y >= 0.0) { *XBehav = 0.0; *YBehav = Hdistance; }
                if (X == 0 && Y <= 0.0) { *XBehav = 0.0; *YBehav = -1.0 * Hdistance; }
                *ZBehav = P_depth + P->Swimdepth;
            } else { *ZBehav = -9999; c->bottom[n] = 1; }
        }
        *bott = c->bottom[n];
    }
}

/* behave (behavior_module.f90:181-551) for ONE particle with everything it reads handed in (tests/test_oracle_differential.py):
 * state = {P_behave, P_swim(n,3), timer, P_Sprev, P_zprev, bottom} in and out, P_S the salinity WCTS_ITPI would return,
 * words = the genrand_int32 values behind the genrand_real1 calls in draw order; out = XBehav, YBehav, ZBehav, bott,
 * words used.  Runs the step's own routine on a one-particle context. */
void ora_behave_case(const ltgpu_params* prm, double state[6], double P_S, const uint32_t* words,
    double Zpar, double P_zc, double P_zetac, double P_age, double P_depth, double P_U, double P_V, double P_angle,
    int32_t it, double daytime, double out[5])
{
    ora_ctx c; memset(&c, 0, sizeof c);
    c.prm = *prm; c.rng_mode = ORA_RNG_PHILOX; c.script = words; c.script_pos = 0; c.script_PS = &P_S;
    int32_t beh = (int32_t)state[0]; double swim3 = state[1], timer = state[2], sprev = state[3], zprev = state[4];
    uint8_t bottom = state[5] != 0.0;
    c.P_behave = &beh; c.P_swim3 = &swim3; c.timer = &timer; c.P_Sprev = &sprev; c.P_zprev = &zprev; c.bottom = &bottom;
    prng g; memset(&g, 0, sizeof g); g.c = &c;
    double ex[3] = {0, 0, 0}, ix[3] = {0, 0, 0};
    int bott = 0;
    behave(&c, NULL, &g, 0.0, 0.0, Zpar, NULL, NULL, NULL, P_zc, P_zc, P_zc, P_zetac, P_age, P_depth, P_U, P_V, P_angle, 0, it,
           ex, ix, daytime, 2, &bott, &out[0], &out[1], &out[2]);
    out[3] = (double)bott; out[4] = (double)c.script_pos;
    state[0] = (double)beh; state[1] = swim3; state[2] = timer; state[3] = sprev; state[4] = zprev; state[5] = (double)bottom;
}

/* ---- boundary_module.f90:1515-1614 mbounds / ibounds ------------------------ */
int32_t ora_mbounds(ora_ctx* c, double Ypos, double Xpos)
{
    return ora_inpoly(Xpos, Ypos, c->maxbound, c->bx, c->by, -1) ? 1 : 0;
}
int32_t ora_ibounds(ora_ctx* c, double claty, double clongx, double* island)
{
    int in_island = 0; *island = 0.0;
    if (c->maxisland <= 0) return 0;                       /* numislands > 0 (:1565) */
    int i = 1, start = 0; int isle = c->hid[0];
    for (;;) {
        i = i + 1;
        int endIsle = 0;
        if (i == c->maxisland) endIsle = 1;
        else if (c->hid[i] != isle) endIsle = 1;           /* hid(i+1) */
        if (endIsle) {
            int count = i - start;
            if (ora_inpoly(clongx, claty, count, c->hx + start, c->hy + start, -1)) {
                in_island = 1; *island = isle; break;
            }
            if (i == c->maxisland) break;
            start = i; isle = c->hid[i];
        }
    }
    return in_island;
}

/* ---- boundary_module.f90:1620-1902 intersect_reflect ------------------------ */
int32_t ora_intersect_reflect(ora_ctx* c, double Xpos, double Ypos, double nXpos, double nYpos,
    double* fintersectX, double* fintersectY, double* freflectX, double* freflectY,
    int32_t* skipbound, int32_t* isWater)
{
    int intersect = 0, intersectf = 0, skipboundi = *skipbound;
    double Mbc = 0.0, Bbc = 0.0, Mp = 0.0, Bp = 0.0, distBC, crossk, dPBC, mBCperp, bBCperp;
    double rx1, rx2, ry1, ry2, dist1, dist2, intersctx = 0.0, interscty = 0.0, rPxyzX = 0.0, rPxyzY = 0.0;
    double xhigh, xlow, yhigh, ylow, bxhigh, bxlow, byhigh, bylow, dtest = 999999.;
    *fintersectX = -999999.; *fintersectY = -999999.; *freflectX = -999999.; *freflectY = -999999.;
    *isWater = 0;
    if (Xpos >= nXpos) { xhigh = Xpos; xlow = nXpos; } else { xhigh = nXpos; xlow = Xpos; }
    if (Ypos >= nYpos) { yhigh = Ypos; ylow = nYpos; } else { yhigh = nYpos; ylow = Ypos; }
#define INBOX() (intersctx <= xhigh && intersctx >= xlow && interscty <= yhigh && interscty >= ylow && \
                 intersctx <= bxhigh && intersctx >= bxlow && interscty <= byhigh && interscty >= bylow)
#define PICK() do { dist1 = sqrt((intersctx - rx1) * (intersctx - rx1) + (interscty - ry1) * (interscty - ry1)); \
                    dist2 = sqrt((intersctx - rx2) * (intersctx - rx2) + (interscty - ry2) * (interscty - ry2)); \
                    if (dist1 < dist2) { rPxyzX = rx1; rPxyzY = ry1; } \
                    else if (dist1 > dist2) { rPxyzX = rx2; rPxyzY = ry2; } \
                    intersect = 1; } while (0)
    for (int i = 1; i <= c->nbounds; ++i) {
        if (i == *skipbound) continue;
        intersect = 0;
        double bcx1 = c->bnd_x[2 * (i - 1)], bcy1 = c->bnd_y[2 * (i - 1)];
        double bcx2 = c->bnd_x[2 * (i - 1) + 1], bcy2 = c->bnd_y[2 * (i - 1) + 1];
        if ((bcx1 > xhigh && bcx2 > xhigh) || (bcx1 < xlow && bcx2 < xlow) ||
            (bcy1 > yhigh && bcy2 > yhigh) || (bcy1 < ylow && bcy2 < ylow)) continue;
        if (bcx1 >= bcx2) { bxhigh = bcx1; bxlow = bcx2; } else { bxhigh = bcx2; bxlow = bcx1; }
        if (bcy1 >= bcy2) { byhigh = bcy1; bylow = bcy2; } else { byhigh = bcy2; bylow = bcy1; }
        if (bcx1 == bcx2 || nXpos == Xpos) {
            if (bcx1 == bcx2 && nXpos == Xpos) continue;
            if (bcx1 == bcx2 && nYpos == Ypos) {                          /* :1703-1726 */
                intersctx = bcx1; interscty = nYpos;
                if (INBOX()) {
                    dPBC = sqrt((intersctx - nXpos) * (intersctx - nXpos) + (interscty - nYpos) * (interscty - nYpos));
                    rx1 = nXpos + (2.0 * dPBC); ry1 = nYpos; rx2 = nXpos - (2.0 * dPBC); ry2 = nYpos;
                    PICK();
                }
            } else if (nXpos == Xpos && bcy1 == bcy2) {                   /* :1727-1750 */
                intersctx = nXpos; interscty = bcy1;
                if (INBOX()) {
                    dPBC = sqrt((intersctx - nXpos) * (intersctx - nXpos) + (interscty - nYpos) * (interscty - nYpos));
                    rx1 = nXpos; ry1 = nYpos + (2.0 * dPBC); rx2 = nXpos; ry2 = nYpos - (2.0 * dPBC);
                    PICK();
                }
            } else if (bcx1 == bcx2 && nYpos != Ypos) {                   /* :1751-1776 */
                Mp = (nYpos - Ypos) / (nXpos - Xpos); Bp = Ypos - Mp * Xpos;
                intersctx = bcx1; interscty = Mp * intersctx + Bp;
                if (INBOX()) {
                    dPBC = nXpos - intersctx;
                    rx1 = nXpos + (2.0 * dPBC); ry1 = nYpos; rx2 = nXpos - (2.0 * dPBC); ry2 = nYpos;
                    PICK();
                }
            } else if (nXpos == Xpos && bcy1 != bcy2) {                   /* :1777-1813 */
                Mbc = (bcy2 - bcy1) / (bcx2 - bcx1); Bbc = bcy2 - Mbc * bcx2;
                intersctx = nXpos; interscty = Mbc * intersctx + Bbc;
                if (INBOX()) {
                    distBC = sqrt((bcx1 - bcx2) * (bcx1 - bcx2) + (bcy1 - bcy2) * (bcy1 - bcy2));
                    crossk = ((nXpos - bcx1) * (bcy2 - bcy1)) - ((bcx2 - bcx1) * (nYpos - bcy1));
                    dPBC = sqrt(crossk * crossk) / distBC;
                    mBCperp = -1.0 / Mbc; bBCperp = nYpos - mBCperp * nXpos;
                    rx1 = sqrt(((2.0 * dPBC) * (2.0 * dPBC)) / (1.0 + mBCperp * mBCperp)) + nXpos;
                    ry1 = mBCperp * rx1 + bBCperp;
                    rx2 = sqrt(((2.0 * dPBC) * (2.0 * dPBC)) / (1.0 + mBCperp * mBCperp)) * -1.0 + nXpos;
                    ry2 = mBCperp * rx2 + bBCperp;
                    PICK();
                }
            }
        } else {                                                          /* :1815-1883 */
            Mbc = (bcy2 - bcy1) / (bcx2 - bcx1); Bbc = bcy2 - Mbc * bcx2;
            Mp = (nYpos - Ypos) / (nXpos - Xpos); Bp = Ypos - Mp * Xpos;
            intersctx = (Bbc - Bp) / (Mp - Mbc);
            interscty = Mp * intersctx + Bp;
            if (Mbc == 0.0) interscty = byhigh;
            if (INBOX()) {
                if (Mbc == 0.0) {
                    dPBC = nYpos - bcy1;
                    rx1 = nXpos; ry1 = nYpos + (2.0 * dPBC); rx2 = nXpos; ry2 = nYpos - (2.0 * dPBC);
                    PICK();
                }
                if (intersect == 0) {
                    distBC = sqrt((bcx1 - bcx2) * (bcx1 - bcx2) + (bcy1 - bcy2) * (bcy1 - bcy2));
                    crossk = ((nXpos - bcx1) * (bcy2 - bcy1)) - ((bcx2 - bcx1) * (nYpos - bcy1));
                    dPBC = sqrt(crossk * crossk) / distBC;
                    mBCperp = -1.0 / Mbc; bBCperp = nYpos - mBCperp * nXpos;
                    rx1 = sqrt(((2.0 * dPBC) * (2.0 * dPBC)) / (1.0 + mBCperp * mBCperp)) + nXpos;
                    ry1 = mBCperp * rx1 + bBCperp;
                    rx2 = sqrt(((2.0 * dPBC) * (2.0 * dPBC)) / (1.0 + mBCperp * mBCperp)) * -1.0 + nXpos;
                    ry2 = mBCperp * rx2 + bBCperp;
                    PICK();
                }
            }
        }
        double d_Pinter = sqrt((Xpos - intersctx) * (Xpos - intersctx) + (Ypos - interscty) * (Ypos - interscty));
        if (intersect == 1 && d_Pinter < dtest) {                         /* :1887-1897 */
            *fintersectX = intersctx; *fintersectY = interscty;
            *freflectX = rPxyzX; *freflectY = rPxyzY;
            intersectf = 1; dtest = d_Pinter; skipboundi = i;
            *isWater = !c->land[i - 1];
        }
    }
#undef INBOX
#undef PICK
    *skipbound = skipboundi;
    return intersectf;
}

/* ---- settlement_module.f90:485-622 ------------------------------------------ */
static int find_poly_index(const int32_t* ids, int n, int id)
{
    for (int i = 0; i < n; ++i) if (ids[i] == id) return i;
    return -1;
}
static int testSettlement(ora_ctx* c, double P_age, int n, double Px, double Py)
{
    int R_ele = c->r_ele[n], polyin = 0, inpoly = 0;
    double settletime = c->prm.pediage;                   /* initSettlement(P_pediage) behavior:154 */
    if (!(P_age >= settletime)) return 0;
    int pidx = -1;
    /* psettle :524-571 */
    for (int q = c->elepoly_ptr[R_ele - 1]; q < c->elepoly_ptr[R_ele]; ++q) {
        int pi = c->elepoly_idx[q];
        int start = c->poly_start[pi], size = c->poly_size[pi];
        const double* col2 = c->polys + (size_t)c->pedges * 1, *col3 = c->polys + (size_t)c->pedges * 2;
        const double* col4 = c->polys + (size_t)c->pedges * 3, *col5 = c->polys + (size_t)c->pedges * 4;
        double dis = sqrt((Px - col2[start - 1]) * (Px - col2[start - 1]) + (Py - col3[start - 1]) * (Py - col3[start - 1]));
        if (dis > c->poly_maxdis[pi]) continue;
        if (ora_inpoly(Px, Py, size, col4 + (start - 1), col5 + (start - 1), -1)) {
            polyin = (int)lround(c->polys[start - 1]); pidx = pi; break;
        }
    }
    if (polyin > 0) {
        inpoly = polyin;
        if (c->prm.holesExist) {                          /* hsettle :574-622 */
            int holein = 0;
            if (pidx < 0) pidx = find_poly_index(c->poly_id, c->npoly, polyin);
            for (int q = c->polyhole_ptr[pidx]; q < c->polyhole_ptr[pidx + 1]; ++q) {
                int hi = c->polyhole_idx[q];
                int start = c->hole_start[hi], size = c->hole_size[hi];
                const double* col2 = c->holes + (size_t)c->hedges * 1, *col3 = c->holes + (size_t)c->hedges * 2;
                const double* col4 = c->holes + (size_t)c->hedges * 3, *col5 = c->holes + (size_t)c->hedges * 4;
                double dis = sqrt((Px - col2[start - 1]) * (Px - col2[start - 1]) + (Py - col3[start - 1]) * (Py - col3[start - 1]));
                if (dis > c->hole_maxdis[hi]) continue;
                if (ora_inpoly(Px, Py, size, col4 + (start - 1), col5 + (start - 1), 0)) {
                    holein = (int)lround(c->holes[start - 1]); break;
                }
            }
            if (holein != 0) inpoly = 0;
        }
    }
    if (inpoly > 0) c->settle[n] = 1;
    return inpoly;
}

/* testSettlement (settlement_module.f90:485-622) for a bare point in rho element R_ele of a context whose habitat
 * has been set (tests/test_oracle_differential.py): the polygon id the point settles in, 0 for none */
int32_t ora_settle_point(ora_ctx* c, int32_t R_ele, double P_age, double Px, double Py)
{
    int32_t re = R_ele; uint8_t st = 0;
    int32_t* keep_r = c->r_ele; uint8_t* keep_s = c->settle;
    c->r_ele = &re; c->settle = &st;
    int r = testSettlement(c, P_age, 0, Px, Py);
    c->r_ele = keep_r; c->settle = keep_s;
    return r;
}

/* ---- error handling shared by the four check sites of update_particles ------ */
/* returns 1 when the reference would STOP */
static int handle_error(ora_ctx* c, int n, int code, double t, double revertZ)
{
    int EF = c->prm.ErrorFlag;
    if (EF < 1 || EF > 3) {
#pragma omp critical(ora_bad)
        { int gid = (int)(c->first_id + n); if (c->bad_particle == 0 || gid < c->bad_particle) c->bad_particle = gid; }
        add_event(c, (int)(c->first_id + n), code, t);
        return 1;
    }
    if (EF == 1) { c->nX[n] = c->X[n]; c->nY[n] = c->Y[n]; c->nZ[n] = revertZ; }
    else if (EF == 2) c->dead[n] = 1;
    else c->oob[n] = 1;
    add_event(c, (int)(c->first_id + n), code, t);
    return 0;
}

/* ---- one particle, one internal step: LTRANS.f90:778-1400 ------------------- */
static int step_particle(ora_ctx* c, int n, int p, int it, const double ex[3], const double ix[3], uint32_t gstep)
{
    const ltgpu_params* P = &c->prm;
    int us = P->us, ws = P->ws, idt = P->idt;
    if (ix[2] <= c->DOB[n]) {                                            /* :790-795 */
        c->nX[n] = c->X[n]; c->nY[n] = c->Y[n]; c->nZ[n] = c->Z[n];
        return 0;
    }
    c->Age[n] = c->Age[n] + (double)(float)idt;                          /* :798 */
    /* updateStatus behavior:162-179 */
    if (c->Age[n] >= P->deadage && P->mortality) {
        if (P->settlementon) { if (!c->settle[n]) c->dead[n] = 1; } else c->dead[n] = 1;
    }
    if (P->settlementon && c->settle[n]) return 0;                       /* :804-816 */
    if (P->mortality && c->dead[n]) return 0;
    if (P->OpenOceanBoundary && c->oob[n]) return 0;

    int64_t gid = c->first_id + n;
    prng g = { c, (uint32_t)((uint64_t)gid & 0xffffffffu), (uint32_t)((uint64_t)gid >> 32), gstep };
    elestate es;
    double Xpar = c->X[n], Ypar = c->Y[n];
    int ele_err = setEle(c, Xpar, Ypar, n, &es);                         /* :830 */
    if (ele_err > 0) {
        int code = ele_err == 4 ? LTGPU_EV_NOT_IN_RHO : ele_err == 5 ? LTGPU_EV_NOT_IN_U : LTGPU_EV_NOT_IN_V;
        return handle_error(c, n, code, ix[2], c->Z[n]);
    }
    setInterp(c, &es, Xpar, Ypar);                                       /* :882 */
    double P_depth = -1.0 * getInterp_static(c, &es, c->depth);          /* :892-896 */
    double P_angle = getInterp_static(c, &es, c->angle);
    double P_zetab = getInterp_zeta(c, &es, c->t_b);
    double P_zetac = getInterp_zeta(c, &es, c->t_c);
    double P_zetaf = getInterp_zeta(c, &es, c->t_f);
    if (c->Z[n] < P_depth) {                                             /* :900-903 */
        c->Z[n] = P_depth + F32(0.001);
        if (P->TrackCollisions) c->hitBottom[n]++;
    }
    double P_zb = c->Z[n], P_zc = c->Z[n], P_zf = c->Z[n];
    if (c->Z[n] > P_zetab) P_zb = P_zetab - F32(0.001);
    if (c->Z[n] > P_zetac) P_zc = P_zetac - F32(0.001);
    if (c->Z[n] > P_zetaf) P_zf = P_zetaf - F32(0.001);
    double ey[3] = { P_zb, P_zc, P_zf };
    double Zpar = ora_polintd(ex, ey, ix[1]);                            /* :914 */
    ey[0] = P_zetab; ey[1] = P_zetac; ey[2] = P_zetaf;
    double P_zeta = ora_polintd(ex, ey, ix[1]); (void)P_zeta;            /* :919 (unused later) */
    c->Z[n] = Zpar;                                                      /* :922 */

    double Pwc_zb[us], Pwc_zc[us], Pwc_zf[us], Pwc_wzb[ws], Pwc_wzc[ws], Pwc_wzf[ws];
    int i;
    for (i = 1; i <= us; ++i) {                                          /* :934-946 */
        Pwc_zb[i - 1] = ora_slevel(P_zetab, P_depth, c->SC[i - 1], c->CS[i - 1], P->hc, P->Vtransform);
        Pwc_zc[i - 1] = ora_slevel(P_zetac, P_depth, c->SC[i - 1], c->CS[i - 1], P->hc, P->Vtransform);
        Pwc_zf[i - 1] = ora_slevel(P_zetaf, P_depth, c->SC[i - 1], c->CS[i - 1], P->hc, P->Vtransform);
        Pwc_wzb[i - 1] = ora_slevel(P_zetab, P_depth, c->SCW[i - 1], c->CSW[i - 1], P->hc, P->Vtransform);
        Pwc_wzc[i - 1] = ora_slevel(P_zetac, P_depth, c->SCW[i - 1], c->CSW[i - 1], P->hc, P->Vtransform);
        Pwc_wzf[i - 1] = ora_slevel(P_zetaf, P_depth, c->SCW[i - 1], c->CSW[i - 1], P->hc, P->Vtransform);
    }
    /* :949-951: index i == us+1 (== ws) receives the ws level */
    Pwc_wzb[i - 1] = ora_slevel(P_zetab, P_depth, c->SCW[ws - 1], c->CSW[ws - 1], P->hc, P->Vtransform);
    Pwc_wzc[i - 1] = ora_slevel(P_zetac, P_depth, c->SCW[ws - 1], c->CSW[ws - 1], P->hc, P->Vtransform);
    Pwc_wzf[i - 1] = ora_slevel(P_zetaf, P_depth, c->SCW[ws - 1], c->CSW[ws - 1], P->hc, P->Vtransform);

    double AdvectX, AdvectY, AdvectZ, TurbHx = 0.0, TurbHy = 0.0, TurbV = 0.0;
    double maxpartdepth = Pwc_wzb[0];                                    /* :981-987 */
    if (Pwc_wzc[0] > maxpartdepth) maxpartdepth = Pwc_wzc[0];
    if (Pwc_wzf[0] > maxpartdepth) maxpartdepth = Pwc_wzf[0];
    double minpartdepth = Pwc_wzb[ws - 1];
    if (Pwc_wzc[ws - 1] < minpartdepth) minpartdepth = Pwc_wzc[ws - 1];
    if (Pwc_wzf[ws - 1] < minpartdepth) minpartdepth = Pwc_wzf[ws - 1];

    double Uad, Vad, Wad, kn1_u, kn1_v, kn1_w, kn2_u, kn2_v, kn2_w, kn3_u, kn3_v, kn3_w, kn4_u, kn4_v, kn4_w;
    find_currents(c, &es, Xpar, Ypar, Zpar, Pwc_zb, Pwc_zc, Pwc_zf, Pwc_wzb, Pwc_wzc, Pwc_wzf,
                  P_zb, P_zc, P_zf, ex, ix, p, 1, &Uad, &Vad, &Wad);     /* :990 */
    kn1_u = Uad; kn1_v = Vad; kn1_w = Wad;
    double x1 = Xpar + (Uad * cos(P_angle) - Vad * sin(P_angle)) * (double)idt / 2.0;
    double y1 = Ypar + (Uad * sin(P_angle) + Vad * cos(P_angle)) * (double)idt / 2.0;
    double z1 = Zpar + Wad * (double)idt / 2.0;
    if (z1 > minpartdepth) z1 = minpartdepth - F32(0.000001);
    if (z1 < maxpartdepth) z1 = maxpartdepth + F32(0.000001);
    find_currents(c, &es, x1, y1, z1, Pwc_zb, Pwc_zc, Pwc_zf, Pwc_wzb, Pwc_wzc, Pwc_wzf,
                  P_zb, P_zc, P_zf, ex, ix, p, 2, &Uad, &Vad, &Wad);     /* :1006 */
    kn2_u = Uad; kn2_v = Vad; kn2_w = Wad;
    double x2 = Xpar + (Uad * cos(P_angle) - Vad * sin(P_angle)) * (double)idt / 2.0;
    double y2 = Ypar + (Uad * sin(P_angle) + Vad * cos(P_angle)) * (double)idt / 2.0;
    double z2 = Zpar + Wad * (double)idt / 2.0;
    if (z2 > minpartdepth) z2 = minpartdepth - F32(0.000001);
    if (z2 < maxpartdepth) z2 = maxpartdepth + F32(0.000001);
    find_currents(c, &es, x2, y2, z2, Pwc_zb, Pwc_zc, Pwc_zf, Pwc_wzb, Pwc_wzc, Pwc_wzf,
                  P_zb, P_zc, P_zf, ex, ix, p, 2, &Uad, &Vad, &Wad);     /* :1022 */
    kn3_u = Uad; kn3_v = Vad; kn3_w = Wad;
    double x3 = Xpar + (Uad * cos(P_angle) - Vad * sin(P_angle)) * (double)idt;
    double y3 = Ypar + (Uad * sin(P_angle) + Vad * cos(P_angle)) * (double)idt;
    double z3 = Zpar + Wad * (double)idt;
    if (z3 > minpartdepth) z3 = minpartdepth - F32(0.000001);
    if (z3 < maxpartdepth) z3 = maxpartdepth + F32(0.000001);
    find_currents(c, &es, x3, y3, z3, Pwc_zb, Pwc_zc, Pwc_zf, Pwc_wzb, Pwc_wzc, Pwc_wzf,
                  P_zb, P_zc, P_zf, ex, ix, p, 3, &Uad, &Vad, &Wad);     /* :1038 */
    kn4_u = Uad; kn4_v = Vad; kn4_w = Wad;
    double P_U = (kn1_u + 2.0 * kn2_u + 2.0 * kn3_u + kn4_u) / 6.0;     /* :1047-1053 */
    double P_V = (kn1_v + 2.0 * kn2_v + 2.0 * kn3_v + kn4_v) / 6.0;
    double P_W = (kn1_w + 2.0 * kn2_w + 2.0 * kn3_w + kn4_w) / 6.0;
    AdvectX = idt * (P_U * cos(P_angle) - P_V * sin(P_angle));
    AdvectY = idt * (P_U * sin(P_angle) + P_V * cos(P_angle));
    AdvectZ = idt * P_W;

    if (P->SaltTempOn) {                                                 /* :1062-1076 */
        for (i = 3; i <= us - 2; ++i)
            if (Zpar < Pwc_zb[i - 1] || Zpar < Pwc_zc[i - 1] || Zpar < Pwc_zf[i - 1]) break;
        int deplvl = i - 2;
        c->P_Salt[n] = WCTS_ITPI(c, &es, FLD_SALT, Xpar, Ypar, deplvl, Pwc_zb, Pwc_zc, Pwc_zf, P_zb, P_zc, P_zf, ex, ix, p, 4);
        c->P_Temp[n] = WCTS_ITPI(c, &es, FLD_TEMP, Xpar, Ypar, deplvl, Pwc_zb, Pwc_zc, Pwc_zf, P_zb, P_zc, P_zf, ex, ix, p, 4);
    }
    if (P->HTurbOn) HTurb(c, &g, &TurbHx, &TurbHy);                      /* :1087 */
    if (P->VTurbOn) TurbV = VTurb(c, &es, &g, P_zc, P_depth, P_zetac, p, ex, ix, Pwc_wzb, Pwc_wzc, Pwc_wzf);   /* :1098 */
    double XBehav = 0.0, YBehav = 0.0, ZBehav = 0.0; int bott = 0;
    if (P->Behavior != 0)                                                /* :1110 */
        behave(c, &es, &g, Xpar, Ypar, Zpar, Pwc_zb, Pwc_zc, Pwc_zf, P_zb, P_zc, P_zf, P_zetac, c->Age[n], P_depth,
               P_U, P_V, P_angle, n, it, ex, ix, ix[2] / 86400.0, p, &bott, &XBehav, &YBehav, &ZBehav);

    double newXpos = c->X[n] + AdvectX + TurbHx;                         /* :1128-1130 */
    double newYpos = c->Y[n] + AdvectY + TurbHy;
    double newZpos = c->Z[n] + AdvectZ + TurbV;
    double reflect;
    if (newZpos > P_zetac) { reflect = P_zetac - newZpos; newZpos = P_zetac + reflect; }    /* :1135-1138 */
    if (newZpos < P_depth) {                                             /* :1141-1145 */
        reflect = P_depth - newZpos; newZpos = P_depth + reflect;
        if (P->TrackCollisions) c->hitBottom[n]++;
    }
    newZpos = newZpos + ZBehav;                                          /* :1148 */
    if (P->Behavior == 7) {                                              /* :1150-1165 */
        if (bott) { newXpos = c->X[n]; newYpos = c->Y[n]; newZpos = P_depth; }
        else { newXpos = newXpos + XBehav; newYpos = newYpos + YBehav; newZpos = P_depth + P->Swimdepth; }
    }
    if (newZpos > P_zetac) newZpos = P_zetac - F32(0.000001);            /* :1170 */
    if (newZpos < P_depth) {                                             /* :1173-1176 */
        newZpos = P_depth + F32(0.000001);
        if (P->TrackCollisions) c->hitBottom[n]++;
    }
    double Xpos = c->X[n], Ypos = c->Y[n], nXpos = newXpos, nYpos = newYpos;   /* :1180-1232 */
    double fiX, fiY, frX, frY;
    int skipbound = -1, reflects = 0, waterFlag = 0, isWater = 0;
    for (;;) {
        int inter = ora_intersect_reflect(c, Xpos, Ypos, nXpos, nYpos, &fiX, &fiY, &frX, &frY, &skipbound, &isWater);
        if (inter == 0) break;
        if (P->TrackCollisions) c->hitLand[n]++;
        if (P->OpenOceanBoundary && isWater) {
            c->nX[n] = fiX; c->nY[n] = fiY; c->nZ[n] = newZpos;
            c->oob[n] = 1; waterFlag = 1; break;
        }
        reflects = reflects + 1;
        if (reflects > 3) {
            if (handle_error(c, n, LTGPU_EV_OUT_3RD, ix[2], c->Z[n])) return 1;
            waterFlag = 1; break;
        }
        Xpos = fiX; Ypos = fiY; nXpos = frX; nYpos = frY;
    }
    if (waterFlag) return 0;
    newXpos = nXpos; newYpos = nYpos;
    if (ora_mbounds(c, newYpos, newXpos) != 1)                           /* :1240-1272 */
        return handle_error(c, n, LTGPU_EV_OUT_MAIN, ix[2], c->Z[n]);
    double island;
    if (ora_ibounds(c, newYpos, newXpos, &island) == 1)                  /* :1275-1307 */
        return handle_error(c, n, LTGPU_EV_IN_ISLAND, ix[2], c->Z[n]);
    c->nX[n] = newXpos; c->nY[n] = newYpos; c->nZ[n] = newZpos;          /* :1312-1314 */
    ele_err = setEle(c, nXpos, nYpos, n, &es);                           /* :1317 */
    if (ele_err > 0) {
        int code = ele_err == 4 ? LTGPU_EV_JUMP_RHO : ele_err == 5 ? LTGPU_EV_JUMP_U : LTGPU_EV_JUMP_V;
        return handle_error(c, n, code, ix[2], c->Z[n]);
    }
    if (P->settlementon) {                                               /* :1373-1382, ledger 14 */
        int inpoly = testSettlement(c, c->Age[n], n, c->X[n], c->Y[n]);
        if (inpoly > 0) { c->nZ[n] = P_depth; c->endpoly[n] = inpoly; c->Lifespan[n] = c->Age[n]; }
    }
    return 0;
}

int32_t ora_step(ora_ctx* c, int32_t p, int32_t it)
{
    if (!c->X || c->npushed < 3) return LTGPU_E_ARG;
    double dt = (double)c->prm.dt, idt = (double)c->prm.idt;
    double ex[3] = { (double)((p - 2) * c->prm.dt), (double)((p - 1) * c->prm.dt), (double)(p * c->prm.dt) };   /* :568-571 */
    double ix[3] = { ex[1] + (double)((it - 2) * c->prm.idt), ex[1] + (double)((it - 1) * c->prm.idt),
                     ex[1] + (double)(it * c->prm.idt) };                                                       /* :588-590 */
    (void)dt; (void)idt;
    int stepIT = c->prm.dt / c->prm.idt;
    uint32_t gstep = (uint32_t)((p - 1) * stepIT + it);
    int stop = 0;
    if (c->nthreads > 1 && c->rng_mode == ORA_RNG_PHILOX) {
#pragma omp parallel for schedule(dynamic, 256) num_threads(c->nthreads) reduction(| : stop)
        for (int n = 0; n < c->n; ++n) { tl_nsig = 0; stop |= step_particle(c, n, p, it, ex, ix, gstep); c->nsig[n] += tl_nsig; }
    } else {
        for (int n = 0; n < c->n; ++n) {
            tl_nsig = 0;
            int st_ = step_particle(c, n, p, it, ex, ix, gstep);
            c->nsig[n] += tl_nsig;
            if (st_) { stop = 1; break; }   /* STOP */
        }
    }
    if (!stop) for (int n = 0; n < c->n; ++n) {                          /* :1407-1414 */
        c->X[n] = c->nX[n]; c->Y[n] = c->nY[n]; c->Z[n] = c->nZ[n];
    }
    c->last_ix3 = ix[2];
    return stop ? LTGPU_E_PARTICLE : LTGPU_OK;
}

int32_t ora_run_external(ora_ctx* c, int32_t p)
{
    int stepIT = c->prm.dt / c->prm.idt;                                 /* :554 */
    for (int it = 1; it <= stepIT; ++it) {
        int rc = ora_step(c, p, it);
        if (rc != LTGPU_OK) return rc;
    }
    return LTGPU_OK;
}

/* The start-up screen of ini_LTRANS (LTRANS.f90:356-452): mbounds, ibounds, then the element
 * check after setEle_all.  The reference reports only the first unlocated particle (its
 * setEle_all leaves the loop, hydro:1594); every one is reported here. */
int32_t ora_screen_initial(ora_ctx* c, int64_t counts[5], int64_t* bad_particle)
{
    int EF = c->prm.ErrorFlag;
    int64_t cnt[5] = {0, 0, 0, 0, 0};
    for (int n = 0; n < c->n; ++n) {
        int code = 0; double island = 0.0;
        if (ora_mbounds(c, c->Y[n], c->X[n]) == 0) code = LTGPU_EV_INIT_OUT_MAIN;               /* :362 */
        else if (ora_ibounds(c, c->Y[n], c->X[n], &island) == 1) code = LTGPU_EV_INIT_IN_ISLAND; /* :385 */
        else if (c->r_ele[n] == 0) code = LTGPU_EV_INIT_NOT_IN_RHO;                              /* :412-452 */
        else if (c->u_ele[n] == 0) code = LTGPU_EV_INIT_NOT_IN_U;
        else if (c->v_ele[n] == 0) code = LTGPU_EV_INIT_NOT_IN_V;
        if (!code) continue;
        int gid = (int)(c->first_id + n);
        if (EF < 1 || EF > 3) { if (c->bad_particle == 0 || gid < c->bad_particle) c->bad_particle = gid; }
        else if (EF == 2) c->dead[n] = 1;                                                        /* die */
        else c->oob[n] = 1;                                                                      /* setOut */
        add_event(c, gid, code, 0.0);
        cnt[code - LTGPU_EV_INIT_OUT_MAIN]++;
    }
    if (counts) memcpy(counts, cnt, sizeof cnt);
    if (c->bad_particle) { if (bad_particle) *bad_particle = c->bad_particle; return LTGPU_E_PARTICLE; }
    return LTGPU_OK;
}

/* conversion_module.f90:322-378, double-precision branches */
int32_t ora_fetch_lonlat(ora_ctx* c, int32_t spherical, double lonmin, double latmin, double earth_radius, double* lon, double* lat)
{
    const double pi = c->prm.PI, RCF = 180.0 / pi, R = earth_radius;
    for (int n = 0; n < c->n; ++n) {
        double x = c->X[n], y = c->Y[n];
        if (spherical) {
            double la = y * 180.0 / (R * pi) + latmin;
            lon[n] = x * 180.0 / (R * pi * cos(la / RCF)) + lonmin;
            lat[n] = y * RCF / R + latmin;
        } else {
            lon[n] = x / R * RCF;
            lat[n] = 2.0 * RCF * (atan(exp(y / R)) - pi / 4.0);
        }
    }
    return LTGPU_OK;
}

int32_t ora_fetch_sigerr(ora_ctx* c, int32_t* count)
{
    memcpy(count, c->nsig, sizeof(int32_t) * (size_t)c->n);
    return LTGPU_OK;
}

int32_t ora_sync(ora_ctx* c, int32_t* bad)
{
    if (bad) *bad = c->bad_particle;
    return c->bad_particle ? LTGPU_E_PARTICLE : LTGPU_OK;
}

static int status_of(const ora_ctx* c, int n)
{   /* getStatus behavior_module.f90:554-574 */
    int s = c->P_behave[n];
    if (c->dead[n]) s = -1;
    if (c->prm.settlementon && c->settle[n]) s = -2;
    if (c->prm.OpenOceanBoundary && c->oob[n]) s = -3;
    return s;
}

int32_t ora_fetch(ora_ctx* c, double* x, double* y, double* z, double* age, int32_t* status,
    double* salt, double* temp, int32_t* hitBottom, int32_t* hitLand, int32_t* endpoly, double* lifespan,
    int32_t* r_ele, int32_t* u_ele, int32_t* v_ele)
{
    size_t n = (size_t)c->n;
    if (x) memcpy(x, c->X, 8 * n);
    if (y) memcpy(y, c->Y, 8 * n);
    if (z) memcpy(z, c->Z, 8 * n);
    if (age) memcpy(age, c->Age, 8 * n);
    if (status) for (size_t i = 0; i < n; ++i) status[i] = status_of(c, (int)i);
    if (salt) memcpy(salt, c->P_Salt, 8 * n);
    if (temp) memcpy(temp, c->P_Temp, 8 * n);
    if (hitBottom) memcpy(hitBottom, c->hitBottom, 4 * n);
    if (hitLand) memcpy(hitLand, c->hitLand, 4 * n);
    if (endpoly) memcpy(endpoly, c->endpoly, 4 * n);
    if (lifespan) memcpy(lifespan, c->Lifespan, 8 * n);
    if (r_ele) memcpy(r_ele, c->r_ele, 4 * n);
    if (u_ele) memcpy(u_ele, c->u_ele, 4 * n);
    if (v_ele) memcpy(v_ele, c->v_ele, 4 * n);
    return LTGPU_OK;
}

int32_t ora_reset_hits(ora_ctx* c)
{   /* LTRANS.f90:1662-1665 */
    memset(c->hitBottom, 0, 4 * (size_t)c->n); memset(c->hitLand, 0, 4 * (size_t)c->n);
    return LTGPU_OK;
}

int32_t ora_stats(ora_ctx* c, int64_t counts[8])
{
    memset(counts, 0, 8 * sizeof(int64_t));
    for (int n = 0; n < c->n; ++n) {
        int settled = c->prm.settlementon && c->settle[n];
        int out = c->prm.OpenOceanBoundary && c->oob[n];
        counts[0] += settled; counts[1] += c->dead[n]; counts[2] += c->oob[n];
        counts[3] += c->hitLand[n]; counts[4] += c->hitBottom[n];
        int unborn = c->last_ix3 <= c->DOB[n];
        if (unborn) counts[7]++;
        else if (!settled && !(c->prm.mortality && c->dead[n]) && !out) counts[6]++;
    }
    counts[5] = c->nev;
    return LTGPU_OK;
}

static int ev_cmp(const void* a, const void* b)
{
    const ltgpu_event* x = (const ltgpu_event*)a; const ltgpu_event* y = (const ltgpu_event*)b;
    if (x->time != y->time) return x->time < y->time ? -1 : 1;
    return (x->particle > y->particle) - (x->particle < y->particle);
}
int32_t ora_drain_events(ora_ctx* c, ltgpu_event* buf, int32_t cap, int32_t* n)
{
    qsort(c->ev, (size_t)c->nev, sizeof(ltgpu_event), ev_cmp);
    int k = c->nev < cap ? c->nev : cap;
    memcpy(buf, c->ev, sizeof(ltgpu_event) * (size_t)k);
    memmove(c->ev, c->ev + k, sizeof(ltgpu_event) * (size_t)(c->nev - k));
    c->nev -= k; *n = k;
    return LTGPU_OK;
}
