// lt_types.h -- device-side view of one context (kernel argument, lives in the
// constant bank).  Layouts are chosen for the B200 memory system, not copied from
// the reference's (3, node, level) Fortran arrays:
//   * hydro fields: [node][level][4 ring slots] -- the three time levels a
//     particle needs for one (node, level) sit in ONE 16-byte (f32) / 32-byte (f64)
//     chunk, so a 4-corner x 4-level stencil is 16 LDG.128 instead of 48 scalar
//     gathers, and the 4th component is the slot being refilled by the copy stream.
//   * element tables: corner coordinates packed as 8 doubles per element (two
//     32-byte sectors); adjacency row-major [element][10].
//   * particles: structure of arrays, one thread per particle.
#pragma once
#include <stdint.h>
#include "../../include/ltrans_b200.h"

#define LT_MAXLEV 64            // us <= 63, ws <= 64

struct LtGridTab {              // one of the rho / u / v grids
    const double* ele;          // [nE][8]  x0..x3,y0..y3 of the element corners
    const int4*   node;         // [nE]     0-based node ids of the corners (RE/UE/VE - 1)
    const int*    adj;          // [nE][10] 1-based neighbour element ids, 0 = none
    const uint8_t* mask;        // [nodes]  only read when FreeSlip
    int nE, nodes;
    int convex;                 // every element is a convex, non-degenerate quad (checked by set_grid)
    // bucket index over element bounding boxes for the whole-grid search (built by set_grid)
    double lx0, ly0, lrcs; int lnx, lny; const int *lptr, *lidx;
};

struct LtDev {
    ltgpu_params P;
    LtGridTab R, U, V;
    const double* depth;        // [rho_nodes]
    const double* angle;        // [rho_nodes]
    // fields: [node][level][4]; element type float or double (P.field_dtype)
    const void *zeta, *u, *v, *w, *kh, *salt, *temp;
    int sb, sc, sf;             // ring slot holding the back / centre / forward record
    // boundary
    const double4* seg;         // [nbounds] x1,y1,x2,y2
    const uint8_t* land;        // [nbounds]
    const double2* bxy;         // [maxbound] closed main polygon
    const double2* hxy;         // [maxisland] closed island polygons, concatenated
    const int* hid;             // [maxisland]
    int nbounds, maxbound, maxisland;
    // habitat (CSR)
    const double* polys; const double* holes; int pedges, hedges, npoly, nhole;
    const int *poly_start, *poly_size, *hole_start, *hole_size;
    const double *poly_maxdis, *hole_maxdis;
    const int *elepoly_ptr, *elepoly_idx, *polyhole_ptr, *polyhole_idx;
    // particles (SoA)
    double *x, *y, *z, *age, *dob, *lifespan, *psalt, *ptemp, *timer, *sprev, *zprev;
    int *r_ele, *u_ele, *v_ele, *hitB, *hitL, *endpoly;
    int *nsig;                  // [n] diagnostic: SigErr (linint) fall-backs taken so far
    uint8_t* flags;             // bit0 settled, bit1 dead, bit2 oob, bit3 bottom (behaviour 7)
    int8_t* behave;             // P_behave
    int n; long long first_id;
    const int* pid;             // slot -> local particle index (particles are periodically re-sorted by cell)
    // events
    ltgpu_event* ev; int* nev; int evcap; int* bad;
    // per-particle scratch passed between the three step kernels (SoA)
    double *s_depth, *s_angle, *s_zeb, *s_zec, *s_zef, *s_pzb, *s_pzc, *s_pzf, *s_zpar,
           *s_nx, *s_ny, *s_advz, *s_pu, *s_pv, *s_turbv;
    uint8_t* s_act;
    // VTurb scratch handed from k_vbuild to k_vwalk for one chunk of particles (lt_vturb.cuh):
    // vw[row][i]: rows 0..VW-1 knot values, VW..2VW-1 knot slopes, 2VW..3VW-1 tension factors of the
    // window of chunk-local particle i; vz1 / vzn the knot line; vka = first knot | SigErr << 16
    double *vw, *vz1, *vzn; int* vka; int vw_stride;
    // the order in which the VTurb kernels visit the slots (finer depth bins than the slot order itself, see
    // resort() in ltrans_b200.cu); nullptr = slot order
    const int* vorder;
    int vb_slot_order;          // 1: k_vbuild keeps the slot order (scratch column = slot), only k_vwalk follows vorder
    // Lagrange form of the 3-point time polynomial (interpolation_module.f90:70-107),
    // LW[v][t] = weight of hydro record t (b,c,f) at internal time ix(v), with the p == 1
    // triplet (b,b,c) folded in; LW4 = (LW[0] + 4 LW[1] + LW[2]) / 6; LWz = raw weights at ix(2)
    double LW[3][3], LW4[3], LWz[3];
    // exact spatial indices over the boundary tables (built by set_bounds)
    double sg_x0, sg_y0, sg_rcs; int sg_nx, sg_ny; const int *sg_ptr, *sg_idx;     // segment buckets
    double mb_y0, mb_rbh; int mb_n; const int *mb_ptr, *mb_idx;                     // main polygon y-bands
    double ib_y0, ib_rbh; int ib_n; const int *ib_ptr, *ib_idx; int ib_ok;          // island y-bands
#ifdef LT_DEBUG_TRACE
    double* dbg; long long dbg_id;   // debug builds only: per-substep VTurb trace of one particle
#endif
    // this step
    int p, it; unsigned gstep;
    double ex[3], ix[3];
    double SC[LT_MAXLEV], CS[LT_MAXLEV], SCW[LT_MAXLEV], CSW[LT_MAXLEV];
};

#define LT_F_SETTLED 1
#define LT_F_DEAD    2
#define LT_F_OOB     4
#define LT_F_BOTTOM  8
