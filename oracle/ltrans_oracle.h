/*
 * ltrans_oracle.h -- CPU restatement of the LTRANS v.2b per-particle time step.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: it
 * may be imported / linked / executed only by tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs, and only as the checker
 * or as the timed CPU baseline.  The product path (ltransv.2b_b200/) never
 * calls it and fails loudly when its CUDA library is missing.
 *
 * PARITY UNPINNED BY THE REFERENCE: the reference (Fortran 90 + NetCDF) has no
 * tests, golden vectors or fixtures for this path and cannot be compiled in this
 * environment (no Fortran compiler, no NetCDF).  This file restates the cited
 * lines of /root/reference/Model/ function by function; what pins it is
 * (1) the published MT19937 known answers for random_module.f90,
 * (2) an independent NumPy restatement of the leaf numerics (oracle/leaf_numpy.py),
 * (3) analytic invariants (tests/test_oracle_invariants.py).
 *
 * The API mirrors include/ltrans_b200.h call for call so that the parity tests
 * drive both with the same arrays.
 */
#ifndef LTRANS_ORACLE_H
#define LTRANS_ORACLE_H
#include <stdint.h>
#include "../include/ltrans_b200.h"   /* ltgpu_params, ltgpu_event, codes */

#ifdef __cplusplus
extern "C" {
#endif

#define ORA_RNG_PHILOX 1   /* same keyed stream as the device                     */
#define ORA_RNG_MT     2   /* the reference's global sequential MT19937           */

typedef struct ora_ctx ora_ctx;

int32_t ora_create(const ltgpu_params* prm, ora_ctx** out);
int32_t ora_destroy(ora_ctx* c);
int32_t ora_set_threads(ora_ctx* c, int32_t nthreads);   /* OpenMP over particles (Philox only) */

int32_t ora_set_grid(ora_ctx* c,
    int32_t vi, int32_t uj, int32_t ui, int32_t vj,
    const double* rx, const double* ry, const double* ux, const double* uy,
    const double* vx, const double* vy, const double* depth, const double* angle,
    const int32_t* rho_mask, const int32_t* u_mask, const int32_t* v_mask,
    const double* SC, const double* CS, const double* SCW, const double* CSW,
    const int32_t* RE, const int32_t* UE, const int32_t* VE,
    int32_t nRE, int32_t nUE, int32_t nVE,
    const int32_t* rAdj, const int32_t* uAdj, const int32_t* vAdj);

int32_t ora_set_bounds(ora_ctx* c,
    int32_t nbounds, const double* bnd_x, const double* bnd_y, const int32_t* land,
    int32_t maxbound, const double* bx, const double* by,
    int32_t maxisland, const double* hx, const double* hy, const int32_t* hid);

int32_t ora_set_habitat(ora_ctx* c,
    int32_t pedges, const double* polys, int32_t hedges, const double* holes,
    int32_t npoly, const int32_t* poly_id, const int32_t* poly_start,
    const int32_t* poly_size, const double* poly_maxdis,
    int32_t nhole, const int32_t* hole_id, const int32_t* hole_start,
    const int32_t* hole_size, const double* hole_maxdis,
    const int32_t* elepoly_ptr, const int32_t* elepoly_idx,
    const int32_t* polyhole_ptr, const int32_t* polyhole_idx);

int32_t ora_set_particles(ora_ctx* c, int32_t n, int64_t first_id,
    const double* x, const double* y, const double* z, const double* dob,
    const int32_t* startpoly,
    const int32_t* r_ele, const int32_t* u_ele, const int32_t* v_ele);

int32_t ora_push_hydro(ora_ctx* c, int32_t dtype,
    const void* zeta, const void* u, const void* v, const void* w,
    const void* aks, const void* salt, const void* temp);
int32_t ora_rotate_hydro(ora_ctx* c);

int32_t ora_step(ora_ctx* c, int32_t p, int32_t it);
int32_t ora_run_external(ora_ctx* c, int32_t p);
int32_t ora_screen_initial(ora_ctx* c, int64_t counts[5], int64_t* bad_particle);
int32_t ora_fetch_lonlat(ora_ctx* c, int32_t spherical, double lonmin, double latmin, double earth_radius, double* lon, double* lat);
int32_t ora_fetch_sigerr(ora_ctx* c, int32_t* count);
int32_t ora_sync(ora_ctx* c, int32_t* bad_particle);

int32_t ora_fetch(ora_ctx* c,
    double* x, double* y, double* z, double* age, int32_t* status,
    double* salt, double* temp, int32_t* hitBottom, int32_t* hitLand,
    int32_t* endpoly, double* lifespan,
    int32_t* r_ele, int32_t* u_ele, int32_t* v_ele);
int32_t ora_reset_hits(ora_ctx* c);
int32_t ora_stats(ora_ctx* c, int64_t counts[8]);
int32_t ora_drain_events(ora_ctx* c, ltgpu_event* buf, int32_t cap, int32_t* n);

/* ---- leaf functions exported for differential / invariant tests ---------- */
double  ora_polintd(const double xa[3], const double ya[3], double x);
void    ora_linint(const double* xa, const double* ya, int32_t n, double x, double* y, double* m);
int32_t ora_gridcell(const double ex[4], const double ey[4], double X, double Y);
int32_t ora_inpoly(double x, double y, int32_t n, const double* ex, const double* ey, int32_t onin);
void    ora_tspsi(int32_t n, const double* x, const double* y, double* yp, double* sigma,
                  int32_t* ier, int32_t* sigerr);
double  ora_hval(double t, int32_t n, const double* x, const double* y, const double* yp,
                 const double* sigma, int32_t* ier);
double  ora_hpval(double t, int32_t n, const double* x, const double* y, const double* yp,
                  const double* sigma, int32_t* ier);
void    ora_snhcsh(double x, double* sinhm, double* coshm, double* coshmm);
double  ora_slevel(double zeta, double depth, double sc, double cs, float hc, int32_t vtransform);
/* intersect_reflect against the context's boundary table; returns intersectf */
int32_t ora_intersect_reflect(ora_ctx* c, double Xpos, double Ypos, double nXpos, double nYpos,
                              double* fiX, double* fiY, double* frX, double* frY,
                              int32_t* skipbound, int32_t* isWater);
int32_t ora_mbounds(ora_ctx* c, double Ypos, double Xpos);
int32_t ora_ibounds(ora_ctx* c, double claty, double clongx, double* island);

/* MT19937 restatement of random_module.f90 (KAT-checked) */
void     ora_mt_init_genrand(uint32_t s);
void     ora_mt_init_by_array(const uint32_t* key, int32_t len);
uint32_t ora_mt_int32(void);
double   ora_mt_real1(void);
double   ora_mt_real3(void);
/* Philox4x32-10 block: ctr[4], key[2] -> out[4] */
void     ora_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* select ORA_RNG_PHILOX (default) or ORA_RNG_MT for ora_step */
int32_t  ora_set_rng(ora_ctx* c, int32_t mode);

#ifdef __cplusplus
}
#endif
#endif
