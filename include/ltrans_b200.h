/*
 * ltrans_b200.h -- C ABI of the B200-native LTRANS v.2b particle-stepping library.
 *
 * This is the drop-in boundary for ONE path of the reference: the per-particle
 * internal time step `update_particles` (reference Model/LTRANS.f90:707-1419)
 * and everything it calls.  The reference has no FFI for this path (it is an
 * internal procedure of PROGRAM main), so the boundary is introduced here: every
 * entry point is `extern "C"`, takes plain pointers / sizes, and can be bound
 * from Fortran with an `INTERFACE ... BIND(C)` block.
 *
 * Conventions (all arrays are host memory owned by the caller; the library
 * copies on every set_ / push_ call and never keeps a host pointer):
 *   - Fortran column-major storage, 1-based node / element ids exactly as the
 *     reference stores them (hydrodynamic_module.f90:23-69).
 *   - double = real(c_double), int32_t = integer(c_int), float = real(c_float).
 *     Fortran LOGICAL is passed as int32_t (0 = .FALSE., non-zero = .TRUE.).
 *   - every function returns int32_t status, 0 = OK (see LTGPU_E_*).
 *   - there is NO CPU fallback: if no CUDA device is usable ltgpu_create fails.
 */
#ifndef LTRANS_B200_H
#define LTRANS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes ------------------------------------------------------ */
#define LTGPU_OK            0
#define LTGPU_E_ARG         1   /* bad argument / call order                   */
#define LTGPU_E_CUDA        2   /* CUDA runtime error (see ltgpu_last_error)   */
#define LTGPU_E_NODEVICE    3   /* no usable CUDA device -- no CPU fallback    */
#define LTGPU_E_PARTICLE    4   /* ErrorFlag==0 semantics: a particle hit a    */
                                /* STOP condition of LTRANS.f90:835-856 etc.   */
#define LTGPU_W_EVENTS_LOST 5   /* warning of ltgpu_drain_events: the device   */
                                /* event log overflowed, some ErrorLog lines   */
                                /* are missing (the drained ones are valid)    */

/* ---- per-particle event codes (ErrorLog.txt formats, LTRANS.f90:761-775) */
#define LTGPU_EV_INIT_OUT_MAIN   11  /* initially outside main bounds  (LTRANS.f90:377) */
#define LTGPU_EV_INIT_IN_ISLAND  12  /* initially inside island bounds (:400)            */
#define LTGPU_EV_INIT_NOT_IN_RHO 13  /* initially not in rho element   (:446)            */
#define LTGPU_EV_INIT_NOT_IN_U   14  /*                                (:448)            */
#define LTGPU_EV_INIT_NOT_IN_V   15  /*                                (:450)            */
#define LTGPU_EV_NOT_IN_RHO     21  /* setEle err 4 at step start  (:870)      */
#define LTGPU_EV_NOT_IN_U       22  /* setEle err 5                (:872)      */
#define LTGPU_EV_NOT_IN_V       23  /* setEle err 6                (:874)      */
#define LTGPU_EV_OUT_3RD        24  /* still out after 3rd reflection (:1221)  */
#define LTGPU_EV_OUT_MAIN       25  /* outside main bounds         (:1268)     */
#define LTGPU_EV_IN_ISLAND      26  /* inside island bounds        (:1303)     */
#define LTGPU_EV_JUMP_RHO       27  /* setEle err at new position  (:1355)     */
#define LTGPU_EV_JUMP_U         28  /*                             (:1357)     */
#define LTGPU_EV_JUMP_V         29  /*                             (:1359)     */

/* ---- RNG modes --------------------------------------------------------- */
/* The reference draws from one global sequential Mersenne Twister
 * (random_module.f90), so its turbulent trajectories depend on particle order and
 * count.  The device uses a counter-based stream instead:
 *   Philox4x32-10, key = (seed, 0),
 *   counter = (particle id low 32, id high 32, global internal step, block)
 *     global internal step = (p - 1) * (dt / idt) + it          (1-based)
 *     block 0              : HTurb, words 0,1 -> x deviate; 2,3 -> y deviate
 *     block 1 + i/2        : VTurb sub-step i (0-based), words 2(i%2), 2(i%2)+1
 *     block 0x80000000     : behave, words 0,1,2 in draw order
 *   uniforms: genrand_real3 = (w + 0.5)/2^32, genrand_real1 = w/(2^32 - 1)
 *   (random_module.f90:239-246, 213-220); norm() = sqrt(-2 ln u1) cos(2 PI u2)
 *   with the namelist PI (norm_module.f90:25-39).
 * particle id = first_id + local index, so results do not depend on sharding. */
#define LTGPU_RNG_PHILOX    1

/* field storage on the device */
#define LTGPU_F32           4   /* ROMS history is float32: lossless           */
#define LTGPU_F64           8

/*
 * Parameters actually read inside the particle loop.  Mirrors the namelists of
 * reference Model/LTRANS.h:45-269 (names kept).  POD, passed once.
 */
typedef struct ltgpu_params {
    int32_t numpar;            /* numparticles                     LTRANS.h:47  */
    int32_t dt;                /* external time step (s)           LTRANS.h:55  */
    int32_t idt;               /* internal time step (s)           LTRANS.h:56  */
    int32_t us;                /* rho s-levels                     LTRANS.h:63  */
    int32_t ws;                /* w s-levels                       LTRANS.h:64  */
    float   hc;                /* REAL(4) in the reference         LTRANS.h:66  */
    int32_t Vtransform;        /* 1,2,3                            LTRANS.h:68  */
    double  z0;                /* roughness                        LTRANS.h:67  */
    int32_t HTurbOn;
    int32_t VTurbOn;
    double  ConstantHTurb;
    int32_t Behavior;          /* 0..7                             LTRANS.h:104 */
    int32_t OpenOceanBoundary;
    int32_t mortality;
    int32_t settlementon;
    double  deadage, pediage, swimstart, swimslow, swimfast;
    double  Sgradient, sink, Hswimspeed, Swimdepth;
    double  twistart, twiend, daylength, Em, Kd, thresh;   /* dvmparam          */
    int32_t holesExist;
    int32_t seed;              /* LTRANS.h:253                                  */
    double  PI;                /* from the namelist, NOT machine pi             */
    int32_t ErrorFlag;         /* 0 stop, 1 revert, 2 kill, 3 set out           */
    int32_t SaltTempOn;
    int32_t TrackCollisions;
    int32_t FreeSlip;
    int32_t rng_mode;          /* LTGPU_RNG_PHILOX (device has no MT19937)      */
    int32_t field_dtype;       /* LTGPU_F32 or LTGPU_F64 device field storage   */
    int32_t vturb_window_sigs; /* 0 (default): reference semantics -- SIGS examines every interval of the
                                * 4*ws-knot VTurb fit and any SigErr sends the whole particle-step to
                                * linint (ver_turb:278-279, 300-336).  1 (opt-in approximation): only the
                                * 32 knots around the particle are examined, so the fall-back fires
                                * ~30 % less often than the reference's */
    int32_t vturb_fp32_walk;   /* 0 (default): the random-displacement walk in FP64 like the reference.
                                * 1 (opt-in): HVAL / HPVAL / Box-Muller of the 60 sub-steps in FP32; the
                                * spline fit, the Philox stream and the position stay FP64.  Statistical
                                * parity only (BASELINE north star, turbulent runs); never the headline */
} ltgpu_params;

/* one buffered per-particle event (drained in ascending particle id) */
typedef struct ltgpu_event {
    int32_t particle;          /* 1-based particle id n                         */
    int32_t code;              /* LTGPU_EV_*                                    */
    double  time;              /* ix(3) at the event, seconds                   */
} ltgpu_event;

typedef struct ltgpu_ctx ltgpu_ctx;

/* Create a context on CUDA device `device` (one context = one GPU = one rank).
 * Replaces nothing in the reference; called from ini_LTRANS after getParams
 * (LTRANS.f90:197). */
int32_t ltgpu_create(const ltgpu_params* prm, int32_t device, ltgpu_ctx** out);
int32_t ltgpu_destroy(ltgpu_ctx* ctx);
const char* ltgpu_last_error(const ltgpu_ctx* ctx);

/* Grid tables built by initGrid (hydrodynamic_module.f90:88-657).
 *  vi,uj = xi_rho,eta_rho ; ui,vj = xi_u,eta_v.
 *  rx,ry[rho_nodes] ux,uy[u_nodes] vx,vy[v_nodes]       metres   (:58)
 *  depth[rho_nodes] (h >= 0), angle[rho_nodes]                    (:38,58)
 *  rho_mask,u_mask,v_mask[nodes]   0/1                            (:69)
 *  SC,CS[us]  SCW,CSW[ws]                                         (:35)
 *  RE,UE,VE (4,n*E) 1-based node ids of each wet element          (:49)
 *  rAdj,uAdj,vAdj (n*E,10) column-major, col 1 = self, 0-filled   (:54,590-647)
 */
int32_t ltgpu_set_grid(ltgpu_ctx* ctx,
    int32_t vi, int32_t uj, int32_t ui, int32_t vj,
    const double* rx, const double* ry, const double* ux, const double* uy,
    const double* vx, const double* vy, const double* depth, const double* angle,
    const int32_t* rho_mask, const int32_t* u_mask, const int32_t* v_mask,
    const double* SC, const double* CS, const double* SCW, const double* CSW,
    const int32_t* RE, const int32_t* UE, const int32_t* VE,
    int32_t nRE, int32_t nUE, int32_t nVE,
    const int32_t* rAdj, const int32_t* uAdj, const int32_t* vAdj);

/* Boundary tables built by createBounds (boundary_module.f90:39-50,1037-1204).
 *  bnd_x,bnd_y (2,nbounds) segments; land[nbounds] (0 = open ocean)
 *  bx,by[maxbound] closed main polygon; hx,hy,hid[maxisland] closed islands. */
int32_t ltgpu_set_bounds(ltgpu_ctx* ctx,
    int32_t nbounds, const double* bnd_x, const double* bnd_y, const int32_t* land,
    int32_t maxbound, const double* bx, const double* by,
    int32_t maxisland, const double* hx, const double* hy, const int32_t* hid);

/* Habitat tables built by initSettlement (settlement_module.f90:29-33,40-480),
 * ragged lists flattened to CSR.
 *  polys (pedges,5) column-major: id, centre x, centre y, edge x, edge y
 *  holes (hedges,6) column-major: id, cx, cy, ex, ey, parent polygon id
 *  npoly / nhole distinct ids, listed in poly_id[] / hole_id[] with
 *    *_start (1-based first row) , *_size (rows) , *_maxdis (reject radius)
 *  elepoly_ptr[rho_elements+1], elepoly_idx[]: 0-based indices into poly_id[]
 *  polyhole_ptr[npoly+1], polyhole_idx[]:     0-based indices into hole_id[] */
int32_t ltgpu_set_habitat(ltgpu_ctx* ctx,
    int32_t pedges, const double* polys, int32_t hedges, const double* holes,
    int32_t npoly, const int32_t* poly_id, const int32_t* poly_start,
    const int32_t* poly_size, const double* poly_maxdis,
    int32_t nhole, const int32_t* hole_id, const int32_t* hole_start,
    const int32_t* hole_size, const double* hole_maxdis,
    const int32_t* elepoly_ptr, const int32_t* elepoly_idx,
    const int32_t* polyhole_ptr, const int32_t* polyhole_idx);

/* Particle table par(numpar,13) columns actually consumed (LTRANS.f90:104-120)
 * plus the element ids found by setEle_all (hydrodynamic_module.f90:1536).
 *  first_id : global 1-based id of local particle 0 (multi-GPU slices; keys the
 *             Philox stream so results do not depend on the sharding).
 *  r_ele,u_ele,v_ele may be NULL: the library then locates every particle on the
 *  device with the result of the whole-grid scan (setEle first=.TRUE.,
 *  hydro:1436-1457; 0 = in no element), through a bucket index over the elements. */
int32_t ltgpu_set_particles(ltgpu_ctx* ctx, int32_t n, int64_t first_id,
    const double* x, const double* y, const double* z, const double* dob,
    const int32_t* startpoly,
    const int32_t* r_ele, const int32_t* u_ele, const int32_t* v_ele);

/* The start-up screen of ini_LTRANS on the device (LTRANS.f90:356-452; replaces the
 * serial mbounds / ibounds loop and the check after setEle_all): particles
 * released outside the main boundary, inside an island or in no rho / u / v
 * element get die (ErrorFlag 2) or setOut (ErrorFlag 1, 3) and one event
 * LTGPU_EV_INIT_* with time 0; counts[0..4] = how many of each code 11..15.
 * ErrorFlag outside 1..3: returns LTGPU_E_PARTICLE and the lowest offending id
 * (the reference STOPs).  Unlike the reference, every unlocated particle is
 * reported, not only the first (its setEle_all leaves the rest unset, hydro:1594).
 * Call after set_bounds and set_particles; optional. */
int32_t ltgpu_screen_initial(ltgpu_ctx* ctx, int64_t counts[5], int64_t* bad_particle);

/* Queue the next hydro record (ROMS memory order: node fastest, then level;
 * exactly one NF90_GET_VAR record, hydro:1140-1364).  The library applies the
 * mask multiply of hydro:1371-1403 and copies asynchronously on a side stream.
 * The first three pushes become back/centre/forward (initHydro, hydro:991-1043);
 * later pushes land in the spare slot until ltgpu_rotate_hydro.
 * dtype = LTGPU_F32 or LTGPU_F64 (type of the host arrays). salt/temp may be NULL. */
int32_t ltgpu_push_hydro(ltgpu_ctx* ctx, int32_t dtype,
    const void* zeta, const void* u, const void* v, const void* w,
    const void* aks, const void* salt, const void* temp);
/* updateHydro's index rotation t_b,t_c,t_f (hydro:1080-1082): the record pushed
 * last becomes "forward".  The compute stream waits for its copy. */
int32_t ltgpu_rotate_hydro(ltgpu_ctx* ctx);

/* One internal time step for every particle = one call of update_particles
 * (LTRANS.f90:596) with external step p (1-based) and internal step it (1-based).
 * Asynchronous. */
int32_t ltgpu_step(ltgpu_ctx* ctx, int32_t p, int32_t it);
/* All dt/idt internal steps of external step p (LTRANS.f90:573-577). */
int32_t ltgpu_run_external(ltgpu_ctx* ctx, int32_t p);
/* Wait for queued work; returns LTGPU_E_PARTICLE if a particle hit a STOP
 * condition under ErrorFlag outside 1..3 (lowest id in *bad_particle). */
int32_t ltgpu_sync(ltgpu_ctx* ctx, int32_t* bad_particle);

/* Gather particle state (printOutput LTRANS.f90:1617, fin_LTRANS :632-658).
 * Any pointer may be NULL.  status = getStatus (behavior_module.f90:554-574).
 * Destinations in page-locked host memory (cudaHostAlloc / cudaHostRegister by the caller)
 * receive the device copy directly; pageable ones go through the library's pinned bounce
 * buffers (double buffered: the copy of one column overlaps the memcpy of the previous one). */
int32_t ltgpu_fetch(ltgpu_ctx* ctx,
    double* x, double* y, double* z, double* age, int32_t* status,
    double* salt, double* temp, int32_t* hitBottom, int32_t* hitLand,
    int32_t* endpoly, double* lifespan,
    int32_t* r_ele, int32_t* u_ele, int32_t* v_ele);
/* Particle positions as longitude / latitude for the writers: x2lon(x, y) and y2lat(y) of
 * conversion_module.f90:322-378 (double-precision branches; `spherical` = SphericalProjection,
 * lonmin / latmin as stored by CONVERT_MOD, i.e. the namelist values minus 1,
 * parameter_module.f90:133-134; PI from ltgpu_params).  Saves the host its per-particle
 * conversion loop of writeOutput (LTRANS.f90:1712-1716). */
int32_t ltgpu_fetch_lonlat(ltgpu_ctx* ctx, int32_t spherical, double lonmin, double latmin,
                           double earth_radius, double* lon, double* lat);

/* Diagnostic (no reference counterpart): how many times each particle's water-column
 * profile fell back from the tension spline to linint because SIGS raised SigErr
 * (hydro:2621-2644, ver_turb:300-336).  The reference takes that branch silently; its
 * verdict hangs on the last bits of T = max(D1/D2, D2/D1) and of exp(), so two correct
 * implementations disagree on WHICH steps fall back (about 4e-4 per particle-step) while
 * agreeing on the rate; the parity tests use this count to show that every trajectory
 * difference above 1e-9 belongs to a particle whose fall-back history differs. */
int32_t ltgpu_fetch_sigerr(ltgpu_ctx* ctx, int32_t* count);

/* Reset hitBottom/hitLand after a print (LTRANS.f90:1662-1665). */
int32_t ltgpu_reset_hits(ltgpu_ctx* ctx);

/* counts[0..7] = settled, dead, out of bounds, land hits, bottom hits,
 * buffered events, active (released, still tracked), unborn.  Device reduction;
 * multi-GPU callers all-reduce these 8 int64 (NCCL sum). */
int32_t ltgpu_stats(ltgpu_ctx* ctx, int64_t counts[8]);

/* Drain buffered per-particle events (by time step, then ascending particle id = the order
 * the serial loop writes ErrorLog.txt, LTRANS.f90:761-775).  Call until *n < cap.
 * The device log holds max(2^20, numpar) events and is emptied into host memory at every
 * ltgpu_sync / ltgpu_drain_events, so it cannot overflow when the host synchronises at least
 * once per internal step; with fewer synchronisations (ltgpu_run_external) an overflow drops
 * the newest events, which is reported ONCE by the return value LTGPU_W_EVENTS_LOST (buf and
 * *n are still valid) and counted by ltgpu_events_lost. */
int32_t ltgpu_drain_events(ltgpu_ctx* ctx, ltgpu_event* buf, int32_t cap, int32_t* n);
int32_t ltgpu_events_lost(ltgpu_ctx* ctx, int64_t* lost);

/* Device pointers of the particle state for device-side gathers (NCCL output
 * gather without a host bounce).  which: 0=x 1=y 2=z 3=age 4=status. */
int32_t ltgpu_device_ptr(ltgpu_ctx* ctx, int32_t which, void** dptr);

/* Output gather without a host bounce (ncclAllGather of the print-interval
 * columns): writes column `which` (0=x 1=y 2=z 3=age as double, 4=status as int32,
 * getStatus behavior_module.f90:554-574) in PARTICLE order into caller-owned DEVICE memory
 * (numpar elements), asynchronously on the compute stream (ltgpu_stream). */
int32_t ltgpu_export_device(ltgpu_ctx* ctx, int32_t which, void* dst_device);

/* Measured FP64 FMA rate of the device in TFLOP/s (8 independent chains per thread, best of
 * 4 launches): the compute roof this FP64 path is reported against next to the HBM one. */
int32_t ltgpu_fp64_peak(ltgpu_ctx* ctx, double* tflops);

/* Timing helpers used by bench.py: CUDA events recorded on the compute stream
 * (torch.cuda.Event only sees torch's current stream). */
int32_t ltgpu_timer_start(ltgpu_ctx* ctx);
int32_t ltgpu_timer_stop(ltgpu_ctx* ctx, float* ms);
/* Per-kernel device times of the step (k_advect, k_vturb, k_finish, reserved), summed over
 * the internal steps since the last call; enable != 0 switches the CUDA-event bracketing on
 * (it synchronises every step, so it is a measurement mode, not for production runs). */
int32_t ltgpu_kernel_times(ltgpu_ctx* ctx, int32_t enable, float ms[4], int64_t* steps);
/* number of kernels this context has launched so far */
int64_t ltgpu_launch_count(const ltgpu_ctx* ctx);
/* raw cudaStream_t of the compute stream */
void* ltgpu_stream(ltgpu_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* LTRANS_B200_H */
