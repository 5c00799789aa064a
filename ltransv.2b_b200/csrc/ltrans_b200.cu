// ltrans_b200.cu -- C ABI (include/ltrans_b200.h) + kernels of the B200-native
// LTRANS v.2b particle step.  Build: nvcc -gencode arch=compute_100a,code=sm_100a.
//
// One context = one GPU.  Streams: `compute` runs the particle step; `copy` refills
// the spare hydro ring slot (pinned staging -> H2D -> transpose/mask kernel) while the
// loop runs; an event fences the slot before the step that first reads it.
// There is no CPU fallback anywhere in this file.
#include <cuda_runtime.h>
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <string>
#include <vector>
#include "lt_step2.cuh"
#include "lt_vturb.cuh"
#include <cub/device/device_radix_sort.cuh>

// ------------------------------------------------------------------ kernels --
// Block shapes (threads, min resident blocks) per kernel, from the sweep in profiles/r01_notes.md:
// large blocks keep the warps of a block in step (same code region -> fewer instruction-fetch
// stalls) and, after the cell re-sort, on neighbouring particles.
#ifndef LT_BLK_ADV
#define LT_BLK_ADV 512
#endif
#ifndef LT_MIN_ADV
#define LT_MIN_ADV 1
#endif
#ifndef LT_BLK_VT
#define LT_BLK_VT 384
#endif
#ifndef LT_MIN_VT
#define LT_MIN_VT 2
#endif
#ifndef LT_BLK_FIN
#define LT_BLK_FIN 128
#endif
#ifndef LT_MIN_FIN
#define LT_MIN_FIN 4
#endif
template <class T, int PH>
__global__ void __launch_bounds__(LT_BLK_ADV, LT_MIN_ADV) k_advect(const __grid_constant__ LtDev D)
{
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    lev_tables_load(D);
    AdvS S;
    bool act = n < D.n;
    if (act) act = advect_prologue<T, PH>(D, n, S);
#pragma unroll 1
    for (int stg = 0; stg < 4; ++stg) {
#ifndef LT_NO_STAGE_BARRIER
        __syncthreads();
#endif
        if (act) advect_stage<T, PH>(D, S, stg);
    }
    if (act) advect_epilogue<T, PH>(D, n, S);
}
template <class T, int PH>
__global__ void __launch_bounds__(LT_BLK_VT, LT_MIN_VT) k_vturb(const __grid_constant__ LtDev D)
{
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    lev_tables_load(D);
    if (n < D.n) vturb_particle<T, PH>(D, n);
}
// VTurb as fit + walk (lt_vturb.cuh): k_vbuild spreads the 32 lanes of a warp over ONE particle's water
// column at a time (shared-memory scratch per warp, no block-level synchronisation), k_vwalk is one
// thread per particle.
#ifndef LT_BLK_VB
#define LT_BLK_VB 128
#endif
#ifndef LT_MIN_VB
#define LT_MIN_VB 5
#endif
#ifndef LT_BLK_VW
#define LT_BLK_VW 256
#endif
#ifndef LT_MIN_VW
#define LT_MIN_VW 3
#endif
// MINB: resident blocks the register allocation is sized for.  Deep columns (ws > ~30) leave room for 4 blocks
// of shared memory only, so the build for them takes the registers of the fifth (118 instead of 96: +3.4 %).
template <class T, int PH, int MINB>
__global__ void __launch_bounds__(LT_BLK_VB, MINB) k_vbuild(const __grid_constant__ LtDev D, int base, int count)
{
    extern __shared__ double vb_smem[];
    lev_tables_load(D);
    const int wib = threadIdx.x >> 5;
    const int warp_first = base + (blockIdx.x * (LT_BLK_VB / 32) + wib) * 32;
    if (warp_first >= base + count) return;
    vbuild_warp<T, PH>(D, vb_smem + (size_t)wib * vb_smem_doubles(D.P.ws), warp_first, base, count);
}
template <class T, int PH>
__global__ void __launch_bounds__(LT_BLK_VW, LT_MIN_VW) k_vwalk(const __grid_constant__ LtDev D, int base, int count)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    lev_tables_load(D);                                                 // the refit of a walked-out window
    if (i < count) { const int n = D.vorder ? D.vorder[base + i] : base + i; vwalk_particle<T, PH>(D, n, D.vb_slot_order ? n - base : i); }
}
template <class T, int PH>
__global__ void __launch_bounds__(LT_BLK_VW, LT_MIN_VW) k_vwalk_f32(const __grid_constant__ LtDev D, int base, int count)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    lev_tables_load(D);
    if (i < count) { const int n = D.vorder ? D.vorder[base + i] : base + i; vwalk_particle_f32<T, PH>(D, n, D.vb_slot_order ? n - base : i); }
}
template <class T, int PH>
__global__ void __launch_bounds__(LT_BLK_FIN, LT_MIN_FIN) k_finish(const __grid_constant__ LtDev D)
{
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    lev_tables_load(D);
    if (n < D.n) finish_particle<T, PH>(D, n);
}

// one ROMS record [level][node] (+ mask multiply, hydro:1371-1403) -> ring slot of the
// [node][level][4] device layout.  Block = 32 nodes x 8 levels through shared memory so
// both the read (node-fastest) and the write (level-fastest) are coalesced.
template <class TI, class TO>
__global__ void k_fill_slot(const TI* __restrict__ in, const uint8_t* __restrict__ mask, TO* __restrict__ out,
                            int nodes, int L, int slot)
{
    __shared__ double tile[32][33];
    int n0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    for (int kk = threadIdx.y; kk < 32; kk += blockDim.y) {
        int n = n0 + threadIdx.x, k = k0 + kk;
        if (n < nodes && k < L) tile[kk][threadIdx.x] = (double)in[(size_t)k * nodes + n] * (double)mask[n];
    }
    __syncthreads();
    for (int nn = threadIdx.y; nn < 32; nn += blockDim.y) {
        int n = n0 + nn, k = k0 + threadIdx.x;
        if (n < nodes && k < L) out[((size_t)n * L + k) * 4 + slot] = (TO)tile[threadIdx.x][nn];
    }
}

// ---- re-sort of the particle slots by (depth bin, rho element), every internal step ---------
// Particles never interact, so slot order is free.  What pays in this FP64 code is lanes that
// take the same branches: level window and log-layer test of find_currents, spline interval and
// tension regime, level walks of the VTurb fit, Newton iteration counts.  Those follow the
// particle's relative depth, so the depth bin is the MAJOR key (32 bins of the local water
// column) and the rho element the minor one; within a bin neighbouring lanes still sit in
// neighbouring elements, which keeps the stencil gathers local.  Vertical turbulence scrambles
// the bins within a few internal steps, so the sort runs before every step: key build, CUB
// radix sort of (key, slot) and ONE kernel that moves all 21 state columns (122 B per particle)
// cost 0.07 ms per step at 1 M particles.  Measured on the 130x130x20 benchmark (particle-steps/s):
// element-major key once per external step 128 M; depth-major once per external step 132 M,
// every 10 / 5 / 2 / 1 steps 141 / 146 / 153 / 157 M; 32 bins and the fused move 163 M.  At Gulf
// scale (ws = 37, 12.5 M particles) depth-major takes `k_vturb` from 64 to 41 ms per step.
// From 2 M particles on the slots take 8 bins and the VTurb kernels their own visiting order of 64 bins (resort()).
// LTGPU_SORT=0 disables, LTGPU_SORT_MODE=<bins> (0 = element-major), LTGPU_VT_BINS=<bins> (0 = slot order),
// LTGPU_SORT_EVERY=<steps>.
// Key of the depth-major orders: (bin << ebits) | element with ebits = bits of the element count, idle particles in
// bin `bins` (one past the last): as few key bits as the grid needs, so the radix sort runs 3 passes of 8 bits
// instead of 4 at Gulf scale (20 + 4 bits).
__global__ void k_sort_keys(const LtDev D, unsigned* __restrict__ key, int* __restrict__ idx, int bins, int ebits)
{
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= D.n) return;
    int re = D.r_ele[n];
    int4 nd = __ldg(D.R.node + (max(re, 1) - 1));
    double h = 0.25 * (__ldg(D.depth + nd.x) + __ldg(D.depth + nd.y) + __ldg(D.depth + nd.z) + __ldg(D.depth + nd.w));
    bool idle = (D.flags[n] & (LT_F_SETTLED | LT_F_DEAD | LT_F_OOB)) != 0;
    idx[n] = n;
    if (bins > 0) {
        int b = (int)(-(double)bins * D.z[n] / fmax(h, 1e-3)); b = max(0, min(bins - 1, b));
        key[n] = ((unsigned)(idle ? bins : b) << ebits) | (unsigned)max(re, 0);   // inactive particles go last
        return;
    }
    if (idle) { key[n] = 0xffffffffu; return; }
    {
        int b = (int)(-8.0 * D.z[n] / fmax(h, 1e-3)); b = max(0, min(7, b));
        key[n] = ((unsigned)re << 3) | (unsigned)b;
    }
}
template <class V>
__global__ void k_gather(const V* __restrict__ in, V* __restrict__ out, const int* __restrict__ perm, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[perm[i]];
}
template <class V>
__global__ void k_scatter(const V* __restrict__ in, V* __restrict__ out, const int* __restrict__ pid, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[pid[i]] = in[i];
}
__global__ void k_iota(int* p, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
}

__global__ void k_status(const LtDev D, int* __restrict__ status)
{   // getStatus behavior_module.f90:554-574
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= D.n) return;
    int s = D.behave[n]; uint8_t f = D.flags[n];
    if (f & LT_F_DEAD) s = -1;
    if (D.P.settlementon && (f & LT_F_SETTLED)) s = -2;
    if (D.P.OpenOceanBoundary && (f & LT_F_OOB)) s = -3;
    status[n] = s;
}

__global__ void k_stats(const LtDev D, double last_ix3, unsigned long long* __restrict__ out)
{
    unsigned long long c[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < D.n; n += gridDim.x * blockDim.x) {
        uint8_t f = D.flags[n];
        bool settled = D.P.settlementon && (f & LT_F_SETTLED), out_ = D.P.OpenOceanBoundary && (f & LT_F_OOB);
        c[0] += settled; c[1] += (f & LT_F_DEAD) != 0; c[2] += (f & LT_F_OOB) != 0;
        c[3] += (unsigned)D.hitL[n]; c[4] += (unsigned)D.hitB[n];
        if (last_ix3 <= D.dob[n]) c[7]++;
        else if (!settled && !(D.P.mortality && (f & LT_F_DEAD)) && !out_) c[6]++;
    }
    for (int k = 0; k < 8; ++k) {
        unsigned long long v = c[k];
        for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(out + k, v);
    }
}

// setEle first=.TRUE. (hydro:1436-1457): whole-grid scan, one thread per particle.
// gridcell without checkele keeps scanning after an on-edge hit whose crossing total is
// even (the inner `exit`), so the answer is the first definite hit, else the last soft one;
// a point on a shared edge is a definite hit of the lower-numbered element, which is what
// a first-hit scan returns, so the two coincide except for degenerate (zero-area) elements.
__global__ void k_locate(const LtDev D, LtGridTab G, int* __restrict__ ele)
{
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= D.n) return;
    double X = D.x[n], Y = D.y[n];
    int found = 0;
    // an element that holds the point has the point inside its bounding box, so it is listed in
    // the point's bucket; bucket lists ascend, so the first hit is the full scan's first hit
    int cx = (int)floor((X - G.lx0) * G.lrcs), cy = (int)floor((Y - G.ly0) * G.lrcs);
    if (cx >= 0 && cy >= 0 && cx < G.lnx && cy < G.lny) {
        int c = cy * G.lnx + cx;
        for (int q = __ldg(G.lptr + c); q < __ldg(G.lptr + c + 1); ++q) {
            int e = __ldg(G.lidx + q);
            if (gridcell_any(G, e, X, Y)) { found = e + 1; break; }
        }
    }
    ele[n] = found;
}

// The start-up screen of ini_LTRANS (LTRANS.f90:356-452): particles released outside the main
// boundary, inside an island, or in no rho / u / v element.  ErrorFlag 2 -> die, 1 or 3 -> setOut,
// anything else is the reference's STOP (lowest id reported through d_bad).
__global__ void k_screen(const LtDev D, unsigned long long* __restrict__ cnt)
{
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= D.n) return;
    double X = D.x[n], Y = D.y[n];
    int code = 0;
    {                                                                    // mbounds :362
        int in = inpoly_banded(X, Y, D.bxy, D.mb_y0, D.mb_rbh, D.mb_n, D.mb_ptr, D.mb_idx);
        if (in == 2) in = inpoly(X, Y, D.maxbound, D.bxy, false) ? 1 : 0;
        if (!in) code = LTGPU_EV_INIT_OUT_MAIN;
    }
    if (!code && D.maxisland > 0) {                                      // ibounds :385
        int in = D.ib_ok ? inpoly_banded(X, Y, D.hxy, D.ib_y0, D.ib_rbh, D.ib_n, D.ib_ptr, D.ib_idx) : 2;
        if (in == 2) in = in_any_island(D, X, Y) ? 1 : 0;
        if (in) code = LTGPU_EV_INIT_IN_ISLAND;
    }
    if (!code) {                                                         // setEle_all :412-452
        if (D.r_ele[n] == 0) code = LTGPU_EV_INIT_NOT_IN_RHO;
        else if (D.u_ele[n] == 0) code = LTGPU_EV_INIT_NOT_IN_U;
        else if (D.v_ele[n] == 0) code = LTGPU_EV_INIT_NOT_IN_V;
    }
    if (!code) return;
    int EF = D.P.ErrorFlag;
    int gid = (int)(D.first_id + D.pid[n]);
    if (EF < 1 || EF > 3) atomicMin(D.bad, gid);
    else if (EF == 2) D.flags[n] |= LT_F_DEAD;
    else D.flags[n] |= LT_F_OOB;
    int k = atomicAdd(D.nev, 1);
    if (k < D.evcap) { D.ev[k].particle = gid; D.ev[k].code = code; D.ev[k].time = 0.0; }
    atomicAdd(cnt + (code - LTGPU_EV_INIT_OUT_MAIN), 1ull);
}

// x2lon / y2lat, double-precision branches (conversion_module.f90:322-378), per slot
__global__ void k_lonlat(const LtDev D, int spherical, double lonmin, double latmin, double R, double* __restrict__ lon, double* __restrict__ lat)
{
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= D.n) return;
    const double pi = D.P.PI, RCF = 180.0 / pi, x = D.x[n], y = D.y[n];
    if (spherical) {
        double la = y * 180.0 / (R * pi) + latmin;
        lon[n] = x * 180.0 / (R * pi * cos(la / RCF)) + lonmin;
        lat[n] = y * RCF / R + latmin;
    } else {
        lon[n] = x / R * RCF;
        lat[n] = 2.0 * RCF * (atan(exp(y / R)) - pi / 4.0);
    }
}

// FP64 FMA throughput of the device: the second roof of this path.  8 independent
// FMA chains per thread, 1024 resident threads per SM.
__global__ void __launch_bounds__(256) k_fp64_peak(double* __restrict__ out, int iters, double a, double b)
{
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

__global__ void k_fill_i32(int* p, int v, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// --------------------------------------------------------------- context -----
struct ltgpu_ctx {
    ltgpu_params prm;
    int device = 0;
    cudaStream_t compute = nullptr, copy = nullptr;
    cudaEvent_t slot_ready[4] = {nullptr, nullptr, nullptr, nullptr}, slot_free = nullptr, t0 = nullptr, t1 = nullptr;
    LtDev D;
    std::vector<void*> owned;                // every cudaMalloc, freed in destroy
    // grid sizes
    int rho_nodes = 0, u_nodes = 0, v_nodes = 0;
    size_t esz = 4;                          // bytes per field element on the device
    uint8_t *m_rho = nullptr, *m_u = nullptr, *m_v = nullptr;      // masks (always uploaded: used by the fill kernel)
    // hydro ring
    int npushed = 0, pending_slot = -1, spare = 3;
    void* h_stage[2] = {nullptr, nullptr}; void* d_stage[2] = {nullptr, nullptr}; size_t stage_bytes = 0; int stage_i = 0;
    cudaEvent_t stage_done[2] = {nullptr, nullptr};
    // events
    std::vector<ltgpu_event> host_events;
    int* d_bad = nullptr; int* d_nev = nullptr; int evcap = 1 << 20;
    long long ev_lost = 0, ev_lost_reported = 0;   // events dropped because the device log was full
    unsigned long long* d_stats = nullptr; int* d_status = nullptr;
    double last_ix3 = -1e300;
    long long launches = 0;
    std::string err;
    bool have_grid = false, have_bounds = false, have_particles = false, have_habitat = false;
    int nthreads_grid = 0;
    // re-sort state
    bool sort_on = true; int sort_mode = 32, sort_every = 1, vt_bins = 0, sort_mode_cfg = -1, vt_bins_cfg = -1;   // slot order / VTurb visiting order (resort); cfg -1: by particle count
    int* d_vorder = nullptr; int ebits = 24, bbits_slot = 8, bbits_vt = 8;   // key layout of k_sort_keys
    unsigned *d_key = nullptr, *d_key2 = nullptr; int *d_idx = nullptr, *d_perm = nullptr, *d_pid = nullptr;
    void* d_cub = nullptr; size_t cub_bytes = 0;
    double* spare8 = nullptr; double* out8 = nullptr;          // bounce buffers of ltgpu_fetch
    bool direct_pending = false;                               // copies straight into page-locked caller memory in flight
    double* alt8[11] = {}; int* alt4[8] = {}; uint8_t* alt1[2] = {};   // second copy of the per-slot state (re-sort target)
    void* h_out[2] = {nullptr, nullptr};     // pinned bounce buffers for fetch (pageable D2H is 4-5x slower)
    cudaEvent_t out_done[2] = {nullptr, nullptr}; int out_i = 0;
    void* pend_host[2] = {nullptr, nullptr}; size_t pend_bytes[2] = {0, 0};
    int key_bits = 32;
    long long sorts = 0;
    // VTurb scratch (one chunk of particles between k_vbuild and k_vwalk)
    int vt_chunk = 0; bool vt_legacy = false; int smem_per_sm = 228 * 1024; int vb_deep = -1;   // LTGPU_VB_DEEP: force the register variant
    // optional per-kernel timing (ltgpu_kernel_times)
    bool timing = false; cudaEvent_t tev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr}; float tacc[4] = {0, 0, 0, 0}; long long tcount = 0;
};

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_); return LTGPU_E_CUDA; } } while (0)
#define ARG(cond, msg) do { if (!(cond)) { ctx->err = msg; return LTGPU_E_ARG; } } while (0)

template <class T>
static int32_t dalloc(ltgpu_ctx* ctx, T** p, size_t count)
{
    void* q = nullptr;
    CK(cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T)));
    ctx->owned.push_back(q);
    *p = (T*)q;
    return LTGPU_OK;
}
template <class T>
static int32_t upload(ltgpu_ctx* ctx, const T** dst, const T* src, size_t count)
{
    T* q = nullptr;
    int32_t rc = dalloc(ctx, &q, count);
    if (rc) return rc;
    if (count) CK(cudaMemcpyAsync(q, src, count * sizeof(T), cudaMemcpyHostToDevice, ctx->compute));
    CK(cudaStreamSynchronize(ctx->compute));       // src may be a temporary
    *dst = q;
    return LTGPU_OK;
}
#define TRY(x) do { int32_t rc_ = (x); if (rc_) return rc_; } while (0)

// ----------------------------------------------------------- hydro refill ----
template <class TI>
static void fill_all(ltgpu_ctx* ctx, const uint8_t* dbase, const size_t off[7], bool have_st, int slot, cudaStream_t s)
{
    LtDev& D = ctx->D;
    int us = ctx->prm.us, ws = ctx->prm.ws;
    struct { int nodes, L; const uint8_t* mask; void* out; } f[7] = {
        {ctx->rho_nodes, 1, ctx->m_rho, (void*)D.zeta}, {ctx->u_nodes, us, ctx->m_u, (void*)D.u},
        {ctx->v_nodes, us, ctx->m_v, (void*)D.v}, {ctx->rho_nodes, ws, ctx->m_rho, (void*)D.w},
        {ctx->rho_nodes, ws, ctx->m_rho, (void*)D.kh}, {ctx->rho_nodes, us, ctx->m_rho, (void*)D.salt},
        {ctx->rho_nodes, us, ctx->m_rho, (void*)D.temp}};
    for (int i = 0; i < 7; ++i) {
        if (i >= 5 && !have_st) break;
        dim3 blk(32, 8), grd((f[i].nodes + 31) / 32, (f[i].L + 31) / 32);
        const TI* in = (const TI*)(dbase + off[i]);
        if (ctx->esz == 4) k_fill_slot<TI, float><<<grd, blk, 0, s>>>(in, f[i].mask, (float*)f[i].out, f[i].nodes, f[i].L, slot);
        else k_fill_slot<TI, double><<<grd, blk, 0, s>>>(in, f[i].mask, (double*)f[i].out, f[i].nodes, f[i].L, slot);
        ctx->launches++;
    }
}

// ---- table of root(x)/x for the SIGS convexity equation (lt_device.cuh::sig_guess) -----
static double convexity_G(double s)
{   // s (cosh s - 1) / (sinh s - s), accurate for all s > 0
    if (s < 0.3) {  // series: numerator s^3/2 (1 + s^2/12 + ...), denominator s^3/6 (1 + s^2/20 + ...)
        double s2 = s * s, num = 0.0, den = 0.0, tn = 0.5, td = 1.0 / 6.0;
        for (int k = 0; k < 12; ++k) { num += tn; den += td; tn *= s2 / ((2 * k + 3) * (2 * k + 4)); td *= s2 / ((2 * k + 4) * (2 * k + 5)); }
        return num / den;
    }
    double sh2 = sinh(0.5 * s);
    return s * (2.0 * sh2 * sh2) / (sinh(s) - s);
}
static void build_sigtab(std::vector<double>& tab)
{
    tab.assign(LT_SIGTAB_N + 1, 1.0);
    for (int i = 1; i <= LT_SIGTAB_N; ++i) {
        double x = (double)i / LT_SIGTAB_INV_H, target = 3.0 + x * x / 10.0;      // T + 1
        double lo = 0.0, hi = target + 1.0;                                        // G(s) >= s - ... and G(0+) = 3
        for (int it = 0; it < 200; ++it) { double mid = 0.5 * (lo + hi); if (convexity_G(mid) < target) lo = mid; else hi = mid; }
        tab[i] = 0.5 * (lo + hi) / x;
    }
}

// ---- exact spatial indices over the boundary tables (used by k_finish) ----------------
// The bucket / band of a coordinate is floor((v - origin) * scale) in IEEE double on both
// host and device; that map is monotone, which is all the exactness argument needs.
static void csr_from_lists(const std::vector<std::vector<int>>& lists, std::vector<int>& ptr, std::vector<int>& idx)
{
    ptr.assign(lists.size() + 1, 0);
    for (size_t i = 0; i < lists.size(); ++i) ptr[i + 1] = ptr[i] + (int)lists[i].size();
    idx.clear(); idx.reserve(ptr.back());
    for (auto& l : lists) idx.insert(idx.end(), l.begin(), l.end());
}
static void band_index(const std::vector<double2>& poly, const std::vector<char>& edge_ok, double& y0, double& rbh, int& nband,
                       std::vector<int>& ptr, std::vector<int>& idx)
{
    int n = (int)poly.size();
    double ymin = 1e300, ymax = -1e300;
    for (auto& q : poly) { ymin = std::min(ymin, q.y); ymax = std::max(ymax, q.y); }
    double ext = std::max(ymax - ymin, 1e-9);
    nband = std::max(1, std::min(8192, n / 2));
    y0 = ymin - 1e-6 * ext;
    rbh = (double)nband / (ext * (1.0 + 2e-6));
    std::vector<std::vector<int>> lists(nband);
    for (int i = 0; i + 1 < n; ++i) {
        if (!edge_ok[i]) continue;
        double lo = std::min(poly[i].y, poly[i + 1].y), hi = std::max(poly[i].y, poly[i + 1].y);
        int b0 = std::max(0, std::min(nband - 1, (int)floor((lo - y0) * rbh)));
        int b1 = std::max(0, std::min(nband - 1, (int)floor((hi - y0) * rbh)));
        for (int bq = b0; bq <= b1; ++bq) lists[bq].push_back(i);
    }
    csr_from_lists(lists, ptr, idx);
}
static bool host_inpoly(double x, double y, const double2* e, int n)
{   // plain crossing test, only used to validate that islands are not nested
    bool in = false;
    for (int i = 0, j = n - 1; i < n; j = i++)
        if (((e[i].y > y) != (e[j].y > y)) && (x < (e[j].x - e[i].x) * (y - e[i].y) / (e[j].y - e[i].y) + e[i].x)) in = !in;
    return in;
}
static int32_t build_indices(ltgpu_ctx* ctx, const std::vector<double4>& seg, const std::vector<double2>& b,
                             const std::vector<double2>& h, const int32_t* hid)
{
    LtDev& D = ctx->D;
    int ns = (int)seg.size();
    // segment buckets
    {
        double xmin = 1e300, xmax = -1e300, ymin = 1e300, ymax = -1e300;
        std::vector<double> len(ns);
        for (int i = 0; i < ns; ++i) {
            xmin = std::min({xmin, seg[i].x, seg[i].z}); xmax = std::max({xmax, seg[i].x, seg[i].z});
            ymin = std::min({ymin, seg[i].y, seg[i].w}); ymax = std::max({ymax, seg[i].y, seg[i].w});
            len[i] = hypot(seg[i].z - seg[i].x, seg[i].w - seg[i].y);
        }
        double cs = 1.0;
        if (ns) { std::vector<double> t = len; std::nth_element(t.begin(), t.begin() + ns / 2, t.end()); cs = std::max(2.0 * t[ns / 2], 1e-9); }
        else { xmin = ymin = 0; xmax = ymax = 1; }
        while ((xmax - xmin) / cs * ((ymax - ymin) / cs) > 4.0e6) cs *= 1.5;
        D.sg_x0 = xmin - cs; D.sg_y0 = ymin - cs; D.sg_rcs = 1.0 / cs;
        D.sg_nx = (int)floor((xmax - D.sg_x0) * D.sg_rcs) + 2; D.sg_ny = (int)floor((ymax - D.sg_y0) * D.sg_rcs) + 2;
        std::vector<std::vector<int>> lists((size_t)D.sg_nx * D.sg_ny);
        for (int i = 0; i < ns; ++i) {
            int cx0 = (int)floor((std::min(seg[i].x, seg[i].z) - D.sg_x0) * D.sg_rcs), cx1 = (int)floor((std::max(seg[i].x, seg[i].z) - D.sg_x0) * D.sg_rcs);
            int cy0 = (int)floor((std::min(seg[i].y, seg[i].w) - D.sg_y0) * D.sg_rcs), cy1 = (int)floor((std::max(seg[i].y, seg[i].w) - D.sg_y0) * D.sg_rcs);
            for (int cy = cy0; cy <= cy1; ++cy) for (int cx = cx0; cx <= cx1; ++cx) lists[(size_t)cy * D.sg_nx + cx].push_back(i);
        }
        std::vector<int> ptr, idx; csr_from_lists(lists, ptr, idx);
        TRY(upload(ctx, &D.sg_ptr, ptr.data(), ptr.size()));
        TRY(upload(ctx, &D.sg_idx, idx.data(), idx.size()));
    }
    // main polygon bands
    {
        std::vector<char> ok(b.size(), 1);
        std::vector<int> ptr, idx;
        if (b.size() >= 2) band_index(b, ok, D.mb_y0, D.mb_rbh, D.mb_n, ptr, idx);
        else { D.mb_n = 0; D.mb_y0 = 0; D.mb_rbh = 0; ptr.assign(1, 0); }
        TRY(upload(ctx, &D.mb_ptr, ptr.data(), ptr.size()));
        TRY(upload(ctx, &D.mb_idx, idx.data(), idx.size()));
    }
    // island bands: one joint crossing count over all islands is "in any island" only if the
    // islands are disjoint and not nested; otherwise the per-island routine is used (ib_ok = 0)
    {
        int nh = (int)h.size();
        std::vector<char> ok(nh, 1);
        std::vector<int> start;
        for (int i = 0; i < nh; ++i) { if (i == 0 || hid[i] != hid[i - 1]) start.push_back(i); if (i + 1 < nh && hid[i + 1] != hid[i]) ok[i] = 0; }
        start.push_back(nh);
        D.ib_ok = 1;
        for (size_t a = 0; a + 1 < start.size() && D.ib_ok; ++a)
            for (size_t c = 0; c + 1 < start.size(); ++c)
                if (a != c && host_inpoly(h[start[a]].x, h[start[a]].y, h.data() + start[c], start[c + 1] - start[c] - 1)) { D.ib_ok = 0; break; }
        std::vector<int> ptr, idx;
        if (nh >= 2) band_index(h, ok, D.ib_y0, D.ib_rbh, D.ib_n, ptr, idx);
        else { D.ib_n = 0; D.ib_y0 = 0; D.ib_rbh = 0; ptr.assign(1, 0); }
        TRY(upload(ctx, &D.ib_ptr, ptr.data(), ptr.size()));
        TRY(upload(ctx, &D.ib_idx, idx.data(), idx.size()));
    }
    return LTGPU_OK;
}

// ---------------------------------------------------------------- re-sort ----
// all per-slot state in one pass: slot i takes the state of old slot perm[i]
#define LT_NP8 11
#define LT_NP4 8
#define LT_NP1 2
struct PermArgs { const double* s8[LT_NP8]; double* d8[LT_NP8]; const int* s4[LT_NP4]; int* d4[LT_NP4]; const uint8_t* s1[LT_NP1]; uint8_t* d1[LT_NP1]; };
__global__ void k_permute_all(const __grid_constant__ PermArgs a, const int* __restrict__ perm, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int j = perm[i];
#pragma unroll
    for (int k = 0; k < LT_NP8; ++k) a.d8[k][i] = a.s8[k][j];
#pragma unroll
    for (int k = 0; k < LT_NP4; ++k) a.d4[k][i] = a.s4[k][j];
#pragma unroll
    for (int k = 0; k < LT_NP1; ++k) a.d1[k][i] = a.s1[k][j];
}
static int32_t resort(ltgpu_ctx* ctx)
{
    LtDev& D = ctx->D; int n = D.n;
    k_sort_keys<<<(n + 255) / 256, 256, 0, ctx->compute>>>(D, ctx->d_key, ctx->d_idx, ctx->sort_mode, ctx->ebits);
    CK(cub::DeviceRadixSort::SortPairs(ctx->d_cub, ctx->cub_bytes, ctx->d_key, ctx->d_key2, ctx->d_idx, ctx->d_perm, n, 0, ctx->sort_mode > 0 ? ctx->ebits + ctx->bbits_slot : 32, ctx->compute));
    double** a8[LT_NP8] = {&D.x, &D.y, &D.z, &D.age, &D.dob, &D.lifespan, &D.psalt, &D.ptemp, &D.timer, &D.sprev, &D.zprev};
    int** a4[LT_NP4] = {&D.r_ele, &D.u_ele, &D.v_ele, &D.hitB, &D.hitL, &D.endpoly, &D.nsig, &ctx->d_pid};
    uint8_t* beh = (uint8_t*)D.behave;
    uint8_t** a1[LT_NP1] = {&D.flags, &beh};
    PermArgs pa;
    for (int k = 0; k < LT_NP8; ++k) { pa.s8[k] = *a8[k]; pa.d8[k] = ctx->alt8[k]; }
    for (int k = 0; k < LT_NP4; ++k) { pa.s4[k] = *a4[k]; pa.d4[k] = ctx->alt4[k]; }
    for (int k = 0; k < LT_NP1; ++k) { pa.s1[k] = *a1[k]; pa.d1[k] = ctx->alt1[k]; }
    k_permute_all<<<(n + 255) / 256, 256, 0, ctx->compute>>>(pa, ctx->d_perm, n);
    for (int k = 0; k < LT_NP8; ++k) std::swap(*a8[k], ctx->alt8[k]);
    for (int k = 0; k < LT_NP4; ++k) std::swap(*a4[k], ctx->alt4[k]);
    for (int k = 0; k < LT_NP1; ++k) std::swap(*a1[k], ctx->alt1[k]);
    D.behave = (int8_t*)beh;
    D.pid = ctx->d_pid;
    ctx->launches += 3;
    // The kernels want different neighbours.  k_advect / k_finish gain from lanes in the same or adjacent
    // ELEMENTS (field windows and element records shared through L1 / L2): few depth bins.  k_vwalk gains from
    // lanes at the same RELATIVE DEPTH (spline interval changes, tension regimes, boundary clamps in step):
    // many bins; k_vbuild fits one column per warp and follows either.  So the slots are ordered by
    // (sort_mode bins, element) and the VTurb kernels visit them through a second index ordered by
    // (vt_bins, element).  Measured at 12.5 M particles with one order for all: 8 bins k_advect 26.2 ms,
    // VTurb 52.6; 32 bins 28.6 / 50.5; 64 bins 30.8 / 49.9.
    if (ctx->d_vorder && ctx->prm.VTurbOn && !ctx->vt_legacy) {
        k_sort_keys<<<(n + 255) / 256, 256, 0, ctx->compute>>>(D, ctx->d_key, ctx->d_idx, ctx->vt_bins, ctx->ebits);
        // the slots were just ordered by (few bins, element) and the radix sort is stable: one pass over the bin
        // byte alone leaves them ordered by (vt_bins, element)
        CK(cub::DeviceRadixSort::SortPairs(ctx->d_cub, ctx->cub_bytes, ctx->d_key, ctx->d_key2, ctx->d_idx, ctx->d_vorder, n, ctx->sort_mode > 0 ? ctx->ebits : 0, ctx->ebits + ctx->bbits_vt, ctx->compute));
        D.vorder = ctx->d_vorder;
        // k_vbuild fits one column per warp: lane coherence means nothing to it, but consecutive slots in the same
        // element re-use the KH columns through L1 / L2.  It keeps the slot order (scratch column = slot) when one
        // chunk holds all particles; only k_vwalk follows vorder (VTurb 50.2 -> 48.9 ms at 12.5 M particles).
        { const char* e = getenv("LTGPU_VB_SLOT_ORDER"); D.vb_slot_order = (!(e && e[0] == '0') && ctx->vt_chunk >= n) ? 1 : 0; }
        ctx->launches += 2;
    }
    CK(cudaGetLastError());
    ctx->sorts++;
    return LTGPU_OK;
}
// copy one per-slot column to the host in particle order: scatter on the device, D2H into a
// pinned bounce buffer (double buffered so the next column's copy overlaps this memcpy)
template <class V>
static int32_t fetch_col(ltgpu_ctx* ctx, V* host, const V* dev)
{
    if (!host) return LTGPU_OK;
    int n = ctx->D.n;
    size_t bytes = sizeof(V) * (size_t)n;
    int b = ctx->out_i; ctx->out_i ^= 1;
    V* tmp = (V*)(b ? (void*)ctx->spare8 : (void*)ctx->out8);       // spare8 is free between re-sorts
    k_scatter<V><<<(n + 255) / 256, 256, 0, ctx->compute>>>(dev, tmp, ctx->d_pid, n);
    ctx->launches++;
    {   // a page-locked destination (cudaHostAlloc / cudaHostRegister by the caller) takes the copy
        // directly: no bounce buffer, no host memcpy; fetch_flush waits for the stream
        cudaPointerAttributes pa;
        if (cudaPointerGetAttributes(&pa, host) == cudaSuccess && pa.type == cudaMemoryTypeHost) {
            CK(cudaMemcpyAsync(host, tmp, bytes, cudaMemcpyDeviceToHost, ctx->compute));
            ctx->direct_pending = true;
            return LTGPU_OK;
        }
        cudaGetLastError();
    }
    CK(cudaMemcpyAsync(ctx->h_out[b], tmp, bytes, cudaMemcpyDeviceToHost, ctx->compute));
    CK(cudaEventRecord(ctx->out_done[b], ctx->compute));
    ctx->pend_host[b] = host; ctx->pend_bytes[b] = bytes;
    // finish the PREVIOUS column while this one is in flight
    int pb = b ^ 1;
    if (ctx->pend_host[pb]) {
        CK(cudaEventSynchronize(ctx->out_done[pb]));
        memcpy(ctx->pend_host[pb], ctx->h_out[pb], ctx->pend_bytes[pb]);
        ctx->pend_host[pb] = nullptr;
    }
    return LTGPU_OK;
}
static int32_t fetch_flush(ltgpu_ctx* ctx)
{
    if (ctx->direct_pending) { CK(cudaStreamSynchronize(ctx->compute)); ctx->direct_pending = false; }
    for (int b = 0; b < 2; ++b)
        if (ctx->pend_host[b]) {
            CK(cudaEventSynchronize(ctx->out_done[b]));
            memcpy(ctx->pend_host[b], ctx->h_out[b], ctx->pend_bytes[b]);
            ctx->pend_host[b] = nullptr;
        }
    return LTGPU_OK;
}

// The launches of one internal step on the compute stream; T = field storage type, PH = ring phase.
template <class T, int PH>
static int32_t launch_step(ltgpu_ctx* ctx)
{
    const LtDev& D = ctx->D;
    cudaStream_t st = ctx->compute;
    const bool vt = ctx->prm.VTurbOn != 0;
    if (ctx->timing) cudaEventRecord(ctx->tev[0], st);
    k_advect<T, PH><<<(D.n + LT_BLK_ADV - 1) / LT_BLK_ADV, LT_BLK_ADV, 0, st>>>(D);
    ctx->launches++;
    if (ctx->timing) cudaEventRecord(ctx->tev[1], st);
    if (vt && ctx->vt_legacy) {
        k_vturb<T, PH><<<(D.n + LT_BLK_VT - 1) / LT_BLK_VT, LT_BLK_VT, 0, st>>>(D);
        ctx->launches++;
    } else if (vt) {
        const size_t smem = sizeof(double) * (size_t)vb_smem_doubles(ctx->prm.ws) * (LT_BLK_VB / 32);
        const bool deep = ctx->vb_deep >= 0 ? ctx->vb_deep != 0 : (smem + sizeof(double) * 4 * LT_MAXLEV + 1024) * LT_MIN_VB > (size_t)ctx->smem_per_sm;
        auto kb = deep ? k_vbuild<T, PH, LT_MIN_VB - 1> : k_vbuild<T, PH, LT_MIN_VB>;
        CK(cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                 // per device and function
        for (int base = 0; base < D.n; base += ctx->vt_chunk) {
            const int count = std::min(ctx->vt_chunk, D.n - base);
            kb<<<(count + LT_BLK_VB - 1) / LT_BLK_VB, LT_BLK_VB, smem, st>>>(D, base, count);
            if (ctx->prm.vturb_fp32_walk) k_vwalk_f32<T, PH><<<(count + LT_BLK_VW - 1) / LT_BLK_VW, LT_BLK_VW, 0, st>>>(D, base, count);
            else k_vwalk<T, PH><<<(count + LT_BLK_VW - 1) / LT_BLK_VW, LT_BLK_VW, 0, st>>>(D, base, count);
            ctx->launches += 2;
        }
    }
    if (ctx->timing) cudaEventRecord(ctx->tev[2], st);
    k_finish<T, PH><<<(D.n + LT_BLK_FIN - 1) / LT_BLK_FIN, LT_BLK_FIN, 0, st>>>(D);
    ctx->launches++;
    if (ctx->timing) {
        cudaEventRecord(ctx->tev[3], st);
        cudaEventSynchronize(ctx->tev[3]);
        for (int k = 0; k < 3; ++k) { float ms = 0; cudaEventElapsedTime(&ms, ctx->tev[k], ctx->tev[k + 1]); ctx->tacc[k] += ms; }
        ctx->tcount++;
    }
    return LTGPU_OK;
}

// Move the device event log to the host and recycle it.  Called at the library's synchronisation
// points (ltgpu_sync, ltgpu_drain_events) with the compute stream idle, so the log only has to
// hold the events of the steps queued between two of them; what did not fit is counted in
// ev_lost and reported by ltgpu_drain_events / ltgpu_events_lost.
static int32_t pull_events(ltgpu_ctx* ctx)
{
    if (!ctx->d_nev) return LTGPU_OK;
    int nev = 0;
    CK(cudaMemcpyAsync(&nev, ctx->d_nev, 4, cudaMemcpyDeviceToHost, ctx->compute));
    CK(cudaStreamSynchronize(ctx->compute));
    if (nev <= 0) return LTGPU_OK;
    const int keep = std::min(nev, ctx->evcap);
    if (nev > ctx->evcap) ctx->ev_lost += (long long)nev - ctx->evcap;
    const size_t old = ctx->host_events.size();
    ctx->host_events.resize(old + keep);
    CK(cudaMemcpyAsync(ctx->host_events.data() + old, ctx->D.ev, sizeof(ltgpu_event) * keep, cudaMemcpyDeviceToHost, ctx->compute));
    CK(cudaMemsetAsync(ctx->d_nev, 0, 4, ctx->compute));
    CK(cudaStreamSynchronize(ctx->compute));
    return LTGPU_OK;
}

extern "C" {

int32_t ltgpu_create(const ltgpu_params* prm, int32_t device, ltgpu_ctx** out)
{
    if (!prm || !out) return LTGPU_E_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev)
        return LTGPU_E_NODEVICE;                   // no CPU fallback
    ltgpu_ctx* ctx = new ltgpu_ctx();
    ctx->prm = *prm; ctx->device = device;
    memset(&ctx->D, 0, sizeof(LtDev));
    ctx->D.P = *prm;
    if (prm->us < 5 || prm->us >= LT_MAXLEV || prm->ws != prm->us + 1 || prm->idt <= 0 || prm->dt < prm->idt ||
        prm->z0 == 0.0 || prm->Vtransform < 1 || prm->Vtransform > 3 ||
        (prm->field_dtype != LTGPU_F32 && prm->field_dtype != LTGPU_F64) || prm->rng_mode != LTGPU_RNG_PHILOX) {
        delete ctx; return LTGPU_E_ARG;            // z0 == 0 is `stop 'dividing by 0'` LTRANS.f90:1512
    }
    ctx->esz = prm->field_dtype == LTGPU_F32 ? 4 : 8;
    if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return LTGPU_E_NODEVICE; }
    cudaDeviceGetAttribute(&ctx->smem_per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, device);
    cudaDeviceProp pr;
    if (cudaGetDeviceProperties(&pr, device) != cudaSuccess) { delete ctx; return LTGPU_E_NODEVICE; }
    ctx->nthreads_grid = pr.multiProcessorCount * 1024;
    if (cudaStreamCreateWithFlags(&ctx->compute, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->copy, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return LTGPU_E_CUDA; }
    {
        static std::vector<double> tab;
        if (tab.empty()) build_sigtab(tab);
        if (cudaMemcpyToSymbol(g_sigtab, tab.data(), sizeof(double) * tab.size()) != cudaSuccess) { delete ctx; return LTGPU_E_CUDA; }
    }
    for (int i = 0; i < 4; ++i) cudaEventCreateWithFlags(&ctx->slot_ready[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->slot_free, cudaEventDisableTiming);
    for (int i = 0; i < 2; ++i) cudaEventCreateWithFlags(&ctx->stage_done[i], cudaEventDisableTiming);
    cudaEventCreate(&ctx->t0); cudaEventCreate(&ctx->t1);
    ctx->D.sb = 0; ctx->D.sc = 1; ctx->D.sf = 2; ctx->spare = 3;
    const char* so = getenv("LTGPU_SORT");
    ctx->sort_on = !(so && so[0] == '0');
    { const char* vd = getenv("LTGPU_VB_DEEP"); if (vd) ctx->vb_deep = atoi(vd); }
    { const char* se = getenv("LTGPU_SORT_EVERY"); if (se && atoi(se) > 0) ctx->sort_every = atoi(se); }
    // -1: chosen by ltgpu_set_particles from the particle count
    { const char* sm = getenv("LTGPU_SORT_MODE"); ctx->sort_mode_cfg = sm ? std::max(0, std::min(127, atoi(sm))) : -1; }
    { const char* vb = getenv("LTGPU_VT_BINS"); ctx->vt_bins_cfg = vb ? std::max(0, std::min(127, atoi(vb))) : -1; }   // 0: VTurb in slot order
    { const char* vl = getenv("LTGPU_VTURB_LEGACY"); ctx->vt_legacy = vl && vl[0] == '1'; }     // round-1 fused k_vturb (A/B only)
    { const char* ec = getenv("LTGPU_EVCAP"); ctx->evcap = ec && atoi(ec) > 0 ? atoi(ec) : 0; }   // 0: sized by set_particles
    *out = ctx;
    return LTGPU_OK;
}

int32_t ltgpu_destroy(ltgpu_ctx* ctx)
{
    if (!ctx) return LTGPU_OK;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (void* p : ctx->owned) cudaFree(p);
    for (int i = 0; i < 2; ++i) { if (ctx->h_out[i]) cudaFreeHost(ctx->h_out[i]); if (ctx->out_done[i]) cudaEventDestroy(ctx->out_done[i]); }
    for (int i = 0; i < 2; ++i) { if (ctx->h_stage[i]) cudaFreeHost(ctx->h_stage[i]); if (ctx->stage_done[i]) cudaEventDestroy(ctx->stage_done[i]); }
    for (int i = 0; i < 4; ++i) if (ctx->slot_ready[i]) cudaEventDestroy(ctx->slot_ready[i]);
    if (ctx->slot_free) cudaEventDestroy(ctx->slot_free);
    if (ctx->t0) cudaEventDestroy(ctx->t0);
    if (ctx->t1) cudaEventDestroy(ctx->t1);
    for (int k = 0; k < 5; ++k) if (ctx->tev[k]) cudaEventDestroy(ctx->tev[k]);
    if (ctx->compute) cudaStreamDestroy(ctx->compute);
    if (ctx->copy) cudaStreamDestroy(ctx->copy);
    delete ctx;
    return LTGPU_OK;
}

const char* ltgpu_last_error(const ltgpu_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

static int32_t make_gridtab(ltgpu_ctx* ctx, LtGridTab* G, int nE, int nodes, const int32_t* E, const int32_t* Adj,
                            const double* nx, const double* ny, const uint8_t* dmask)
{
    std::vector<double> ele((size_t)nE * 8);
    std::vector<int4> nd(nE);
    std::vector<int> adj((size_t)nE * 10);
    for (int e = 0; e < nE; ++e) {
        int q[4];
        for (int i = 0; i < 4; ++i) {
            q[i] = E[4 * (size_t)e + i] - 1;
            if (q[i] < 0 || q[i] >= nodes) { ctx->err = "element node id out of range"; return LTGPU_E_ARG; }
            ele[(size_t)e * 8 + i] = nx[q[i]]; ele[(size_t)e * 8 + 4 + i] = ny[q[i]];     // hydro:561-580
        }
        nd[e] = make_int4(q[0], q[1], q[2], q[3]);
        for (int i = 0; i < 10; ++i) {
            int a = Adj[(size_t)i * nE + e];                       // Fortran (nE,10) column-major
            if (a < 0 || a > nE) { ctx->err = "adjacency id out of range"; return LTGPU_E_ARG; }
            adj[(size_t)e * 10 + i] = a;
        }
    }
    TRY(upload(ctx, &G->ele, ele.data(), ele.size()));
    TRY(upload(ctx, &G->node, nd.data(), nd.size()));
    TRY(upload(ctx, &G->adj, adj.data(), adj.size()));
    G->mask = dmask; G->nE = nE; G->nodes = nodes;
    {   // gridcell_quick is only valid for convex, non-degenerate quads: all four corner turns must
        // have the sign of the area, each by a clear margin
        int convex = 1;
        for (int e = 0; e < nE && convex; ++e) {
            const double* q = &ele[(size_t)e * 8];
            double A = (q[2] - q[0]) * (q[7] - q[5]) - (q[3] - q[1]) * (q[6] - q[4]);
            if (!(fabs(A) > 0.0)) { convex = 0; break; }
            for (int i = 0; i < 4; ++i) {
                int a = i, b = (i + 1) & 3, c = (i + 2) & 3;
                double turn = (q[b] - q[a]) * (q[4 + c] - q[4 + b]) - (q[4 + b] - q[4 + a]) * (q[c] - q[b]);
                if (!(turn * (A < 0 ? -1.0 : 1.0) > 1e-6 * fabs(A))) { convex = 0; break; }
            }
        }
        G->convex = convex;
    }
    {   // bucket index for k_locate: about one element per bucket
        double xmin = 1e300, xmax = -1e300, ymin = 1e300, ymax = -1e300;
        for (int e = 0; e < nE; ++e)
            for (int i = 0; i < 4; ++i) {
                xmin = std::min(xmin, ele[(size_t)e * 8 + i]); xmax = std::max(xmax, ele[(size_t)e * 8 + i]);
                ymin = std::min(ymin, ele[(size_t)e * 8 + 4 + i]); ymax = std::max(ymax, ele[(size_t)e * 8 + 4 + i]);
            }
        double cs = std::max(sqrt(std::max((xmax - xmin) * (ymax - ymin), 1e-18) / (double)nE), 1e-9);
        while ((xmax - xmin) / cs * ((ymax - ymin) / cs) > 1.6e7) cs *= 1.5;
        G->lx0 = xmin - cs; G->ly0 = ymin - cs; G->lrcs = 1.0 / cs;
        G->lnx = (int)floor((xmax - G->lx0) * G->lrcs) + 2; G->lny = (int)floor((ymax - G->ly0) * G->lrcs) + 2;
        std::vector<int> ptr((size_t)G->lnx * G->lny + 1, 0), idx;
        auto range = [&](int e, int& cx0, int& cx1, int& cy0, int& cy1) {
            const double* q = &ele[(size_t)e * 8];
            double x0 = std::min({q[0], q[1], q[2], q[3]}), x1 = std::max({q[0], q[1], q[2], q[3]});
            double y0 = std::min({q[4], q[5], q[6], q[7]}), y1 = std::max({q[4], q[5], q[6], q[7]});
            cx0 = (int)floor((x0 - G->lx0) * G->lrcs); cx1 = (int)floor((x1 - G->lx0) * G->lrcs);
            cy0 = (int)floor((y0 - G->ly0) * G->lrcs); cy1 = (int)floor((y1 - G->ly0) * G->lrcs);
        };
        for (int e = 0; e < nE; ++e) {
            int cx0, cx1, cy0, cy1; range(e, cx0, cx1, cy0, cy1);
            for (int cy = cy0; cy <= cy1; ++cy) for (int cx = cx0; cx <= cx1; ++cx) ptr[(size_t)cy * G->lnx + cx + 1]++;
        }
        for (size_t c = 0; c + 1 < ptr.size(); ++c) ptr[c + 1] += ptr[c];
        idx.resize(ptr.back());
        std::vector<int> fill(ptr.begin(), ptr.end() - 1);
        for (int e = 0; e < nE; ++e) {                     // ascending e: every bucket list ascends
            int cx0, cx1, cy0, cy1; range(e, cx0, cx1, cy0, cy1);
            for (int cy = cy0; cy <= cy1; ++cy) for (int cx = cx0; cx <= cx1; ++cx) idx[fill[(size_t)cy * G->lnx + cx]++] = e;
        }
        if (idx.empty()) idx.push_back(0);
        TRY(upload(ctx, &G->lptr, ptr.data(), ptr.size()));
        TRY(upload(ctx, &G->lidx, idx.data(), idx.size()));
    }
    return LTGPU_OK;
}

static int32_t upload_mask(ltgpu_ctx* ctx, uint8_t** d, const int32_t* m, int n)
{
    std::vector<uint8_t> t(n);
    for (int i = 0; i < n; ++i) t[i] = m[i] != 0;
    const uint8_t* q = nullptr;
    TRY(upload(ctx, &q, t.data(), (size_t)n));
    *d = (uint8_t*)q;
    return LTGPU_OK;
}

int32_t ltgpu_set_grid(ltgpu_ctx* ctx, int32_t vi, int32_t uj, int32_t ui, int32_t vj,
    const double* rx, const double* ry, const double* ux, const double* uy,
    const double* vx, const double* vy, const double* depth, const double* angle,
    const int32_t* rho_mask, const int32_t* u_mask, const int32_t* v_mask,
    const double* SC, const double* CS, const double* SCW, const double* CSW,
    const int32_t* RE, const int32_t* UE, const int32_t* VE,
    int32_t nRE, int32_t nUE, int32_t nVE,
    const int32_t* rAdj, const int32_t* uAdj, const int32_t* vAdj)
{
    if (!ctx) return LTGPU_E_ARG;
    ARG(!ctx->have_grid, "set_grid called twice");
    ARG(rx && ry && ux && uy && vx && vy && depth && angle && rho_mask && u_mask && v_mask && SC && CS && SCW && CSW &&
        RE && UE && VE && rAdj && uAdj && vAdj, "set_grid: null array");
    ARG(vi > 1 && uj > 1 && ui > 0 && vj > 0 && nRE > 0 && nUE > 0 && nVE > 0, "set_grid: bad sizes");
    CK(cudaSetDevice(ctx->device));
    ctx->rho_nodes = vi * uj; ctx->u_nodes = ui * uj; ctx->v_nodes = vi * vj;
    TRY(upload_mask(ctx, &ctx->m_rho, rho_mask, ctx->rho_nodes));
    TRY(upload_mask(ctx, &ctx->m_u, u_mask, ctx->u_nodes));
    TRY(upload_mask(ctx, &ctx->m_v, v_mask, ctx->v_nodes));
    LtDev& D = ctx->D;
    TRY(make_gridtab(ctx, &D.R, nRE, ctx->rho_nodes, RE, rAdj, rx, ry, ctx->m_rho));
    TRY(make_gridtab(ctx, &D.U, nUE, ctx->u_nodes, UE, uAdj, ux, uy, ctx->m_u));
    TRY(make_gridtab(ctx, &D.V, nVE, ctx->v_nodes, VE, vAdj, vx, vy, ctx->m_v));
    TRY(upload(ctx, &D.depth, depth, (size_t)ctx->rho_nodes));
    TRY(upload(ctx, &D.angle, angle, (size_t)ctx->rho_nodes));
    int us = ctx->prm.us, ws = ctx->prm.ws;
    for (int k = 0; k < us; ++k) { D.SC[k] = SC[k]; D.CS[k] = CS[k]; }
    for (int k = 0; k < ws; ++k) { D.SCW[k] = SCW[k]; D.CSW[k] = CSW[k]; }
    // hydro ring: [node][level][4]
    size_t e = ctx->esz, rn = ctx->rho_nodes, un = ctx->u_nodes, vn = ctx->v_nodes;
    struct { const void** p; size_t count; } f[] = {
        {&D.zeta, rn * 4}, {&D.u, un * us * 4}, {&D.v, vn * us * 4}, {&D.w, rn * ws * 4}, {&D.kh, rn * ws * 4},
        {&D.salt, rn * us * 4}, {&D.temp, rn * us * 4}};
    for (auto& q : f) {
        uint8_t* p = nullptr;
        TRY(dalloc(ctx, &p, q.count * e));
        CK(cudaMemsetAsync(p, 0, q.count * e, ctx->compute));
        *q.p = p;
    }
    // staging: one record = zeta + u + v + w + aks + salt + temp, at the host dtype (<= 8 B)
    // (push_hydro pads every field to 16 bytes)
    {
        const size_t cnt[7] = {rn, un * us, vn * us, rn * ws, rn * ws, rn * us, rn * us};
        ctx->stage_bytes = 0;
        for (size_t c : cnt) ctx->stage_bytes += (c * 8 + 15) & ~(size_t)15;
    }
    for (int i = 0; i < 2; ++i) {
        CK(cudaMallocHost(&ctx->h_stage[i], ctx->stage_bytes));
        void* d = nullptr; CK(cudaMalloc(&d, ctx->stage_bytes)); ctx->owned.push_back(d); ctx->d_stage[i] = d;
    }
    CK(cudaStreamSynchronize(ctx->compute));
    ctx->have_grid = true;
    return LTGPU_OK;
}

int32_t ltgpu_set_bounds(ltgpu_ctx* ctx, int32_t nbounds, const double* bnd_x, const double* bnd_y, const int32_t* land,
    int32_t maxbound, const double* bx, const double* by, int32_t maxisland, const double* hx, const double* hy,
    const int32_t* hid)
{
    if (!ctx) return LTGPU_E_ARG;
    ARG(nbounds >= 0 && maxbound >= 0 && maxisland >= 0, "set_bounds: bad sizes");
    CK(cudaSetDevice(ctx->device));
    LtDev& D = ctx->D;
    std::vector<double4> seg(nbounds); std::vector<uint8_t> ld(nbounds);
    for (int i = 0; i < nbounds; ++i) {
        seg[i] = make_double4(bnd_x[2 * i], bnd_y[2 * i], bnd_x[2 * i + 1], bnd_y[2 * i + 1]);
        ld[i] = land[i] != 0;
    }
    std::vector<double2> b(maxbound), h(maxisland);
    for (int i = 0; i < maxbound; ++i) b[i] = make_double2(bx[i], by[i]);
    for (int i = 0; i < maxisland; ++i) h[i] = make_double2(hx[i], hy[i]);
    TRY(upload(ctx, &D.seg, seg.data(), seg.size()));
    TRY(upload(ctx, &D.land, ld.data(), ld.size()));
    TRY(upload(ctx, &D.bxy, b.data(), b.size()));
    TRY(upload(ctx, &D.hxy, h.data(), h.size()));
    TRY(upload(ctx, &D.hid, hid, (size_t)maxisland));
    D.nbounds = nbounds; D.maxbound = maxbound; D.maxisland = maxisland;
    TRY(build_indices(ctx, seg, b, h, hid));
    ctx->have_bounds = true;
    return LTGPU_OK;
}

int32_t ltgpu_set_habitat(ltgpu_ctx* ctx, int32_t pedges, const double* polys, int32_t hedges, const double* holes,
    int32_t npoly, const int32_t* poly_id, const int32_t* poly_start, const int32_t* poly_size, const double* poly_maxdis,
    int32_t nhole, const int32_t* hole_id, const int32_t* hole_start, const int32_t* hole_size, const double* hole_maxdis,
    const int32_t* elepoly_ptr, const int32_t* elepoly_idx, const int32_t* polyhole_ptr, const int32_t* polyhole_idx)
{
    if (!ctx) return LTGPU_E_ARG;
    ARG(ctx->have_grid, "set_habitat before set_grid");
    (void)poly_id; (void)hole_id;
    CK(cudaSetDevice(ctx->device));
    LtDev& D = ctx->D;
    TRY(upload(ctx, &D.polys, polys, (size_t)pedges * 5));
    TRY(upload(ctx, &D.holes, holes, (size_t)hedges * 6));
    TRY(upload(ctx, &D.poly_start, poly_start, (size_t)npoly));
    TRY(upload(ctx, &D.poly_size, poly_size, (size_t)npoly));
    TRY(upload(ctx, &D.poly_maxdis, poly_maxdis, (size_t)npoly));
    TRY(upload(ctx, &D.hole_start, hole_start, (size_t)nhole));
    TRY(upload(ctx, &D.hole_size, hole_size, (size_t)nhole));
    TRY(upload(ctx, &D.hole_maxdis, hole_maxdis, (size_t)nhole));
    TRY(upload(ctx, &D.elepoly_ptr, elepoly_ptr, (size_t)D.R.nE + 1));
    TRY(upload(ctx, &D.elepoly_idx, elepoly_idx, (size_t)elepoly_ptr[D.R.nE]));
    TRY(upload(ctx, &D.polyhole_ptr, polyhole_ptr, (size_t)npoly + 1));
    TRY(upload(ctx, &D.polyhole_idx, polyhole_idx, (size_t)polyhole_ptr[npoly]));
    D.pedges = pedges; D.hedges = hedges; D.npoly = npoly; D.nhole = nhole;
    ctx->have_habitat = true;
    return LTGPU_OK;
}

int32_t ltgpu_set_particles(ltgpu_ctx* ctx, int32_t n, int64_t first_id,
    const double* x, const double* y, const double* z, const double* dob, const int32_t* startpoly,
    const int32_t* r_ele, const int32_t* u_ele, const int32_t* v_ele)
{
    if (!ctx) return LTGPU_E_ARG;
    ARG(ctx->have_grid, "set_particles before set_grid");
    ARG(!ctx->have_particles, "set_particles called twice");
    ARG(n > 0 && x && y && z && dob, "set_particles: null array");
    (void)startpoly;                               // never changes: stays with the host (LTRANS.f90:650)
    CK(cudaSetDevice(ctx->device));
    LtDev& D = ctx->D;
    D.n = n; D.first_id = first_id;
    size_t N = (size_t)n;
    double** dd[] = {&D.x, &D.y, &D.z, &D.age, &D.dob, &D.lifespan, &D.psalt, &D.ptemp, &D.timer, &D.sprev, &D.zprev};
    for (auto p : dd) { TRY(dalloc(ctx, p, N)); CK(cudaMemsetAsync(*p, 0, N * 8, ctx->compute)); }
    int** ii[] = {&D.r_ele, &D.u_ele, &D.v_ele, &D.hitB, &D.hitL, &D.endpoly, &D.nsig};
    for (auto p : ii) { TRY(dalloc(ctx, p, N)); CK(cudaMemsetAsync(*p, 0, N * 4, ctx->compute)); }
    TRY(dalloc(ctx, &D.flags, N)); TRY(dalloc(ctx, &D.behave, N));
    CK(cudaMemsetAsync(D.flags, ctx->prm.Behavior == 7 ? LT_F_BOTTOM : 0, N, ctx->compute));   // behavior:113
    CK(cudaMemsetAsync(D.behave, ctx->prm.Behavior, N, ctx->compute));                          // behavior:118
    CK(cudaMemcpyAsync(D.x, x, N * 8, cudaMemcpyHostToDevice, ctx->compute));
    CK(cudaMemcpyAsync(D.y, y, N * 8, cudaMemcpyHostToDevice, ctx->compute));
    CK(cudaMemcpyAsync(D.z, z, N * 8, cudaMemcpyHostToDevice, ctx->compute));
    CK(cudaMemcpyAsync(D.dob, dob, N * 8, cudaMemcpyHostToDevice, ctx->compute));
    if (r_ele && u_ele && v_ele) {
        CK(cudaMemcpyAsync(D.r_ele, r_ele, N * 4, cudaMemcpyHostToDevice, ctx->compute));
        CK(cudaMemcpyAsync(D.u_ele, u_ele, N * 4, cudaMemcpyHostToDevice, ctx->compute));
        CK(cudaMemcpyAsync(D.v_ele, v_ele, N * 4, cudaMemcpyHostToDevice, ctx->compute));
    } else {
        int blocks = (n + 127) / 128;
        k_locate<<<blocks, 128, 0, ctx->compute>>>(D, D.R, D.r_ele);
        k_locate<<<blocks, 128, 0, ctx->compute>>>(D, D.U, D.u_ele);
        k_locate<<<blocks, 128, 0, ctx->compute>>>(D, D.V, D.v_ele);
        ctx->launches += 3;
    }
    // a particle raises at most one event per internal step, so max(2^20, n) entries cannot
    // overflow between two synchronisation points that are at most one step apart
    if (ctx->evcap <= 0) ctx->evcap = std::max(1 << 20, n);
    TRY(dalloc(ctx, &D.ev, (size_t)ctx->evcap)); D.evcap = ctx->evcap;
    TRY(dalloc(ctx, &ctx->d_nev, 1)); TRY(dalloc(ctx, &ctx->d_bad, 1));
    D.nev = ctx->d_nev; D.bad = ctx->d_bad;
    CK(cudaMemsetAsync(ctx->d_nev, 0, 4, ctx->compute));
    k_fill_i32<<<1, 32, 0, ctx->compute>>>(ctx->d_bad, INT_MAX, 1);
    TRY(dalloc(ctx, &ctx->d_stats, 8)); TRY(dalloc(ctx, &ctx->d_status, N));
    double** sc[] = {&D.s_depth, &D.s_angle, &D.s_zeb, &D.s_zec, &D.s_zef, &D.s_pzb, &D.s_pzc, &D.s_pzf, &D.s_zpar,
                     &D.s_nx, &D.s_ny, &D.s_advz, &D.s_pu, &D.s_pv, &D.s_turbv};
    for (auto p : sc) TRY(dalloc(ctx, p, N));
    TRY(dalloc(ctx, &D.s_act, N));
    TRY(dalloc(ctx, &ctx->d_pid, N)); TRY(dalloc(ctx, &ctx->d_perm, N)); TRY(dalloc(ctx, &ctx->d_idx, N));
    TRY(dalloc(ctx, &ctx->d_key, N)); TRY(dalloc(ctx, &ctx->d_key2, N));
    TRY(dalloc(ctx, &ctx->spare8, N)); TRY(dalloc(ctx, &ctx->out8, N));
    for (auto& q : ctx->alt8) TRY(dalloc(ctx, &q, N));
    for (auto& q : ctx->alt4) TRY(dalloc(ctx, &q, N));
    for (auto& q : ctx->alt1) TRY(dalloc(ctx, &q, N));
    for (int b = 0; b < 2; ++b) {          // pinned bounce buffers of ltgpu_fetch (allocated here: cudaMallocHost is slow)
        CK(cudaMallocHost(&ctx->h_out[b], sizeof(double) * N));
        CK(cudaEventCreateWithFlags(&ctx->out_done[b], cudaEventDisableTiming));
    }
    if (ctx->prm.VTurbOn && !ctx->vt_legacy) {
        int chunk = 1 << 24;            // 788 B of scratch per particle; fewer, longer launches: 1 M 52.3, 2 M 51.2, 4 M 50.9, all 12.5 M 50.5 ms
        { const char* vc = getenv("LTGPU_VTURB_CHUNK"); if (vc && atoi(vc) > 0) chunk = atoi(vc); }
        chunk = std::min((n + 31) / 32 * 32, (chunk + 31) / 32 * 32);
        ctx->vt_chunk = chunk; D.vw_stride = chunk;
        TRY(dalloc(ctx, &D.vw, (size_t)3 * VW * chunk)); TRY(dalloc(ctx, &D.vz1, (size_t)chunk)); TRY(dalloc(ctx, &D.vzn, (size_t)chunk));
        TRY(dalloc(ctx, &D.vka, (size_t)chunk));
    }
    {   // Two orders (slots by 8 depth bins x element, VTurb visiting order by 64 bins x element) pay when the
        // fields do not fit L2 - k_advect's windows then come from DRAM unless neighbouring lanes share them - and
        // there are enough particles to carry the second radix sort (0.1 ms per step): +3.4 % at 12.5 M particles
        // on the Gulf grid (1.9 GB of u, v, w, AKs).  With L2-resident fields lanes at the same relative depth are
        // worth more to k_advect than lanes in the same element: -8 % at 10 M particles on 120x80x20, -1.5 % at
        // 1 M on 130x130x20.  There: one order, 32 bins.
        int l2 = 0; cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, ctx->device);
        const size_t field_bytes = (size_t)ctx->rho_nodes * (size_t)ctx->prm.ws * 4 * ctx->esz * 4;
        const bool big = n >= (1 << 20) && field_bytes > (size_t)l2;
        ctx->sort_mode = ctx->sort_mode_cfg >= 0 ? ctx->sort_mode_cfg : (big ? 8 : 32);
        ctx->vt_bins = ctx->vt_bins_cfg >= 0 ? ctx->vt_bins_cfg : (big ? 64 : 0);
        D.vorder = nullptr; ctx->d_vorder = nullptr;
        if (ctx->sort_on && ctx->prm.VTurbOn && !ctx->vt_legacy && ctx->vt_bins > 0 && ctx->vt_bins != ctx->sort_mode)
            TRY(dalloc(ctx, &ctx->d_vorder, N));
    }
    k_iota<<<(n + 255) / 256, 256, 0, ctx->compute>>>(ctx->d_pid, n);
    D.pid = ctx->d_pid;
    {
        auto nbits = [](long long v) { int b = 1; while ((1ll << b) <= v) ++b; return b; };   // bits that hold 0 .. v
        ctx->ebits = nbits(D.R.nE); ctx->bbits_slot = nbits(std::max(ctx->sort_mode, 1)); ctx->bbits_vt = nbits(std::max(ctx->vt_bins, 1));
        if (ctx->ebits + std::max(ctx->bbits_slot, ctx->bbits_vt) > 32) return LTGPU_E_ARG;      // > 2^24 elements with 127 bins
        ctx->key_bits = 32;
        cub::DeviceRadixSort::SortPairs(nullptr, ctx->cub_bytes, ctx->d_key, ctx->d_key2, ctx->d_idx, ctx->d_perm, n, 0, ctx->key_bits, ctx->compute);
        void* q = nullptr; CK(cudaMalloc(&q, std::max<size_t>(ctx->cub_bytes, 16))); ctx->owned.push_back(q); ctx->d_cub = q;
    }
    CK(cudaStreamSynchronize(ctx->compute));
    ctx->have_particles = true;
    return LTGPU_OK;
}

int32_t ltgpu_screen_initial(ltgpu_ctx* ctx, int64_t counts[5], int64_t* bad_particle)
{
    if (!ctx) return LTGPU_E_ARG;
    ARG(ctx->have_particles && ctx->have_bounds, "screen_initial before set_particles / set_bounds");
    CK(cudaSetDevice(ctx->device));
    LtDev& D = ctx->D;
    CK(cudaMemsetAsync(ctx->d_stats, 0, 8 * sizeof(unsigned long long), ctx->compute));
    k_screen<<<(D.n + 127) / 128, 128, 0, ctx->compute>>>(D, ctx->d_stats);
    ctx->launches += 1;
    unsigned long long c[8]; int bad = INT_MAX;
    CK(cudaMemcpyAsync(c, ctx->d_stats, sizeof c, cudaMemcpyDeviceToHost, ctx->compute));
    CK(cudaMemcpyAsync(&bad, ctx->d_bad, 4, cudaMemcpyDeviceToHost, ctx->compute));
    CK(cudaStreamSynchronize(ctx->compute));
    if (counts) for (int k = 0; k < 5; ++k) counts[k] = (int64_t)c[k];
    if (bad != INT_MAX) {
        if (bad_particle) *bad_particle = bad;
        ctx->err = "particle " + std::to_string(bad) + ": bad initial location and ErrorFlag outside 1..3 (the reference STOPs)";
        return LTGPU_E_PARTICLE;
    }
    return LTGPU_OK;
}

int32_t ltgpu_push_hydro(ltgpu_ctx* ctx, int32_t dtype, const void* zeta, const void* u, const void* v, const void* w,
                         const void* aks, const void* salt, const void* temp)
{
    if (!ctx) return LTGPU_E_ARG;
    ARG(ctx->have_grid, "push_hydro before set_grid");
    ARG(dtype == LTGPU_F32 || dtype == LTGPU_F64, "push_hydro: bad dtype");
    ARG(zeta && u && v && w && aks, "push_hydro: null array");
    bool need_st = ctx->prm.SaltTempOn || ctx->prm.Behavior == 4 || ctx->prm.Behavior == 5 || ctx->prm.Behavior == 7;
    bool have_st = salt && temp;
    ARG(have_st || !need_st, "push_hydro: salt/temp required by SaltTempOn / Behavior 4,5,7");
    ARG(ctx->pending_slot < 0, "push_hydro: previous record not rotated in yet");
    CK(cudaSetDevice(ctx->device));
    int slot = ctx->npushed < 3 ? ctx->npushed : ctx->spare;
    size_t e = dtype == LTGPU_F32 ? 4 : 8, us = ctx->prm.us, ws = ctx->prm.ws;
    size_t rn = ctx->rho_nodes, un = ctx->u_nodes, vn = ctx->v_nodes;
    size_t cnt[7] = {rn, un * us, vn * us, rn * ws, rn * ws, rn * us, rn * us}, off[7];
    const void* src[7] = {zeta, u, v, w, aks, salt, temp};
    int si = ctx->stage_i; ctx->stage_i ^= 1;
    CK(cudaEventSynchronize(ctx->stage_done[si]));             // staging buffer free again?
    uint8_t* hs = (uint8_t*)ctx->h_stage[si];
    size_t o = 0;
    for (int i = 0; i < 7; ++i) {
        off[i] = o;
        if (i >= 5 && !have_st) continue;
        ARG(o + cnt[i] * e <= ctx->stage_bytes, "push_hydro: record larger than the staging buffer");
        memcpy(hs + o, src[i], cnt[i] * e);                    // pageable -> pinned
        o += (cnt[i] * e + 15) & ~(size_t)15;
    }
    // the spare slot may still be read as "back" by steps queued before the last rotate
    CK(cudaStreamWaitEvent(ctx->copy, ctx->slot_free, 0));
    CK(cudaMemcpyAsync(ctx->d_stage[si], hs, o, cudaMemcpyHostToDevice, ctx->copy));
    if (dtype == LTGPU_F32) fill_all<float>(ctx, (const uint8_t*)ctx->d_stage[si], off, have_st, slot, ctx->copy);
    else fill_all<double>(ctx, (const uint8_t*)ctx->d_stage[si], off, have_st, slot, ctx->copy);
    CK(cudaEventRecord(ctx->stage_done[si], ctx->copy));
    CK(cudaEventRecord(ctx->slot_ready[slot], ctx->copy));
    CK(cudaGetLastError());
    if (ctx->npushed < 3) CK(cudaStreamWaitEvent(ctx->compute, ctx->slot_ready[slot], 0));
    else ctx->pending_slot = slot;
    ctx->npushed++;
    return LTGPU_OK;
}

int32_t ltgpu_rotate_hydro(ltgpu_ctx* ctx)
{
    if (!ctx) return LTGPU_E_ARG;
    ARG(ctx->pending_slot >= 0, "rotate_hydro: no pushed record pending");
    CK(cudaSetDevice(ctx->device));
    LtDev& D = ctx->D;
    int old_b = D.sb;
    D.sb = D.sc; D.sc = D.sf; D.sf = ctx->pending_slot;        // hydro:1080-1082
    ctx->spare = old_b; ctx->pending_slot = -1;
    CK(cudaStreamWaitEvent(ctx->compute, ctx->slot_ready[D.sf], 0));
    CK(cudaEventRecord(ctx->slot_free, ctx->compute));          // steps queued so far were the last readers of old_b
    return LTGPU_OK;
}

// ---------------------------------------------------------------- stepping ---
int32_t ltgpu_step(ltgpu_ctx* ctx, int32_t p, int32_t it)
{
    if (!ctx) return LTGPU_E_ARG;
    ARG(ctx->have_grid && ctx->have_bounds && ctx->have_particles, "step: grid / bounds / particles not set");
    ARG(ctx->npushed >= 3, "step: fewer than 3 hydro records pushed");
    ARG(!ctx->prm.settlementon || ctx->have_habitat, "step: settlementon without set_habitat");
    ARG(p >= 1 && it >= 1, "step: p and it are 1-based");
    CK(cudaSetDevice(ctx->device));
    LtDev& D = ctx->D;
    const int dt = ctx->prm.dt, idt = ctx->prm.idt;
    if (ctx->sort_on && (it - 1) % ctx->sort_every == 0) TRY(resort(ctx));
    D.p = p; D.it = it;
    D.ex[0] = (double)((p - 2) * dt); D.ex[1] = (double)((p - 1) * dt); D.ex[2] = (double)(p * dt);   // LTRANS.f90:568-571
    D.ix[0] = D.ex[1] + (double)((it - 2) * idt);                                                      // :588-590
    D.ix[1] = D.ex[1] + (double)((it - 1) * idt);
    D.ix[2] = D.ex[1] + (double)(it * idt);
    D.gstep = (unsigned)((p - 1) * (dt / idt) + it);
    {   // Lagrange weights of the 3-point time polynomial at ix(1..3); p == 1 uses (b,b,c) (LTRANS.f90:1534-1544)
        const double e0 = D.ex[0], e1 = D.ex[1], e2 = D.ex[2];
        for (int v = 0; v < 3; ++v) {
            double x = D.ix[v];
            double L0 = (x - e1) * (x - e2) / ((e0 - e1) * (e0 - e2)), L1 = (x - e0) * (x - e2) / ((e1 - e0) * (e1 - e2)),
                   L2 = (x - e0) * (x - e1) / ((e2 - e0) * (e2 - e1));
            if (v == 1) { D.LWz[0] = L0; D.LWz[1] = L1; D.LWz[2] = L2; }
            // lag() uses y_c + w0 (y_b - y_c) + w2 (y_f - y_c); the p == 1 triplet (b,b,c) is
            // b + L2 (c - b) = y_c + (1 - L2)(y_b - y_c) + 0 (y_f - y_c)
            if (p == 1) { D.LW[v][0] = 1.0 - L2; D.LW[v][1] = L2; D.LW[v][2] = 0.0; }
            else { D.LW[v][0] = L0; D.LW[v][1] = L1; D.LW[v][2] = L2; }
        }
        for (int t = 0; t < 3; ++t) D.LW4[t] = (D.LW[0][t] + 4.0 * D.LW[1][t] + D.LW[2][t]) / 6.0;
    }
    {
        int32_t rc = LTGPU_OK;
        if (ctx->esz == 4) {
            switch (D.sb) { case 0: rc = launch_step<float, 0>(ctx); break; case 1: rc = launch_step<float, 1>(ctx); break;
                            case 2: rc = launch_step<float, 2>(ctx); break; default: rc = launch_step<float, 3>(ctx); break; }
        } else {
            switch (D.sb) { case 0: rc = launch_step<double, 0>(ctx); break; case 1: rc = launch_step<double, 1>(ctx); break;
                            case 2: rc = launch_step<double, 2>(ctx); break; default: rc = launch_step<double, 3>(ctx); break; }
        }
        if (rc) return rc;
    }
    CK(cudaGetLastError());
    ctx->last_ix3 = D.ix[2];
    return LTGPU_OK;
}

int32_t ltgpu_run_external(ltgpu_ctx* ctx, int32_t p)
{
    if (!ctx) return LTGPU_E_ARG;
    int stepIT = ctx->prm.dt / ctx->prm.idt;                   // LTRANS.f90:554
    for (int it = 1; it <= stepIT; ++it) TRY(ltgpu_step(ctx, p, it));
    return LTGPU_OK;
}

int32_t ltgpu_sync(ltgpu_ctx* ctx, int32_t* bad_particle)
{
    if (!ctx) return LTGPU_E_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->copy));
    CK(cudaStreamSynchronize(ctx->compute));
    int bad = INT_MAX;
    if (ctx->d_bad) CK(cudaMemcpy(&bad, ctx->d_bad, 4, cudaMemcpyDeviceToHost));
    if (bad_particle) *bad_particle = bad == INT_MAX ? 0 : bad;
    if (bad != INT_MAX) { ctx->err = "a particle hit a STOP condition (ErrorFlag outside 1..3)"; return LTGPU_E_PARTICLE; }
    return pull_events(ctx);
}

int32_t ltgpu_fetch(ltgpu_ctx* ctx, double* x, double* y, double* z, double* age, int32_t* status,
    double* salt, double* temp, int32_t* hitBottom, int32_t* hitLand, int32_t* endpoly, double* lifespan,
    int32_t* r_ele, int32_t* u_ele, int32_t* v_ele)
{
    if (!ctx) return LTGPU_E_ARG;
    ARG(ctx->have_particles, "fetch before set_particles");
    CK(cudaSetDevice(ctx->device));
    LtDev& D = ctx->D; size_t N = D.n; cudaStream_t s = ctx->compute;
    if (status) { k_status<<<(D.n + 255) / 256, 256, 0, s>>>(D, ctx->d_status); ctx->launches++; }
    TRY(fetch_col(ctx, x, (const double*)D.x)); TRY(fetch_col(ctx, y, (const double*)D.y)); TRY(fetch_col(ctx, z, (const double*)D.z));
    TRY(fetch_col(ctx, age, (const double*)D.age)); TRY(fetch_col(ctx, status, (const int*)ctx->d_status));
    TRY(fetch_col(ctx, salt, (const double*)D.psalt)); TRY(fetch_col(ctx, temp, (const double*)D.ptemp));
    TRY(fetch_col(ctx, hitBottom, (const int*)D.hitB)); TRY(fetch_col(ctx, hitLand, (const int*)D.hitL));
    TRY(fetch_col(ctx, endpoly, (const int*)D.endpoly)); TRY(fetch_col(ctx, lifespan, (const double*)D.lifespan));
    TRY(fetch_col(ctx, r_ele, (const int*)D.r_ele)); TRY(fetch_col(ctx, u_ele, (const int*)D.u_ele)); TRY(fetch_col(ctx, v_ele, (const int*)D.v_ele));
    (void)N; (void)s;
    return fetch_flush(ctx);
}

int32_t ltgpu_fetch_lonlat(ltgpu_ctx* ctx, int32_t spherical, double lonmin, double latmin, double earth_radius, double* lon, double* lat)
{
    if (!ctx || !lon || !lat) return LTGPU_E_ARG;
    ARG(ctx->have_particles, "fetch_lonlat before set_particles");
    ARG(earth_radius > 0.0, "fetch_lonlat: bad earth radius");
    CK(cudaSetDevice(ctx->device));
    LtDev& D = ctx->D;
    // s_nx / s_ny are per-slot scratch of the step kernels, rewritten by k_advect before they are read
    k_lonlat<<<(D.n + 255) / 256, 256, 0, ctx->compute>>>(D, spherical, lonmin, latmin, earth_radius, D.s_nx, D.s_ny);
    ctx->launches++;
    TRY(fetch_col(ctx, lon, (const double*)D.s_nx)); TRY(fetch_col(ctx, lat, (const double*)D.s_ny));
    return fetch_flush(ctx);
}

int32_t ltgpu_fetch_sigerr(ltgpu_ctx* ctx, int32_t* count)
{
    if (!ctx || !count) return LTGPU_E_ARG;
    ARG(ctx->have_particles, "fetch_sigerr before set_particles");
    CK(cudaSetDevice(ctx->device));
    TRY(fetch_col(ctx, count, (const int*)ctx->D.nsig));
    return fetch_flush(ctx);
}

int32_t ltgpu_reset_hits(ltgpu_ctx* ctx)
{
    if (!ctx) return LTGPU_E_ARG;
    ARG(ctx->have_particles, "reset_hits before set_particles");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemsetAsync(ctx->D.hitB, 0, (size_t)ctx->D.n * 4, ctx->compute));     // LTRANS.f90:1662-1665
    CK(cudaMemsetAsync(ctx->D.hitL, 0, (size_t)ctx->D.n * 4, ctx->compute));
    return LTGPU_OK;
}

int32_t ltgpu_stats(ltgpu_ctx* ctx, int64_t counts[8])
{
    if (!ctx || !counts) return LTGPU_E_ARG;
    ARG(ctx->have_particles, "stats before set_particles");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemsetAsync(ctx->d_stats, 0, 64, ctx->compute));
    int blocks = std::min((ctx->D.n + 255) / 256, 1184);
    k_stats<<<blocks, 256, 0, ctx->compute>>>(ctx->D, ctx->last_ix3, ctx->d_stats);
    ctx->launches++;
    unsigned long long h[8]; int nev = 0;
    CK(cudaMemcpyAsync(h, ctx->d_stats, 64, cudaMemcpyDeviceToHost, ctx->compute));
    CK(cudaMemcpyAsync(&nev, ctx->d_nev, 4, cudaMemcpyDeviceToHost, ctx->compute));
    CK(cudaStreamSynchronize(ctx->compute));
    for (int k = 0; k < 8; ++k) counts[k] = (int64_t)h[k];
    counts[5] = (int64_t)std::min(nev, ctx->evcap) + (int64_t)ctx->host_events.size();
    return LTGPU_OK;
}

int32_t ltgpu_drain_events(ltgpu_ctx* ctx, ltgpu_event* buf, int32_t cap, int32_t* n)
{
    if (!ctx || !buf || !n || cap < 0) return LTGPU_E_ARG;
    ARG(ctx->have_particles, "drain_events before set_particles");
    CK(cudaSetDevice(ctx->device));
    TRY(pull_events(ctx));
    // ErrorLog.txt order of the serial loop: by time step, then ascending particle id
    std::sort(ctx->host_events.begin(), ctx->host_events.end(), [](const ltgpu_event& a, const ltgpu_event& b) {
        return a.time != b.time ? a.time < b.time : a.particle < b.particle; });
    int k = std::min<int>(cap, (int)ctx->host_events.size());
    memcpy(buf, ctx->host_events.data(), sizeof(ltgpu_event) * k);
    ctx->host_events.erase(ctx->host_events.begin(), ctx->host_events.begin() + k);
    *n = k;
    if (ctx->ev_lost > ctx->ev_lost_reported) {
        ctx->err = std::to_string(ctx->ev_lost - ctx->ev_lost_reported) + " per-particle events were dropped: the device log (" +
                   std::to_string(ctx->evcap) + " entries) filled up between two synchronisation points";
        ctx->ev_lost_reported = ctx->ev_lost;
        return LTGPU_W_EVENTS_LOST;
    }
    return LTGPU_OK;
}

int32_t ltgpu_events_lost(ltgpu_ctx* ctx, int64_t* lost)
{
    if (!ctx || !lost) return LTGPU_E_ARG;
    *lost = (int64_t)ctx->ev_lost;
    return LTGPU_OK;
}

int32_t ltgpu_device_ptr(ltgpu_ctx* ctx, int32_t which, void** dptr)
{
    if (!ctx || !dptr) return LTGPU_E_ARG;
    ARG(ctx->have_particles, "device_ptr before set_particles");
    CK(cudaSetDevice(ctx->device));
    LtDev& D = ctx->D;
    // slots are re-sorted by cell; hand out a particle-order copy (valid until the next call)
    int n = D.n;
    if (which >= 0 && which <= 3) {
        const double* src = which == 0 ? D.x : which == 1 ? D.y : which == 2 ? D.z : D.age;
        k_scatter<double><<<(n + 255) / 256, 256, 0, ctx->compute>>>(src, ctx->out8, ctx->d_pid, n);
        ctx->launches++; *dptr = ctx->out8;
    } else if (which == 4) {
        k_status<<<(n + 255) / 256, 256, 0, ctx->compute>>>(D, ctx->d_status);
        k_scatter<int><<<(n + 255) / 256, 256, 0, ctx->compute>>>(ctx->d_status, (int*)ctx->out8, ctx->d_pid, n);
        ctx->launches += 2; *dptr = ctx->out8;
    } else { ctx->err = "device_ptr: which out of range"; return LTGPU_E_ARG; }
    return LTGPU_OK;
}

int32_t ltgpu_export_device(ltgpu_ctx* ctx, int32_t which, void* dst_device)
{
    if (!ctx || !dst_device) return LTGPU_E_ARG;
    ARG(ctx->have_particles, "export_device before set_particles");
    CK(cudaSetDevice(ctx->device));
    LtDev& D = ctx->D; const int n = D.n, blocks = (n + 255) / 256;
    if (which >= 0 && which <= 3) {
        const double* src = which == 0 ? D.x : which == 1 ? D.y : which == 2 ? D.z : D.age;
        k_scatter<double><<<blocks, 256, 0, ctx->compute>>>(src, (double*)dst_device, ctx->d_pid, n);
        ctx->launches++;
    } else if (which == 4) {
        k_status<<<blocks, 256, 0, ctx->compute>>>(D, ctx->d_status);
        k_scatter<int><<<blocks, 256, 0, ctx->compute>>>(ctx->d_status, (int*)dst_device, ctx->d_pid, n);
        ctx->launches += 2;
    } else { ctx->err = "export_device: which out of range"; return LTGPU_E_ARG; }
    CK(cudaGetLastError());
    return LTGPU_OK;
}

int32_t ltgpu_fp64_peak(ltgpu_ctx* ctx, double* tflops)
{
    if (!ctx || !tflops) return LTGPU_E_ARG;
    CK(cudaSetDevice(ctx->device));
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, ctx->device));
    const int blocks = pr.multiProcessorCount * 4, iters = 4096;
    double* d = nullptr; CK(cudaMalloc(&d, sizeof(double) * blocks * 256));
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(a, ctx->compute));
        k_fp64_peak<<<blocks, 256, 0, ctx->compute>>>(d, iters, 0.999999, 1e-6);
        CK(cudaEventRecord(b, ctx->compute));
        CK(cudaEventSynchronize(b));
        float ms = 0; CK(cudaEventElapsedTime(&ms, a, b));
        ctx->launches++;
        const double flops = 2.0 * 64.0 * iters * 256.0 * blocks;
        if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(d);
    *tflops = best;
    return LTGPU_OK;
}

int32_t ltgpu_timer_start(ltgpu_ctx* ctx)
{
    if (!ctx) return LTGPU_E_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaEventRecord(ctx->t0, ctx->compute));
    return LTGPU_OK;
}
int32_t ltgpu_timer_stop(ltgpu_ctx* ctx, float* ms)
{
    if (!ctx || !ms) return LTGPU_E_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaEventRecord(ctx->t1, ctx->compute));
    CK(cudaEventSynchronize(ctx->t1));
    CK(cudaEventElapsedTime(ms, ctx->t0, ctx->t1));
    return LTGPU_OK;
}
#ifdef LT_DEBUG_TRACE
int32_t ltgpu_debug_trace(ltgpu_ctx* ctx, int64_t id, double* out, int32_t n)
{   // debug builds only (never compiled into the shipped library)
    if (!ctx->D.dbg) { void* q; cudaMalloc(&q, 4096 * 8); cudaMemset(q, 0, 4096 * 8); ctx->D.dbg = (double*)q; }
    cudaStreamSynchronize(ctx->compute);
    if (out) cudaMemcpy(out, ctx->D.dbg, sizeof(double) * n, cudaMemcpyDeviceToHost);
    ctx->D.dbg_id = id;
    return 0;
}
#endif
#ifdef LT_DEBUG_TRACE
int32_t ltgpu_debug_counters(ltgpu_ctx* ctx, unsigned long long* out, int32_t reset)
{   // debug builds only: [0] Newton cap, [1] Newton cycle, [2] max Newton NIT, [3] secant cap, [4] max secant NIT
    cudaStreamSynchronize(ctx->compute);
    cudaMemcpyFromSymbol(out, g_dbgcnt, sizeof(unsigned long long) * 8);
    cudaMemcpyFromSymbol(out + 8, g_dbgcase, sizeof(double) * 8);
    cudaMemcpyFromSymbol(out + 16, g_dbgcnt2, sizeof(unsigned long long) * 8);
    cudaMemcpyFromSymbol(out + 24, g_dbghist, sizeof(unsigned long long) * 64);
    if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(g_dbgcnt, z, sizeof z); cudaMemcpyToSymbol(g_dbgcnt2, z, sizeof z); unsigned long long zz[64] = {0}; cudaMemcpyToSymbol(g_dbghist, zz, sizeof zz); }
    return 0;
}
#endif
int32_t ltgpu_kernel_times(ltgpu_ctx* ctx, int32_t enable, float ms[4], int64_t* steps)
{
    if (!ctx) return LTGPU_E_ARG;
    CK(cudaSetDevice(ctx->device));
    if (ms) { for (int k = 0; k < 4; ++k) ms[k] = ctx->tacc[k]; }
    if (steps) *steps = ctx->tcount;
    for (int k = 0; k < 4; ++k) ctx->tacc[k] = 0; ctx->tcount = 0;
    if (enable && !ctx->tev[0]) for (int k = 0; k < 5; ++k) CK(cudaEventCreate(&ctx->tev[k]));
    ctx->timing = enable != 0;
    return LTGPU_OK;
}
int64_t ltgpu_launch_count(const ltgpu_ctx* ctx) { return ctx ? ctx->launches : 0; }
void* ltgpu_stream(ltgpu_ctx* ctx) { return ctx ? (void*)ctx->compute : nullptr; }

}  // extern "C"
