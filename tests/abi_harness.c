/* abi_harness.c -- drives libltrans_b200.so through include/ltrans_b200.h from plain C, the way the
 * Fortran host of fortran/ does through ISO_C_BINDING: no Python, no ctypes, no torch.
 *
 *   abi_harness CASE.bin OUT.bin [device]
 *
 * CASE.bin is written by tests/test_abi_harness.py (dump_case): the parameter struct, the grid /
 * boundary / particle tables and the hydro records of a small run, each array as int64 byte count +
 * raw bytes in the fixed order read below.  The harness replays run_LTRANS' loop
 * (LTRANS.f90:156-161, 548-614) and writes x, y, z, age, status, rho element, the 8 statistics
 * counters and the number of events to OUT.bin; the test compares them bit for bit with the same run
 * made through the ctypes binding.  Test infrastructure only. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "ltrans_b200.h"

static FILE* in;
static void* rd(int64_t* nbytes)
{
    int64_t nb = 0;
    if (fread(&nb, sizeof nb, 1, in) != 1) { fprintf(stderr, "abi_harness: truncated case file\n"); exit(2); }
    void* p = malloc(nb > 0 ? (size_t)nb : 1);
    if (nb > 0 && fread(p, 1, (size_t)nb, in) != (size_t)nb) { fprintf(stderr, "abi_harness: truncated array\n"); exit(2); }
    if (nbytes) *nbytes = nb;
    return p;
}
#define CHECK(call) do { int32_t rc_ = (call); if (rc_ != LTGPU_OK) { \
    fprintf(stderr, "abi_harness: %s -> status %d: %s\n", #call, (int)rc_, ltgpu_last_error(ctx)); return 3; } } while (0)

int main(int argc, char** argv)
{
    if (argc < 3) { fprintf(stderr, "usage: abi_harness CASE.bin OUT.bin [device]\n"); return 1; }
    in = fopen(argv[1], "rb");
    if (!in) { perror(argv[1]); return 1; }
    int device = argc > 3 ? atoi(argv[3]) : 0;
    int64_t nb;
    ltgpu_params* prm = (ltgpu_params*)rd(&nb);
    if (nb != (int64_t)sizeof(ltgpu_params)) { fprintf(stderr, "abi_harness: ltgpu_params is %zu bytes here, %lld in the case\n", sizeof(ltgpu_params), (long long)nb); return 2; }
    int32_t* d = (int32_t*)rd(NULL);            /* vi uj ui vj nRE nUE nVE nbounds maxbound maxisland n nrec nexternal */
    const int32_t vi = d[0], uj = d[1], ui = d[2], vj = d[3], nRE = d[4], nUE = d[5], nVE = d[6];
    const int32_t nbounds = d[7], maxbound = d[8], maxisland = d[9], n = d[10], nrec = d[11], nexternal = d[12];
    double *rx = rd(NULL), *ry = rd(NULL), *ux = rd(NULL), *uy = rd(NULL), *vx = rd(NULL), *vy = rd(NULL), *depth = rd(NULL), *angle = rd(NULL);
    int32_t *rmask = rd(NULL), *umask = rd(NULL), *vmask = rd(NULL);
    double *SC = rd(NULL), *CS = rd(NULL), *SCW = rd(NULL), *CSW = rd(NULL);
    int32_t *RE = rd(NULL), *UE = rd(NULL), *VE = rd(NULL), *rAdj = rd(NULL), *uAdj = rd(NULL), *vAdj = rd(NULL);
    double *bnd_x = rd(NULL), *bnd_y = rd(NULL); int32_t* land = rd(NULL);
    double *bx = rd(NULL), *by = rd(NULL), *hx = rd(NULL), *hy = rd(NULL); int32_t* hid = rd(NULL);
    double *x = rd(NULL), *y = rd(NULL), *z = rd(NULL), *dob = rd(NULL);
    int32_t *re = rd(NULL), *ue = rd(NULL), *ve = rd(NULL);

    ltgpu_ctx* ctx = NULL;
    int32_t rc = ltgpu_create(prm, device, &ctx);
    if (rc != LTGPU_OK) { fprintf(stderr, "abi_harness: ltgpu_create -> %d (no CPU fallback)\n", (int)rc); return 3; }
    CHECK(ltgpu_set_grid(ctx, vi, uj, ui, vj, rx, ry, ux, uy, vx, vy, depth, angle, rmask, umask, vmask, SC, CS, SCW, CSW,
                         RE, UE, VE, nRE, nUE, nVE, rAdj, uAdj, vAdj));
    CHECK(ltgpu_set_bounds(ctx, nbounds, bnd_x, bnd_y, land, maxbound, bx, by, maxisland, hx, hy, hid));
    CHECK(ltgpu_set_particles(ctx, n, 1, x, y, z, dob, NULL, re, ue, ve));
    float** rec = (float**)malloc(sizeof(float*) * 5 * (size_t)nrec);
    for (int k = 0; k < nrec; ++k) for (int f = 0; f < 5; ++f) rec[5 * k + f] = (float*)rd(NULL);
    fclose(in);
    for (int k = 0; k < 3; ++k) CHECK(ltgpu_push_hydro(ctx, LTGPU_F32, rec[5 * k], rec[5 * k + 1], rec[5 * k + 2], rec[5 * k + 3], rec[5 * k + 4], NULL, NULL));
    for (int p = 1; p <= nexternal; ++p) {
        if (p > 2) {
            CHECK(ltgpu_push_hydro(ctx, LTGPU_F32, rec[5 * p], rec[5 * p + 1], rec[5 * p + 2], rec[5 * p + 3], rec[5 * p + 4], NULL, NULL));
            CHECK(ltgpu_rotate_hydro(ctx));
        }
        CHECK(ltgpu_run_external(ctx, p));
        int32_t bad = 0;
        CHECK(ltgpu_sync(ctx, &bad));
    }
    double *ox = malloc(8 * (size_t)n), *oy = malloc(8 * (size_t)n), *oz = malloc(8 * (size_t)n), *oage = malloc(8 * (size_t)n);
    int32_t *ost = malloc(4 * (size_t)n), *ore = malloc(4 * (size_t)n);
    CHECK(ltgpu_fetch(ctx, ox, oy, oz, oage, ost, NULL, NULL, NULL, NULL, NULL, NULL, ore, NULL, NULL));
    int64_t counts[8];
    CHECK(ltgpu_stats(ctx, counts));
    ltgpu_event* ev = malloc(sizeof(ltgpu_event) * 65536); int32_t nev = 0;
    rc = ltgpu_drain_events(ctx, ev, 65536, &nev);
    if (rc != LTGPU_OK && rc != LTGPU_W_EVENTS_LOST) { fprintf(stderr, "abi_harness: drain_events -> %d\n", (int)rc); return 3; }
    FILE* out = fopen(argv[2], "wb");
    if (!out) { perror(argv[2]); return 1; }
    fwrite(ox, 8, (size_t)n, out); fwrite(oy, 8, (size_t)n, out); fwrite(oz, 8, (size_t)n, out); fwrite(oage, 8, (size_t)n, out);
    fwrite(ost, 4, (size_t)n, out); fwrite(ore, 4, (size_t)n, out); fwrite(counts, 8, 8, out); fwrite(&nev, 4, 1, out);
    fclose(out);
    printf("abi_harness: %d particles, %d external steps, %lld kernel launches, %d events\n", (int)n, (int)nexternal,
           (long long)ltgpu_launch_count(ctx), (int)nev);
    ltgpu_destroy(ctx);
    return 0;
}
