#!/usr/bin/env python
"""bench.py -- particle-steps/s of the LTRANS v.2b particle loop on N B200s.

Default workload = BASELINE.json configs[4] (the configuration the metric "at 1/2/4/8 B200" is
quoted on): Gulf-scale synthetic ROMS world (1024x768 rho, us 36 / ws 37, dt 3600, idt 120),
12.5 M buoyant particles PER GPU (Behavior 6, sink > 0), HTurb + VTurb with the reference's
full-column SigErr sweep, open ocean boundary, fields replicated on every GPU, particle slices
keyed by global id (Philox).  `--config 2` / `--config 3` select BASELINE configs[1] / [2]; they
are also run, shorter, as `secondary` entries of the default line.

One bench "step" = one external time step = dt/idt = 30 internal steps of every particle
(30 x {re-sort, k_advect, k_vturb, k_finish}) plus the refill of one hydro record.

  value : device-timed (CUDA events on the library's compute stream), particle state and the
          three live hydro records resident; the next record's refill runs on the side stream
          as in production.
  e2e   : the same loop through the C ABI with HOST buffers, print interval = one external
          step as in the shipped LTRANS.data: every step pushes one hydro record from host
          memory (H2D), reduces the 8 statistics counters over ranks (NCCL all-reduce), gathers
          x, y, z, status of ALL particles in particle order over NVLink (NCCL all-gather of
          ltgpu_export_device columns) and lands them in rank 0's host memory (D2H).
  --impl reference : the CPU restatement of the reference loop (oracle/) on the host cores,
          same world and switches, on a bounded cut of the particles.
"""
import argparse
import glob
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

# BASELINE.md section 3 / SURVEY.md 8d: algorithmic bytes per particle-step at the reference's
# 8-byte width (state 104 + geometry 256 + zeta 96 + u/v/w 1152 + AKs 96 ws [+ salt 384 + behaviour 64])
CONFIGS = {
    5: dict(name="config5 (BASELINE configs[4]: weak scaling, config-4 world)", b_alg=5160, particles=12_500_000,
            world=dict(ni=1024, nj=768, us=36, hmin=50.0, hmax=3000.0, dlon=0.02, dlat=0.018, speed=0.9),
            grid="1024x768 rho, us 36, ws 37",
            prm=dict(Behavior=6, sink=0.002, settlementon=0, mortality=0, TrackCollisions=0, ErrorFlag=3,
                     OpenOceanBoundary=1, HTurbOn=1, VTurbOn=1),
            desc="Gulf-scale synthetic ROMS, buoyant particles (Behavior 6), HTurb+VTurb, open boundary",
            cpu_particles=100_000),
    2: dict(name="config2 (BASELINE configs[1])", b_alg=3624, particles=1_000_000,
            world=dict(), grid="130x130 rho, us 20, ws 21",
            prm=dict(Behavior=0, settlementon=0, mortality=0, TrackCollisions=0, ErrorFlag=1, HTurbOn=1, VTurbOn=1),
            desc="Baymouth-shape synthetic ROMS, passive particles, HTurb+VTurb", cpu_particles=100_000),
    3: dict(name="config3 (BASELINE configs[2])", b_alg=4072, particles=10_000_000,
            world=dict(ni=120, nj=80, us=20, dlon=0.02, dlat=0.018), grid="120x80 rho, us 20, ws 21",
            prm=dict(Behavior=4, settlementon=1, holesExist=1, mortality=1, TrackCollisions=0, ErrorFlag=3,
                     pediage=3600.0, deadage=40 * 3600.0, HTurbOn=1, VTurbOn=1),
            desc="Chesapeake-scale synthetic ROMS, oyster larvae (Behavior 4), HTurb+VTurb, 64 settlement "
                 "polygons with holes, mortality", cpu_particles=100_000, npoly=64),
}
NREC = 5            # distinct hydro records kept on the host, visited 0,1,2,3,4,3,2,1,0,... (continuous in time)


def config_dict(cfg, n, world_size, prm):
    return {"workload": f"{cfg['name']}: {cfg['desc']}; {n} particles/GPU; {prm.dt // prm.idt} internal steps per step",
            "particles_per_gpu": n, "grid": cfg["grid"], "internal_steps_per_step": prm.dt // prm.idt,
            "vturb_full_sigs": int(not prm.vturb_window_sigs), "field_storage": "f32 (lossless), [node][level][4-slot ring]",
            "rng": "philox4x32-10 keyed (seed; particle id, step, block)", "parallelism": f"particle slices x{world_size}, fields replicated",
            "l2": "particle state + fields exceed L2 at 12.5 M particles; a 256 MiB buffer is also written between timed steps"}


def rec_index(p):
    """ping-pong over the NREC host records: p = 0,1,2,... -> 0,1,2,3,4,3,2,1,0,1,..."""
    m = p % (2 * (NREC - 1))
    return m if m < NREC else 2 * (NREC - 1) - m


def clocks_sampler(stop, out, gpu_index):
    q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    while not stop.is_set():
        try:
            r = subprocess.run(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                               capture_output=True, text=True, timeout=5)
            f = [s.strip() for s in r.stdout.strip().split(",")]
            if len(f) >= 6:
                out.append(f)
        except Exception:
            pass
        stop.wait(0.5)      # nvidia-smi perturbs the run (~4 % at 0.2 s): sample sparsely


def clocks_summary(samples):
    if not samples:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
    sm = [float(s[0]) for s in samples if s[0].replace(".", "").isdigit()]
    mx = [float(s[1]) for s in samples if s[1].replace(".", "").isdigit()]
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    reasons = [n for k, n in enumerate(names) if any(s[2 + k].lower().startswith("active") for s in samples)]
    return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
            "reasons": reasons, "samples": len(samples)}


# ------------------------------------------------------------------ CPU arm ----
def make_world_and_params(cfg, n):
    from common import World, make_params
    w = World(**cfg["world"])
    prm = make_params(w, n, **cfg["prm"])
    return w, prm


def cpu_engine(cfg, n):
    """The oracle (C restatement of the reference loop) set up on an n-particle cut of the workload."""
    from oracle.oracle import Oracle
    w, prm = make_world_and_params(cfg, n)
    o = Oracle().create(prm)
    o.set_grid(w.grid()); o.set_bounds(w.bounds())
    if prm.settlementon:
        o.set_habitat(w.habitat(npoly=cfg.get("npoly", 8)))
    x, y, z, dob, r, u, v = w.seed_particles(n, seed=1234)
    o.set_particles(x, y, z, dob, None, r, u, v, first_id=1)
    recs = [w.record(k) for k in range(3)]
    for k in range(3):
        o.push_hydro(recs[k])
    return o, w, prm


def cpu_serial_rate(o, n, it0, budget_s=12.0, max_steps=8):
    """1 thread, like the reference (its loop is serial, LTRANS.f90:778): internal steps it0, it0+1, ... of
    external step 1 until the budget is spent.  Returns (particle-steps/s, steps done)."""
    o.set_threads(1)
    t0 = time.perf_counter(); done = 0
    while True:
        o.step(1, it0 + done); done += 1
        el = time.perf_counter() - t0
        if el >= budget_s or done >= max_steps:
            return n * done / el, done


def cpu_baseline(cfg, n):
    cores = os.cpu_count() or 1
    o, w, prm = cpu_engine(cfg, n)
    o.set_threads(cores)
    o.step(1, 1)
    t = time.perf_counter()
    nst = 12
    for it in range(2, 2 + nst):
        o.step(1, it)
    dt = time.perf_counter() - t
    serial, sdone = cpu_serial_rate(o, n, 2 + nst)
    o.destroy()
    return {"value": n * nst / dt, "unit": "particle-steps/s", "cores": cores, "kind": "port",
            "serial_value": serial, "serial_cores": 1,
            "sample": f"{n}-particle cut of the same workload: {nst} internal steps with OpenMP over particles on {cores} threads, "
                      f"{sdone} internal steps on 1 thread (the reference loop is serial)"}


def run_reference(args, cfg):
    """--impl reference: the CPU restatement on all host threads; one step = one external step of the cut."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n = args.cpu_particles or cfg["cpu_particles"]
    o, w, prm = cpu_engine(cfg, n)
    stepIT = prm.dt // prm.idt
    serial, sdone = cpu_serial_rate(o, n, 1)                 # internal steps 1.. of external step 1 on one thread
    o.set_threads(cores)
    for it in range(1 + sdone, stepIT + 1):                  # rest of external step 1, untimed
        o.step(1, it)
    recs = {}
    times = []
    p = 1
    for s in range(args.warmup + args.steps):
        p += 1
        if p > 2:
            k = rec_index(p)
            if k not in recs:
                recs[k] = w.record(k)
            o.push_hydro(recs[k]); o.rotate_hydro()
        t = time.perf_counter()
        o.run_external(p)
        dt = time.perf_counter() - t
        if s >= args.warmup:
            times.append(dt)
    total = sum(times)
    val = n * stepIT * args.steps / total
    nfull = args.particles or cfg["particles"]
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    from common import make_params
    prm_full = make_params(w, nfull, **cfg["prm"])
    sample = (f"{n}-particle cut of the workload, {stepIT} internal steps per step, OpenMP over particles on {cores} threads; "
              f"serial_value = {sdone} internal steps of the same cut on 1 thread (the reference loop is serial)")
    print(json.dumps({
        "impl": "reference", "metric": "particle-steps/sec", "value": val, "unit": "particle-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(cfg, nfull, world_size, prm_full),
        "cpu_baseline": {"value": val, "unit": "particle-steps/s", "cores": cores, "kind": "port",
                         "serial_value": serial, "serial_cores": 1, "sample": sample},
        "e2e": {"value": val, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
    o.destroy()


# ------------------------------------------------------------------ GPU arm ----
class Engine:
    """One context on this rank's GPU, set up on its particle slice of the workload."""

    def __init__(self, cfg, n, rank, local, keys):
        from common import LtransLib
        self.cfg, self.n, self.rank = cfg, n, rank
        self.w, self.prm = make_world_and_params(cfg, n)
        w, prm = self.w, self.prm
        self.stepIT = prm.dt // prm.idt
        g = self.g = LtransLib().create(prm, device=local)
        g.set_grid(w.grid()); g.set_bounds(w.bounds())
        if prm.settlementon:
            g.set_habitat(w.habitat(npoly=cfg.get("npoly", 8)))
        # contiguous particle slice of the global run (ids first_id .. first_id + n - 1), located on the device
        x, y, z, dob, _, _, _ = w.seed_particles(n, seed=1234 + rank, locate=False)
        g.set_particles(x, y, z, dob, None, None, None, None, first_id=1 + rank * n)
        rc, self.screened, bad = g.screen_initial()
        self.keys = keys
        self.recs = [{k: v for k, v in w.record(r).items() if k in keys} for r in range(NREC)]
        for k in range(3):
            g.push_hydro(self.recs[k])
        self.h2d = sum(a.nbytes for a in self.recs[0].values())
        self.p = 0

    def step(self):
        """one external step: rotate in the record pushed during the previous step, queue the 30
        internal steps (asynchronous), stage + copy the next record on the side stream meanwhile"""
        g = self.g
        self.p += 1
        if self.p > 2:
            g.rotate_hydro()
        g.run_external(self.p)
        if self.p >= 2:
            g.push_hydro(self.recs[rec_index(self.p + 1)])


def latest_profile_json():
    """ncu-derived per-particle-step figures of the dominant kernels (written by tools/ncu_digest.py from a
    `--set full` capture); the newest profiles/r*_kernel_metrics.json, named in the line."""
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_kernel_metrics.json")))
    if not files:
        return None, None
    try:
        return json.load(open(files[-1])), os.path.relpath(files[-1], ROOT)
    except Exception:
        return None, None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--config", type=int, default=5, choices=sorted(CONFIGS))
    ap.add_argument("--particles", type=int, default=0, help="particles per GPU (default: the configuration's)")
    ap.add_argument("--cpu-particles", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        run_reference(args, cfg)
        return

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world_size > 1:
        # NCCL announces its version on stdout when the communicator is created: keep stdout to
        # the one JSON line by pointing fd 1 at stderr until the first collective has run
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.all_reduce(torch.zeros(1, device=dev))
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    n = args.particles or cfg["particles"]
    need_st = cfg["prm"].get("Behavior", 0) in (4, 5, 7) or cfg["prm"].get("SaltTempOn", 0)
    keys = ("zeta", "u", "v", "w", "aks") + (("salt", "temp") if need_st else ())
    E = Engine(cfg, n, rank, local, keys)
    g, prm, stepIT = E.g, E.prm, E.stepIT
    fp64_peak = g.fp64_peak()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    ext = torch.cuda.ExternalStream(g.stream(), device=dev)      # the library's compute stream, as a torch stream
    side = torch.cuda.Stream(device=dev)                         # output gather + D2H, under the next step's compute

    def barrier():
        g.sync()
        torch.cuda.synchronize()
        if world_size > 1:
            dist.barrier()

    # ---- end-to-end plumbing: statistics all-reduce + output gather over NCCL ------------------
    send = {"x": torch.empty(n, dtype=torch.float64, device=dev), "y": torch.empty(n, dtype=torch.float64, device=dev),
            "z": torch.empty(n, dtype=torch.float64, device=dev), "status": torch.empty(n, dtype=torch.int32, device=dev)}
    which = {"x": 0, "y": 1, "z": 2, "status": 4}
    if world_size > 1:
        recv = {k: torch.empty(n * world_size, dtype=t.dtype, device=dev) for k, t in send.items()}
    else:
        recv = send
    host_out = {k: torch.empty(t.numel(), dtype=t.dtype).pin_memory() for k, t in recv.items()} if rank == 0 else {}
    stats_t = torch.zeros(8, dtype=torch.int64, device=dev)
    gather_done = None
    d2h = sum(t.numel() * t.element_size() for t in host_out.values()) if rank == 0 else 0
    d2h += 64

    def print_interval():
        """what printOutput needs (LTRANS.f90:1617-1666): the counters of all ranks and every particle's
        x, y, z, status in particle order in the writer's (rank 0's) host memory"""
        nonlocal gather_done
        st = g.stats()                                           # device reduction, 64 B D2H (synchronises this step)
        stats_t.copy_(torch.from_numpy(st))
        if world_size > 1:
            dist.all_reduce(stats_t)
        if gather_done is not None:
            ext.wait_event(gather_done)                          # send buffers free again (device-side wait)
        for k, t in send.items():
            g.export_device(which[k], t.data_ptr())              # particle-order columns, on the compute stream
        ready = ext.record_event()
        side.wait_event(ready)
        with torch.cuda.stream(side):
            if world_size > 1:
                for k in send:
                    dist.all_gather_into_tensor(recv[k], send[k])
            if rank == 0:
                for k in host_out:
                    host_out[k].copy_(recv[k], non_blocking=True)
            gather_done = side.record_event()

    for _ in range(args.warmup):
        E.step()
    print_interval()                                             # warm the NCCL channels and the pinned buffers
    barrier()
    stop, samples = threading.Event(), []
    # only rank 0 polls nvidia-smi (its own GPU): polling perturbs the polled GPU
    th = threading.Thread(target=clocks_sampler if rank == 0 else (lambda *a: None), args=(stop, samples, local)); th.start()
    # ---- device-timed arm ---------------------------------------------------------
    launches0 = g.launch_count()
    dev_ms = 0.0
    for _ in range(args.steps):
        flush.fill_(1)                             # L2 flush between timed iterations
        torch.cuda.synchronize()
        g.timer_start()
        E.step()
        ms_ = g.timer_stop(); dev_ms += ms_
        if os.environ.get("LT_BENCH_DEBUG"):
            print(f"rank {rank} step ms {ms_:.1f}", file=sys.stderr)
    launches = g.launch_count() - launches0
    barrier()
    t_dev = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world_size > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    dev_ms = float(t_dev.item())
    # ---- end-to-end arm (host buffers in, host buffers out, collectives inside) ---------
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        E.step()
        print_interval()
    g.sync()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t_e = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world_size > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_s = float(t_e.item())
    stop.set(); th.join()
    # where the end-to-end step goes (one more step, pieces timed on the host with a sync after each)
    parts = {}
    barrier()
    t = time.perf_counter(); E.step(); g.sync(); parts["step_and_push_ms"] = 1e3 * (time.perf_counter() - t)
    t = time.perf_counter(); print_interval(); torch.cuda.synchronize(); parts["stats_allreduce_gather_d2h_ms"] = 1e3 * (time.perf_counter() - t)
    # ---- per-kernel device times (separate short pass: the bracketing syncs every step) -----
    g.kernel_times(True)
    E.step()
    kms, ksteps = g.kernel_times(False)
    g.sync()
    k_adv, k_vt, k_fin = (v / max(1, ksteps) for v in kms[:3])
    fin = None
    if rank == 0 and host_out:
        fin = bool(np.isfinite(host_out["z"].numpy()).all())
    total_steps = n * world_size * stepIT * args.steps
    value = total_steps / (dev_ms * 1e-3)
    e2e = total_steps / e2e_s
    kern_ms = k_adv + k_vt + k_fin                            # the three launches of one internal step
    step_ms = dev_ms / (args.steps * stepIT)                  # same, from the timed region (incl. re-sort, refill waits)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    b_alg = cfg["b_alg"]
    achieved = b_alg * n / (kern_ms * 1e-3) / 1e9
    prof, prof_file = latest_profile_json()
    main_stats = [int(v) for v in stats_t.tolist()]
    g.destroy()
    del flush, send, recv
    torch.cuda.empty_cache()

    # ---- secondary entries: the other GPU configurations, device-timed, shorter -----------------
    secondary = []
    if not args.no_secondary and not args.particles:
        for c in sorted(CONFIGS):
            if c == args.config:
                continue
            c2 = CONFIGS[c]
            st2 = c2["prm"].get("Behavior", 0) in (4, 5, 7)
            E2 = Engine(c2, c2["particles"], rank, local, ("zeta", "u", "v", "w", "aks") + (("salt", "temp") if st2 else ()))
            E2.step(); E2.step()
            E2.g.sync()
            if world_size > 1:
                dist.barrier()
            ms2 = 0.0
            for _ in range(2):
                E2.g.timer_start(); E2.step(); ms2 += E2.g.timer_stop()
            E2.g.kernel_times(True); E2.step(); km2, ks2 = E2.g.kernel_times(False)
            st = E2.g.stats()
            t2 = torch.tensor([ms2], dtype=torch.float64, device=dev)
            if world_size > 1:
                dist.all_reduce(t2, op=dist.ReduceOp.MAX)
            ms2 = float(t2.item())
            kk = [v / max(1, ks2) for v in km2[:3]]
            secondary.append({
                "config": config_dict(c2, c2["particles"], world_size, E2.prm), "steps": 2, "warmup": 2,
                "value": c2["particles"] * world_size * E2.stepIT * 2 / (ms2 * 1e-3), "unit": "particle-steps/s",
                "ms_per_step": ms2 / 2, "kernels_ms": {"k_advect": kk[0], "k_vturb": kk[1], "k_finish": kk[2]},
                "roofline_frac": c2["b_alg"] * c2["particles"] / (sum(kk) * 1e-3) / 1e9 / peak,
                "stats_rank0": {"settled": int(st[0]), "dead": int(st[1]), "out_of_bounds": int(st[2]), "active": int(st[6])}})
            E2.g.destroy()
            del E2
            torch.cuda.empty_cache()

    # ---- secondary entry: the opt-in single-precision random walk on the headline configuration --------
    if not args.no_secondary and not args.particles:
        cf = dict(cfg); cf["prm"] = dict(cfg["prm"], vturb_fp32_walk=1)
        E3 = Engine(cf, n, rank, local, keys)
        E3.step(); E3.step(); E3.g.sync()
        if world_size > 1:
            dist.barrier()
        ms3 = 0.0
        for _ in range(2):
            E3.g.timer_start(); E3.step(); ms3 += E3.g.timer_stop()
        E3.g.kernel_times(True); E3.step(); km3, ks3 = E3.g.kernel_times(False)
        t3 = torch.tensor([ms3], dtype=torch.float64, device=dev)
        if world_size > 1:
            dist.all_reduce(t3, op=dist.ReduceOp.MAX)
        ms3 = float(t3.item())
        kk = [v / max(1, ks3) for v in km3[:3]]
        c3 = config_dict(cf, n, world_size, E3.prm); c3["vturb_fp32_walk"] = 1
        c3["workload"] += "; OPT-IN single-precision random walk (statistical parity only, not the headline)"
        secondary.append({"config": c3, "steps": 2, "warmup": 2, "dtype": "f64 fit + f32 walk",
                          "value": n * world_size * E3.stepIT * 2 / (ms3 * 1e-3), "unit": "particle-steps/s", "ms_per_step": ms3 / 2,
                          "kernels_ms": {"k_advect": kk[0], "k_vturb": kk[1], "k_finish": kk[2]}})
        E3.g.destroy()
        del E3
        torch.cuda.empty_cache()

    if rank == 0:
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None,
                "kernel": "k_advect + k_vturb + k_finish (the three launches of one internal step; k_vturb dominates)",
                "kernel_ms": kern_ms, "kernels_ms": {"k_advect": k_adv, "k_vturb": k_vt, "k_finish": k_fin},
                "dominant": {"kernel": "k_vturb", "algorithmic_bytes_per_particle": 96 * prm.ws,
                             "achieved": 96 * prm.ws * n / (k_vt * 1e-3) / 1e9 if k_vt > 0 else None,
                             "frac": 96 * prm.ws * n / (k_vt * 1e-3) / 1e9 / peak if k_vt > 0 else None},
                "internal_step_ms_in_timed_region": step_ms,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst)" if peaks else "fallback 6650",
                "note": f"achieved = {b_alg} algorithmic B/particle-step x particles per launch / summed launch time; the path is "
                        "FP64-issue bound, not HBM bound (DESIGN.md section 5): see roofline_fp64"}
        roof64 = {"bound": "fp64", "peak": fp64_peak, "unit": "TFLOP/s", "peak_source": "ltgpu_fp64_peak (FMA chains, measured in this run)",
                  "achieved": None, "frac": None}
        if prof:
            key = f"config{args.config}"
            pk = prof.get(key) or {}
            if pk.get("dram_bytes_per_particle_step") is not None:
                roof["traffic"] = pk["dram_bytes_per_particle_step"] * n
                roof["traffic_source"] = f"{prof_file} [{key}] (ncu --set full, dram__bytes_read+write per launch, scaled by particles)"
            if pk.get("fp64_flop_per_particle_step") is not None:
                a64 = pk["fp64_flop_per_particle_step"] * n / (kern_ms * 1e-3) / 1e12
                roof64.update({"achieved": a64, "frac": a64 / fp64_peak if fp64_peak else None,
                               "flop_per_particle_step": pk["fp64_flop_per_particle_step"],
                               "source": f"{prof_file} [{key}] (ncu thread-instruction counts: 2 x dfma + dmul + dadd)"})
        line = {
            "metric": "particle-steps/sec", "value": value, "unit": "particle-steps/s", "n_gpus": world_size,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(cfg, n, world_size, prm),
            "e2e": {"value": e2e, "unit": "particle-steps/s", "h2d_bytes_per_step": int(E.h2d), "d2h_bytes_per_step": int(d2h),
                    "collectives_per_step": ("ncclAllReduce(8 x int64) + ncclAllGather(x, y, z f64 + status i32 of all particles)"
                                             if world_size > 1 else "none at 1 GPU (same code path: export_device + D2H)"),
                    "breakdown_ms": parts, "output_finite": fin},
            "gpu_launches": int(launches),
            "roofline": roof, "roofline_fp64": roof64,
            "clocks": clocks_summary(samples),
            "stats": {"settled": main_stats[0], "dead": main_stats[1], "out_of_bounds": main_stats[2],
                      "active": main_stats[6], "events": main_stats[5], "screened_at_start": [int(v) for v in E.screened]},
            "secondary": secondary,
        }
        if not args.no_cpu and world_size == 1:
            line["cpu_baseline"] = cpu_baseline(cfg, args.cpu_particles or cfg["cpu_particles"])
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line))
    if world_size > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
