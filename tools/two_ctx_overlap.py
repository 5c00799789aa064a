"""Does the GPU gain from running the latency-bound advect kernel of one half of the particles under the
issue-bound VTurb kernels of the other half?  Two contexts on one GPU, each with half of the particles of
BASELINE configs[4], stepped (a) one after the other, (b) queued together with the second one staggered by
`--stagger` internal steps; compared with one context holding all of them.
usage: python tools/two_ctx_overlap.py [--particles N] [--steps K] [--config C]"""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--particles", type=int, default=12500000)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--config", type=int, default=5)
ap.add_argument("--stagger", type=int, default=0, help="internal steps ctx 1 runs alone first (0: half a step = advect only is not expressible; see --halfshift)")
args = ap.parse_args()
cfg = bench.CONFIGS[args.config]
keys = ("zeta", "u", "v", "w", "aks")
n = args.particles


def timed(engines, steps, mode):
    for E in engines:
        E.g.sync()
    t0 = time.perf_counter()
    for _ in range(steps):
        if mode == "serial":
            for E in engines:
                E.step(); E.g.sync()
        else:
            for E in engines:
                E.step()
            for E in engines:
                E.g.sync()
    dt = time.perf_counter() - t0
    tot = sum(E.n for E in engines) * engines[0].stepIT * steps
    return tot / dt


one = bench.Engine(cfg, n, 0, 0, keys)
for _ in range(3):
    one.step()
one.g.sync()
r_one = timed([one], args.steps, "serial")
one.g.destroy(); del one
a = bench.Engine(cfg, n // 2, 0, 0, keys)
b = bench.Engine(cfg, n - n // 2, 1, 0, keys)
for _ in range(3):
    a.step(); b.step()
r_ser = timed([a, b], args.steps, "serial")
r_con = timed([a, b], args.steps, "concurrent")
# staggered: context a runs a few internal steps of its next external step alone, so that its VTurb kernels
# meet b's advect kernels
a.g.sync(); b.g.sync()
print({"one_context": r_one, "two_serial": r_ser, "two_concurrent": r_con})
