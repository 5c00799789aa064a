"""The C ABI driven without Python in the data path: tests/abi_harness.c is compiled
with gcc against include/ltrans_b200.h, linked to libltrans_b200.so and replays a dumped case; its
output must equal the ctypes run bit for bit.  Also: one process, two contexts at once."""
import ctypes as C
import os
import struct
import subprocess

import numpy as np
import pytest

from common import ROOT, SMALL, World, LtransLib, make_params, setup, run
from ltrans_b200.host.binding import DEFAULT_LIB, Params

CASE_KW = dict(HTurbOn=1, VTurbOn=1, Behavior=0, settlementon=0, mortality=0, ErrorFlag=1, TrackCollisions=0, ConstantHTurb=20.0)


def dump_case(path, w, prm, n, nexternal):
    g, b = w.grid(), w.bounds()
    x, y, z, dob, r, u, v = w.seed_particles(n)
    nrec = nexternal + 1
    recs = [w.record(k) for k in range(nrec)]

    def f64(a): return np.ascontiguousarray(a, np.float64).tobytes()
    def i32(a): return np.ascontiguousarray(a, np.int32).tobytes()
    dims = np.array([g["vi"], g["uj"], g["ui"], g["vj"], g["nRE"], g["nUE"], g["nVE"], len(b["land"]), len(b["bx"]), len(b["hx"]),
                     n, nrec, nexternal], np.int32)
    blobs = [bytes(prm), dims.tobytes()]
    blobs += [f64(g[k]) for k in ("rx", "ry", "ux", "uy", "vx", "vy", "depth", "angle")]
    blobs += [i32(g[k]) for k in ("rho_mask", "u_mask", "v_mask")]
    blobs += [f64(g[k]) for k in ("SC", "CS", "SCW", "CSW")]
    blobs += [i32(g[k]) for k in ("RE", "UE", "VE", "rAdj", "uAdj", "vAdj")]
    blobs += [f64(b["bnd_x"]), f64(b["bnd_y"]), i32(b["land"]), f64(b["bx"]), f64(b["by"]), f64(b["hx"]), f64(b["hy"]), i32(b["hid"])]
    blobs += [f64(x), f64(y), f64(z), f64(dob), i32(r), i32(u), i32(v)]
    for rec in recs:
        blobs += [np.ascontiguousarray(rec[k], np.float32).tobytes() for k in ("zeta", "u", "v", "w", "aks")]
    with open(path, "wb") as f:
        for bl in blobs:
            f.write(struct.pack("<q", len(bl))); f.write(bl)


def build_harness(tmp):
    exe = os.path.join(tmp, "abi_harness")
    libdir = os.path.dirname(DEFAULT_LIB)
    subprocess.check_call(["gcc", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), "-o", exe, os.path.join(ROOT, "tests", "abi_harness.c"),
                           "-L", libdir, "-lltrans_b200", "-Wl,-rpath," + libdir])
    return exe


def test_c_harness_compiles_and_links_against_the_header(tmp_path):
    """CPU: the harness builds with a C compiler from the public header alone and refuses to run without a
    device (no CPU fallback); the case file round-trips."""
    import torch
    exe = build_harness(str(tmp_path))
    w = World(**SMALL); n = 64
    prm = make_params(w, n, **CASE_KW)
    case = str(tmp_path / "case.bin")
    dump_case(case, w, prm, n, 2)
    assert os.path.getsize(case) > 100000
    if not torch.cuda.is_available():
        r = subprocess.run([exe, case, str(tmp_path / "out.bin")], capture_output=True, text=True)
        assert r.returncode == 3 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_c_harness_equals_ctypes_run_bit_for_bit(tmp_path):
    exe = build_harness(str(tmp_path))
    w = World(**SMALL); n = 1200; nexternal = 3
    prm = make_params(w, n, **CASE_KW)
    case, out = str(tmp_path / "case.bin"), str(tmp_path / "out.bin")
    dump_case(case, w, prm, n, nexternal)
    r = subprocess.run([exe, case, out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    raw = open(out, "rb").read()
    o = 0
    got = {}
    for k, t in (("x", np.float64), ("y", np.float64), ("z", np.float64), ("age", np.float64), ("status", np.int32), ("r_ele", np.int32)):
        nb = n * np.dtype(t).itemsize
        got[k] = np.frombuffer(raw[o:o + nb], t); o += nb
    counts = np.frombuffer(raw[o:o + 64], np.int64); o += 64
    nev = struct.unpack("<i", raw[o:o + 4])[0]
    g = LtransLib()
    setup(g, w, prm, n)
    assert run(g, w, nexternal) == (0, 0)
    f = g.fetch(); st = g.stats(); ev = g.drain_events(1 << 16)
    for k in got:
        assert np.array_equal(got[k], f[k]), k
    assert nev == len(ev) and np.array_equal(counts[[0, 1, 2, 3, 4, 6, 7]], st[[0, 1, 2, 3, 4, 6, 7]])
    assert counts[5] == nev
    g.destroy()


@pytest.mark.gpu
def test_two_contexts_in_one_process():
    """The reference host is one process; nothing in the ABI requires one process per GPU.  Two contexts
    (on two devices when the box has them, else both on device 0) stepped alternately from one thread, each
    on half of the particles, give the single-context run bit for bit (Philox is keyed on the global id)."""
    import torch
    w = World(**SMALL); n = 2000; h = n // 2
    prm = make_params(w, n, **CASE_KW)
    x, y, z, dob, r, u, v = w.seed_particles(n)

    def make(sl, first, device):
        g = LtransLib().create(make_params(w, sl.stop - sl.start, **CASE_KW), device=device)
        g.set_grid(w.grid()); g.set_bounds(w.bounds())
        g.set_particles(x[sl], y[sl], z[sl], dob[sl], None, r[sl], u[sl], v[sl], first_id=first)
        for k in range(3):
            g.push_hydro(w.record(k))
        return g
    dev2 = 1 if torch.cuda.device_count() > 1 else 0
    a, b, whole = make(slice(0, h), 1, 0), make(slice(h, n), h + 1, dev2), make(slice(0, n), 1, 0)
    for it in range(1, 13):
        a.step(1, it); b.step(1, it); whole.step(1, it)          # all three in flight at once
    assert a.sync() == (0, 0) and b.sync() == (0, 0) and whole.sync() == (0, 0)
    fa, fb, fw = a.fetch(), b.fetch(), whole.fetch()
    for k in ("x", "y", "z", "status", "r_ele"):
        assert np.array_equal(np.concatenate([fa[k], fb[k]]), fw[k]), k
    assert np.array_equal(a.stats()[:5] + b.stats()[:5], whole.stats()[:5])
    for g in (a, b, whole):
        g.destroy()
