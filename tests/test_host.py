"""CPU-side checks of the product's host layer: the C-ABI library loads and exports every
symbol include/ltrans_b200.h declares, struct layouts agree, there is no CPU fallback,
the synthetic world's setup tables are self-consistent, and the multi-rank host logic
(particle slices, statistics reduction, output gather) is exact under gloo."""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from common import ROOT, SMALL, World, make_params, setup, run
from ltrans_b200.host.binding import DEFAULT_LIB, LtransLib, LtransError, Params, Event
from ltrans_b200.host import shard

HEADER = os.path.join(ROOT, "include", "ltrans_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ltgpu_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(DEFAULT_LIB), "run __graft_entry__.build()"
    lib = C.CDLL(DEFAULT_LIB)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n


def test_struct_layouts_match_the_header():
    code = r'''#include <stdio.h>
#include <stddef.h>
#include "ltrans_b200.h"
int main(void){ printf("%zu %zu %zu %zu %zu %zu\n", sizeof(ltgpu_params), offsetof(ltgpu_params, z0),
  offsetof(ltgpu_params, deadage), offsetof(ltgpu_params, PI), offsetof(ltgpu_params, field_dtype), sizeof(ltgpu_event)); return 0; }'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(code)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "t"), os.path.join(d, "t.c")])
        out = subprocess.check_output([os.path.join(d, "t")]).decode().split()
    got = [C.sizeof(Params), Params.z0.offset, Params.deadage.offset, Params.PI.offset, Params.field_dtype.offset, C.sizeof(Event)]
    assert [int(v) for v in out] == got


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = LtransLib()
    with pytest.raises(LtransError) as e:
        lib.create(Params.shipped())
    assert "status 3" in str(e.value)              # LTGPU_E_NODEVICE
    with pytest.raises(LtransError):
        LtransLib(path="/nonexistent/libltrans_b200.so")


def test_world_tables_are_consistent():
    w = World(**SMALL)
    g, b = w.grid(), w.bounds()
    for E, adj, n in ((g["RE"], g["rAdj"], g["nRE"]), (g["UE"], g["uAdj"], g["nUE"]), (g["VE"], g["vAdj"], g["nVE"])):
        assert E.shape == (n, 4) and adj.shape == (10, n)
        assert np.all(adj[0] == np.arange(1, n + 1)) and np.all(adj[9] == 0)
        for e in (0, n // 2, n - 1):                       # neighbours share a node, ascending ids
            nb = adj[1:, e][adj[1:, e] > 0]
            assert np.all(np.diff(nb) > 0)
            for q in nb:
                assert len(set(E[e]) & set(E[q - 1])) >= 1
    assert b["bx"][0] == b["bx"][-1] and b["by"][0] == b["by"][-1]
    assert len(b["land"]) == len(b["bx"]) - 1 + len(b["hx"]) - b["nislands"]
    assert (b["land"] == 0).sum() > 0 and b["nislands"] >= 2
    for k in np.unique(b["hid"]):
        sel = b["hid"] == k
        assert b["hx"][sel][0] == b["hx"][sel][-1]
    x, y, z, dob, r, u, v = w.seed_particles(500)
    assert r.min() >= 1 and u.min() >= 1 and v.min() >= 1
    rec = w.record(2)
    assert rec["u"].shape == (w.us, w.nj, w.ni - 1) and rec["aks"].shape == (w.ws, w.nj, w.ni)
    assert rec["zeta"].dtype == np.float32 and np.abs(rec["u"]).max() < 1.0


def test_slices_cover_all_particles():
    for n, ws in ((10, 3), (1000, 8), (7, 8), (12500000 * 8, 8)):
        seen = 0
        for r in range(ws):
            lo, hi = shard.slice_for_rank(n, r, ws)
            assert lo == seen or lo == n
            seen = max(seen, hi)
        assert seen == n


_WORKER = r'''
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import numpy as np, torch.distributed as dist
from common import SMALL, World, make_params, setup, run
from oracle.oracle import Oracle
from ltrans_b200.host import shard
rank, ws = int(sys.argv[1]), int(sys.argv[2])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + sys.argv[3], rank=rank, world_size=ws)
w = World(**SMALL); n = 240
prm = make_params(w, n, Behavior=0, settlementon=0, mortality=0, VTurbOn=0)
lo, hi = shard.slice_for_rank(n, rank, ws)
o = Oracle()
setup(o, w, prm, n, sl=slice(lo, hi), first_id=shard.first_id(lo))
run(o, w, 1, nint=6)
f = o.fetch()
st = shard.allreduce_stats(o.stats(), dist)
x = shard.gather_output(f["x"], n, dist)
s = shard.gather_output(f["status"], n, dist)
if rank == 0:
    np.savez({out!r}, x=x, status=s, stats=st)
dist.destroy_process_group()
'''


def test_two_rank_gloo_run_equals_single_rank():
    """N > 1 host path on CPU: each rank steps its particle slice (the oracle stands in for
    the GPU engine), statistics are all-reduced, output is all-gathered.  Philox is keyed on
    the GLOBAL particle id, so the sharded run must equal the unsharded one exactly."""
    from oracle.oracle import Oracle
    w = World(**SMALL); n = 240
    prm = make_params(w, n, Behavior=0, settlementon=0, mortality=0, VTurbOn=0)
    o = Oracle()
    setup(o, w, prm, n)
    run(o, w, 1, nint=6)
    f1, st1 = o.fetch(), o.stats()
    with tempfile.TemporaryDirectory() as d:
        out = os.path.join(d, "o.npz")
        script = os.path.join(d, "w.py")
        open(script, "w").write(_WORKER.format(root=ROOT, out=out))
        port = str(29500 + os.getpid() % 2000)
        procs = [subprocess.Popen([sys.executable, script, str(r), "2", port]) for r in range(2)]
        for p in procs:
            assert p.wait(timeout=300) == 0
        got = np.load(out)
        assert np.array_equal(got["x"], f1["x"])
        assert np.array_equal(got["status"], f1["status"])
        assert np.array_equal(got["stats"], st1)
