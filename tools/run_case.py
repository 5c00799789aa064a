"""Run a case directory written by tools/make_reference_case.py (LTRANS.data + ./input) through the
run-loop driver with the CUDA library (or `--engine oracle`), writing ./output_b200/para*.csv in
the reference's format, and optionally compare with the para*.csv of a real LTRANS run.

    python tools/run_case.py CASE_DIR [--engine cuda|oracle] [--compare REFERENCE_OUTPUT_DIR]"""
import argparse, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ltrans_b200  # noqa: F401
from ltrans_b200.host import roms_io, formats
from ltrans_b200.host.driver import Run
from ltrans_b200.host.world import Projection


def run_case(case, engine="cuda", outdir=None):
    nml = formats.read_namelist(os.path.join(case, "LTRANS.data"))
    prm, flat = formats.params_from_namelist(nml)
    j = lambda p: os.path.normpath(os.path.join(case, p))
    proj = Projection(lonmin_nml=float(flat["lonmin"]), latmin_nml=float(flat["latmin"]), pi=float(flat["pi"]), radius=float(flat["earth_radius"]))
    const = {}
    for key, name in (("zeta", "zeta"), ("salt", "salt"), ("temp", "temp"), ("u", "u"), ("v", "v"), ("w", "w"), ("aks", "aks")):
        if not flat.get("read" + name, True):
            const[key] = float(flat.get("const" + name, 0.0))
    w = roms_io.RomsWorld(j(flat["ncgridfile"]), j(flat["prefix"]), flat["suffix"], int(flat["filenum"]), int(flat["numdigits"]),
                          int(flat["tdim"]), bool(flat.get("startfile", False)), proj=proj, dt_hydro=float(flat["dt"]), const=const)
    lon, lat, z, dob, startpoly = formats.read_particles_csv(j(flat["parfile"]), prm.settlementon)
    if engine == "oracle":
        sys.path.insert(0, ROOT)
        from oracle.oracle import Oracle
        e = Oracle()
    else:
        from ltrans_b200.host.binding import LtransLib
        e = LtransLib()
    outdir = outdir or os.path.join(case, "output_b200" if engine == "cuda" else "output_oracle")
    run = Run(e, w, prm, outdir, days=float(flat["days"]), iprint=int(flat["iprint"]), write_csv=bool(flat.get("writecsv", True)),
              write_nc=bool(flat.get("writenc", False)), NCOutFile=flat.get("ncoutfile", "output"), NCtime=int(flat.get("nctime", 0)))
    run.init(lon, lat, z, dob, startpoly=startpoly if prm.settlementon else None)
    run.run(); w.close()
    return outdir


def compare(a, b):
    """para*.csv of two output directories: depth (F10.3), status, lon, lat (F9.4)"""
    names = sorted(n for n in os.listdir(a) if n.startswith("para") and n.endswith(".csv"))
    worst = np.zeros(4)
    for n in names:
        if not os.path.exists(os.path.join(b, n)):
            print("missing in", b, ":", n); continue
        x, y = np.loadtxt(os.path.join(a, n), delimiter=",", ndmin=2), np.loadtxt(os.path.join(b, n), delimiter=",", ndmin=2)
        d = np.abs(x[:, :4] - y[:, :4]).max(axis=0)
        worst = np.maximum(worst, d)
    print("%d files; largest |difference| depth %.3f m, status %d, lon %.4f deg, lat %.4f deg (print resolution 1e-3 m, 1e-4 deg)" % (
        len(names), worst[0], int(worst[1]), worst[2], worst[3]))
    return worst


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("case"); ap.add_argument("--engine", default="cuda", choices=("cuda", "oracle")); ap.add_argument("--compare", default=None)
    a = ap.parse_args()
    out = run_case(a.case, a.engine)
    print("wrote", out)
    if a.compare:
        compare(out, a.compare)
