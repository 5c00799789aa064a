"""Write a complete LTRANS v.2b input set for the synthetic world, so that anyone with gfortran +
netcdf-fortran can run the REAL reference on exactly the inputs the CUDA library and the oracle
are tested on (mitigation 4: the oracle is "parity unpinned" here because no
Fortran compiler exists in the build image).  Nothing of the reference is copied: the files are
generated.

    python tools/make_reference_case.py CASE_DIR [--particles 5000] [--days 2] [--small]
    cd CASE_DIR && /path/to/LTRANS.exe            # reads ./LTRANS.data, writes ./output/para*.csv
    python tools/run_case.py CASE_DIR --compare CASE_DIR/output

Default = BASELINE configs[0]: 130x130x20 grid, passive particles, RK4 advection only (HTurb and
VTurb off: the reference draws from MT19937, the library from Philox, so only turbulence-free
runs are comparable particle by particle).  Build the reference with -O2 -fno-fast-math."""
import argparse, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ltrans_b200  # noqa: F401
from ltrans_b200.host import roms_io, formats
from ltrans_b200.host.world import World

SMALL = dict(ni=40, nj=36, us=10)


def fbool(v):
    return ".TRUE." if v else ".FALSE."


def write_namelist(path, w, n, days, tdim, habitat=None, **over):
    p = dict(numpar=n, days=days, iprint=3600, dt=3600, idt=120, us=w.us, ws=w.ws, tdim=tdim, hc=0.2, z0=0.0005,
             Vtransform=1, HTurbOn=False, VTurbOn=False, ConstantHTurb=1.0, Behavior=0, OpenOceanBoundary=True,
             mortality=False, deadage=367200, pediage=302400, swimstart=0.0, swimslow=0.005, swimfast=0.005,
             Sgradient=1.0, sink=-0.0003, Hswimspeed=0.9, Swimdepth=2, settlementon=habitat is not None,
             holesExist=habitat is not None and habitat["hedges"] > 0, seed=9, ErrorFlag=1, SaltTempOn=False,
             TrackCollisions=False, FreeSlip=False)
    p.update(over)
    hab = habitat or dict(poly_id=[0], hole_id=[0], pedges=0, hedges=0)
    pid, hid = list(hab["poly_id"]) or [0], list(hab["hole_id"]) or [0]
    L = []
    def grp(name, items):
        L.append("$" + name)
        for k, v in items:
            L.append("  %s = %s" % (k, fbool(v) if isinstance(v, bool) else ("'%s'" % v if isinstance(v, str) else repr(v))))
        L.append("$end")
    grp("numparticles", [("numpar", p["numpar"])])
    grp("timeparam", [(k, p[k]) for k in ("days", "iprint", "dt", "idt")])
    grp("hydroparam", [(k, p[k]) for k in ("us", "ws", "tdim", "hc", "z0", "Vtransform")] +
        [(a + b, v) for b in ("Zeta", "Salt", "Temp", "U", "V", "W", "Aks") for a, v in (("read", True), ("const", 0.0))] +
        [("readDens", False), ("constDens", 0.0)])           # no `rho` variable in the synthetic history files
    grp("turbparam", [(k, p[k]) for k in ("HTurbOn", "VTurbOn", "ConstantHTurb")])
    grp("behavparam", [(k, p[k]) for k in ("Behavior", "OpenOceanBoundary", "mortality", "deadage", "pediage", "swimstart", "swimslow",
                                           "swimfast", "Sgradient", "sink", "Hswimspeed", "Swimdepth")])
    grp("dvmparam", [("twistart", 4.801821), ("twiend", 19.19956), ("daylength", 14.39774), ("Em", 1814.328), ("Kd", 1.07), ("thresh", 0.0166)])
    grp("settleparam", [("settlementon", p["settlementon"]), ("holesExist", p["holesExist"]), ("minpolyid", int(min(pid))),
                        ("maxpolyid", int(max(pid))), ("minholeid", int(min(hid))), ("maxholeid", int(max(hid))),
                        ("pedges", int(hab["pedges"])), ("hedges", int(hab["hedges"]))])
    grp("convparam", [("PI", 3.14159265358979), ("Earth_Radius", 6378000), ("SphericalProjection", True),
                      ("latmin", int(w.proj.latmin + 1)), ("lonmin", int(w.proj.lonmin + 1))])
    grp("romsgrid", [("NCgridfile", "./input/grid.nc")])
    grp("romsoutput", [("prefix", "./input/his_"), ("suffix", ".nc"), ("filenum", 1), ("numdigits", 4), ("startfile", False)])
    grp("parloc", [("parfile", "./input/Initial_particle_locations.csv")])
    grp("habpolyloc", [("habitatfile", "./input/End_polygons.csv"), ("holefile", "./input/End_holes.csv")])
    grp("output", [("outpath", "./output/"), ("NCOutFile", "output"), ("outpathGiven", True), ("writeCSV", True), ("writeNC", False),
                   ("NCtime", 0), ("Github", "n/a"), ("RunName", "synthetic world"), ("ExeDir", "."), ("OutDir", "./output"),
                   ("RunBy", "n/a"), ("Institution", "n/a"), ("StartedOn", "n/a")])
    grp("other", [("seed", p["seed"]), ("ErrorFlag", p["ErrorFlag"]), ("BoundaryBLNs", False), ("SaltTempOn", p["SaltTempOn"]),
                  ("TrackCollisions", p["TrackCollisions"]), ("WriteHeaders", False), ("WriteModelTiming", False), ("ijbuff", 4),
                  ("FreeSlip", p["FreeSlip"])])
    with open(path, "w", newline="") as f:
        f.write("\r\n".join(L) + "\r\n")                      # the shipped file has CRLF line ends


def make_case(case, n=5000, days=2.0, small=False, tdim=12, seed=1234, **over):
    w = World(**SMALL) if small else World()
    os.makedirs(os.path.join(case, "input"), exist_ok=True); os.makedirs(os.path.join(case, "output"), exist_ok=True)
    nrec = int(days * 86400 / 3600) + 2                       # records 0 .. stepT + 1 (updateHydro reads record p at step p)
    roms_io.write_grid_nc(os.path.join(case, "input", "grid.nc"), w)
    roms_io.write_history_nc(w, os.path.join(case, "input", "his_"), ".nc", 1, 4, nrec=nrec, tdim=tdim)
    x, y, z, dob, r, u, v = w.seed_particles(n, seed=seed)
    lon, lat = w.proj.x2lon(x, y), w.proj.y2lat(y)
    formats.write_particles_csv(os.path.join(case, "input", "Initial_particle_locations.csv"), lon, lat, z, dob)
    write_namelist(os.path.join(case, "LTRANS.data"), w, n, days, tdim, **over)
    return w


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("case"); ap.add_argument("--particles", type=int, default=5000); ap.add_argument("--days", type=float, default=2.0)
    ap.add_argument("--small", action="store_true")
    a = ap.parse_args()
    make_case(a.case, a.particles, a.days, a.small)
    print("wrote", a.case, sorted(os.listdir(os.path.join(a.case, "input")))[:6], "...")
