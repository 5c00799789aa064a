"""Host-side formats and run loop: the namelist
reader, CSV inputs, Fortran-formatted outputs and the external/internal loop with print
intervals.  The CPU oracle stands in for the engine here; the GPU counterpart is in
tests/test_parity_gpu.py::test_driver_outputs_match."""
import os
import re

import numpy as np

from common import ROOT, SMALL, World, make_params
from ltrans_b200.host import formats
from ltrans_b200.host.driver import Run


def test_namelist_reader_matches_shipped_defaults():
    nml = formats.read_namelist(os.path.join(ROOT, "tests", "golden", "LTRANS_sample.data"))
    assert list(nml)[:4] == ["numparticles", "timeparam", "hydroparam", "turbparam"]
    assert nml["timeparam"] == {"days": 4.5, "iprint": 3600, "dt": 3600, "idt": 120}
    assert nml["output"]["outpath"] == "./output/" and nml["parloc"]["parfile"].endswith(".csv")
    p, flat = formats.params_from_namelist(nml)
    from ltrans_b200.host.binding import Params
    q = Params.shipped()
    for name, _ in Params._fields_:
        assert getattr(p, name) == getattr(q, name), name
    assert flat["pi"] == 3.14159265358979 and flat["latmin"] == 36


def test_namelist_reader_crlf_and_ampersand_groups(tmp_path):
    f = tmp_path / "x.data"
    f.write_bytes(b"&timeparam\r\n  dt = 1800 ! c\r\n  days=2.d0\r\n/\r\n$other\r\n FreeSlip=.T.\r\n$end\r\n")
    nml = formats.read_namelist(str(f))
    assert nml == {"timeparam": {"dt": 1800, "days": 2.0}, "other": {"freeslip": True}}


def test_fortran_edit_descriptors():
    assert formats._F(-12.3456, 10, 3) == "   -12.346" and formats._F(0.5, 9, 4) == "   0.5000"
    assert formats._I(-3, 7) == "     -3" and formats._I(12345678, 7) == "*******"
    assert formats._F(123456.0, 8, 4) == "********"
    assert formats.para_filename(2, "out") == os.path.join("out", "para10000002.csv")


def test_particle_csv_round_trip(tmp_path):
    lon = np.array([-76.01, -75.9]); lat = np.array([37.0, 37.2]); z = np.array([-3.5, -0.25]); dob = np.array([0.0, 3600.0])
    p = str(tmp_path / "Initial_particle_locations.csv")
    formats.write_particles_csv(p, lon, lat, z, dob, np.array([101001, 101002]))
    a = formats.read_particles_csv(p, True)
    assert np.allclose(a[0], lon) and np.allclose(a[2], z) and list(a[4]) == [101001, 101002]
    assert list(formats.read_particles_csv(p, False)[4]) == [0, 0]


def run_driver(engine, outdir, n=120, world=None, run_kw=None, **kw):
    w = world or World(**SMALL)
    base = dict(Behavior=4, HTurbOn=1, VTurbOn=0, pediage=3600.0, deadage=9000.0, SaltTempOn=1,
                TrackCollisions=1, ErrorFlag=1)
    base.update(kw)
    prm = make_params(w, n, **base)
    x, y, z, dob, r, u, v = w.seed_particles(n, seed=77)
    lon, lat = w.proj.x2lon(x, y), w.proj.y2lat(y)
    run = Run(engine, w, prm, outdir, days=3 / 24.0, iprint=1800, **(run_kw or {}))
    run.init(lon, lat, z, dob, startpoly=np.full(n, 101001, np.int32))
    f = run.run()
    return run, f


def test_run_loop_writes_reference_formats(tmp_path):
    from oracle.oracle import Oracle
    out = str(tmp_path / "o")
    run, f = run_driver(Oracle(), out)
    files = sorted(os.listdir(out))
    # 3 h at iprint = 1800 s -> 6 prints, first file para10000002.csv
    assert [x for x in files if x.startswith("para")] == ["para1000000%d.csv" % k for k in range(2, 8)]
    row = re.compile(r"^ *-?\d+\.\d{3}, *-?\d+, *-?\d+\.\d{4}, *-?\d+\.\d{4}, *-?\d+\.\d{4}, *-?\d+\.\d{4}$")
    lines = open(os.path.join(out, "para10000007.csv")).read().splitlines()
    assert len(lines) == 120 and all(row.match(s) and len(s) == 10 + 8 + 10 + 10 + 9 + 9 for s in lines)
    end = open(os.path.join(out, "endfile.csv")).read().splitlines()
    assert len(end) == 120 and all(len(s.split(",")) == 6 for s in end)
    st = np.array([int(s.split(",")[2]) for s in end])
    assert set(st) <= {4, 2, -1, -2, -3} and (st == -2).sum() == (f["status"] == -2).sum()
    hits = open(os.path.join(out, "LandHits.csv")).read().splitlines()
    assert hits[0] == " numpar,lon,lat,depth,age,time,hitLand" and all(len(s.split(",")) == 7 for s in hits[1:])


def test_run_loop_from_netcdf_files(tmp_path):
    """The same run fed from ROMS grid / history NetCDF files (file sequencing of updateHydro) and
    writing the particle NetCDF output next to the CSVs: identical CSVs, NetCDF rows = CSV rows."""
    from oracle.oracle import Oracle
    from scipy.io import netcdf_file
    from ltrans_b200.host import roms_io
    w = World(**SMALL)
    roms_io.write_grid_nc(str(tmp_path / "grid.nc"), w)
    roms_io.write_history_nc(w, str(tmp_path / "his_"), ".nc", 1, 4, nrec=5, tdim=2)
    rw = roms_io.RomsWorld(str(tmp_path / "grid.nc"), str(tmp_path / "his_"), ".nc", 1, 4, tdim=2)
    a, b = str(tmp_path / "a"), str(tmp_path / "b")
    run_driver(Oracle(), a, n=60)
    run_driver(Oracle(), b, n=60, world=rw, run_kw=dict(write_nc=True, NCOutFile="out"))
    rw.close()
    for name in sorted(os.listdir(a)):
        assert open(os.path.join(a, name)).read() == open(os.path.join(b, name)).read(), name
    with netcdf_file(os.path.join(b, "out.nc"), "r", mmap=False) as f:
        t = f.variables["model_time"][:]
        assert np.array_equal(t, np.arange(0, 3 * 3600 + 1, 1800.0))          # t = 0 plus 6 prints
        assert np.all(f.variables["color"][0] == 4.0) and np.all(f.variables["age"][0] == 0.0)
        last = np.loadtxt(os.path.join(b, "para10000007.csv"), delimiter=",")
        assert np.allclose(f.variables["depth"][-1], last[:, 0], rtol=0, atol=5.01e-4)
        assert np.array_equal(f.variables["color"][-1], last[:, 1])
        assert np.allclose(f.variables["lon"][-1], last[:, 2], rtol=0, atol=5.01e-5)
        assert np.allclose(f.variables["salinity"][-1], last[:, 4], rtol=0, atol=5.01e-5)


def test_lonlat_conversion_matches_projection():
    """ora_fetch_lonlat / ltgpu_fetch_lonlat = x2lon, y2lat of conversion_module.f90:322-378"""
    from oracle.oracle import Oracle
    w = World(**SMALL); n = 500
    prm = make_params(w, n, Behavior=0, settlementon=0)
    o = Oracle(); o.create(prm); o.set_grid(w.grid()); o.set_bounds(w.bounds())
    x, y, z, dob, r, u, v = w.seed_particles(n, seed=3)
    o.set_particles(x, y, z, dob, None, r, u, v)
    lon, lat = o.fetch_lonlat(w.proj)
    assert np.allclose(lon, w.proj.x2lon(x, y), rtol=0, atol=1e-12) and np.allclose(lat, w.proj.y2lat(y), rtol=0, atol=1e-12)
    assert np.allclose(w.proj.lon2x(lon, lat), x, rtol=1e-12) and np.allclose(w.proj.lat2y(lat), y, rtol=1e-12)
    lon_m, lat_m = o.fetch_lonlat(w.proj, spherical=False)       # Mercator branch
    R, pi = w.proj.R, w.proj.pi
    assert np.allclose(lon_m, x / R * 180.0 / pi, rtol=1e-14) and np.allclose(lat_m, 2 * 180.0 / pi * (np.arctan(np.exp(y / R)) - pi / 4), rtol=1e-12, atol=1e-12)


def test_reference_case_round_trip(tmp_path):
    """tools/make_reference_case.py writes a full LTRANS v.2b input set (LTRANS.data, grid and history
    NetCDF, particle CSV); tools/run_case.py runs it from those files.  Same para files as a run fed
    from memory with the positions read back from the CSV."""
    import importlib.util
    from oracle.oracle import Oracle
    def load(name):
        spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "tools", name + ".py"))
        m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m); return m
    mk, rc = load("make_reference_case"), load("run_case")
    case = str(tmp_path / "case")
    w = mk.make_case(case, n=80, days=0.125, small=True, tdim=3)
    nml = formats.read_namelist(os.path.join(case, "LTRANS.data"))
    prm, flat = formats.params_from_namelist(nml)
    assert (prm.numpar, prm.us, prm.ws, prm.HTurbOn, prm.VTurbOn, prm.Behavior, prm.ErrorFlag) == (80, w.us, w.ws, 0, 0, 0, 1)
    assert flat["readdens"] is False and flat["ncgridfile"] == "./input/grid.nc" and flat["tdim"] == 3
    assert sorted(os.listdir(os.path.join(case, "input"))) == ["Initial_particle_locations.csv", "grid.nc", "his_0001.nc", "his_0002.nc"]
    out = rc.run_case(case, engine="oracle")
    lon, lat, z, dob, _ = formats.read_particles_csv(os.path.join(case, "input", "Initial_particle_locations.csv"), False)
    mem = str(tmp_path / "mem")
    run = Run(Oracle(), w, prm, mem, days=0.125, iprint=3600)
    run.init(lon, lat, z, dob); run.run()
    names = sorted(n for n in os.listdir(mem) if n.startswith("para"))
    assert names == ["para1000000%d.csv" % k for k in (2, 3, 4)]
    for n in names + ["endfile.csv"]:
        assert open(os.path.join(mem, n)).read() == open(os.path.join(out, n)).read(), n
    assert np.all(rc.compare(out, mem) == 0)
