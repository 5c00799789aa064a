"""NetCDF side of the harness: grid / history files in the
reference's layout and file sequencing, and the particle NetCDF output."""
import os

import numpy as np
import pytest
from scipy.io import netcdf_file

from common import SMALL, World
from ltrans_b200.host import roms_io


def test_history_filename_and_sequence():
    assert roms_io.history_filename("his_", 7, ".nc", 4) == "his_0007.nc"          # hydro:270-290
    with pytest.raises(ValueError):
        roms_io.history_filename("his_", 7, ".nc", 9)
    # tdim = 4: files hold records 0-3, 4-7, ...; with startfile the first holds 0-4 (hydro:1090-1126)
    assert [roms_io.record_location(k, 4, False) for k in (0, 3, 4, 7, 8)] == [(0, 0), (0, 3), (1, 0), (1, 3), (2, 0)]
    assert [roms_io.record_location(k, 4, True) for k in (0, 4, 5, 8, 9)] == [(0, 0), (0, 4), (1, 0), (1, 3), (2, 0)]
    # the same walk as updateHydro's stepf / iint bookkeeping
    for startfile in (False, True):
        stepf, iint, tdim = 3, 0, 4
        for k in range(3, 20):
            if (startfile and iint == 0 and stepf == tdim) or stepf < tdim:
                stepf += 1
            else:
                iint += 1; stepf = 1
            assert roms_io.record_location(k, tdim, startfile) == (iint, stepf - 1)


@pytest.mark.parametrize("startfile", [False, True])
def test_grid_and_records_round_trip(tmp_path, startfile):
    w = World(**SMALL)
    gridfile = str(tmp_path / "grid.nc")
    prefix = str(tmp_path / "his_")
    roms_io.write_grid_nc(gridfile, w)
    names = roms_io.write_history_nc(w, prefix, ".nc", 3, 4, nrec=9, tdim=4, startfile=startfile)
    assert [os.path.basename(n) for n in names] == ["his_0003.nc", "his_0004.nc"] + ([] if startfile else ["his_0005.nc"])
    r = roms_io.RomsWorld(gridfile, prefix, ".nc", 3, 4, tdim=4, startfile=startfile)
    assert (r.ni, r.nj, r.us, r.ws) == (w.ni, w.nj, w.us, w.ws)
    g0, g1 = w.grid(), r.grid()
    for k, a in g0.items():
        assert np.array_equal(np.asarray(a), np.asarray(g1[k])), k
    b0, b1 = w.bounds(), r.bounds()
    for k, a in b0.items():
        assert np.array_equal(np.asarray(a), np.asarray(b1[k])), k
    for k in (0, 3, 4, 5, 8):
        a, b = w.record(k), r.record(k)
        for name in a:
            assert a[name].dtype == b[name].dtype and np.array_equal(a[name], b[name]), (k, name)
    rc = roms_io.RomsWorld(gridfile, prefix, ".nc", 3, 4, tdim=4, startfile=startfile, const={"zeta": 0.25, "aks": 1e-3})
    rec = rc.record(1)
    assert np.all(rec["zeta"] == np.float32(0.25)) and np.all(rec["aks"] == np.float32(1e-3))      # readZeta = .FALSE.
    assert np.array_equal(rec["u"], w.record(1)["u"])
    r.close(); rc.close()


def test_eroded_mask_is_boundary_only(tmp_path):
    """hydro keeps the file's mask; createBounds erodes its own copy (boundary:166-196)."""
    w = World(**SMALL)
    w.mask_rho = w.mask_rho.copy()
    j, i = 1, w.ni // 2                      # a one-node inlet into the southern land rim: < 2 water neighbours
    assert w.mask_rho[j, i] == 0 and w.mask_rho[j + 1, i] == 1
    w.mask_rho[j, i] = 1
    w.mask_u = w.mask_rho[:, :-1] * w.mask_rho[:, 1:]; w.mask_v = w.mask_rho[:-1, :] * w.mask_rho[1:, :]
    roms_io.write_grid_nc(str(tmp_path / "g.nc"), w)
    roms_io.write_history_nc(w, str(tmp_path / "h"), ".nc", 1, 1, nrec=1, tdim=1)
    r = roms_io.RomsWorld(str(tmp_path / "g.nc"), str(tmp_path / "h"), ".nc", 1, 1, tdim=1)
    assert r.mask_rho[j, i] == 1 and r.mask_bnd[j, i] == 0
    ref = World(**SMALL).bounds()
    got = r.bounds()
    assert np.array_equal(ref["bnd_x"], got["bnd_x"]) and np.array_equal(ref["land"], got["land"])
    r.close()


def test_particle_netcdf(tmp_path):
    n = 7
    out = roms_io.ParticleNetCDF(str(tmp_path), "output", n, NCtime=0, SaltTempOn=True, TrackCollisions=True)
    dob = np.arange(n) * 120.0
    out.create(dob)
    for k, t in enumerate((0, 3600, 7200)):
        a = np.full(n, float(k))
        out.write(t, a, a + 1, a + 2, -a, a * 0 + 1, np.arange(n), np.arange(n) * 2, a + 20, a + 10)
    out.close()
    with netcdf_file(str(tmp_path / "output.nc"), "r", mmap=False) as f:
        assert f.dimensions["numpar"] == n and f.dimensions["time"] is None
        assert np.array_equal(f.variables["model_time"][:], [0.0, 3600.0, 7200.0])
        assert np.array_equal(f.variables["dob"][:], dob)
        assert f.variables["lon"][:].shape == (3, n) and np.all(f.variables["lon"][2] == 3.0)
        assert np.array_equal(f.variables["hitLand"][1], np.arange(n) * 2.0)
        assert f.variables["depth"].units == b"meters below surface" and f.title == b"LTRANS output"
        assert set(f.variables) == {"model_time", "dob", "age", "lon", "lat", "depth", "color", "hitBottom", "hitLand",
                                    "salinity", "temperature"}
    # numbered files every NCtime seconds of model time (hydro:2984-3001, 3329-3333): dob only in the first
    seq = roms_io.ParticleNetCDF(str(tmp_path), "seq", n, NCtime=7200)
    seq.create(dob)
    for t in (0, 3600, 7200, 10800, 14400):
        z = np.zeros(n)
        seq.write(t, z, z, z, z, z)
    seq.close()
    files = sorted(p for p in os.listdir(tmp_path) if p.startswith("seq_"))
    assert files == ["seq_001.nc", "seq_002.nc", "seq_003.nc"]
    with netcdf_file(str(tmp_path / "seq_002.nc"), "r", mmap=False) as f:
        assert np.array_equal(f.variables["model_time"][:], [7200.0, 10800.0]) and "dob" not in f.variables
