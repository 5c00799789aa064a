"""ROMS NetCDF on either side of the particle loop: the grid and history files the reference
reads (`initGrid` hydrodynamic_module.f90:95-345, `initHydro` :660-1043, `updateHydro`
:1052-1410) and the particle NetCDF file it writes (`createNetCDF` :2953-3297, `writeNetCDF`
:3300-3434).  In production these stay with the Fortran host; this module is the harness-side
mirror so the Python driver can run from files laid out like
the reference's inputs.  NetCDF-3 classic through `scipy.io.netcdf_file` (no netCDF4 / HDF5
library exists in the image); no scale/offset handling, like the plain NF90_GET_VAR reads."""
import os

import numpy as np
from scipy.io import netcdf_file

from .world import World, Projection


def history_filename(prefix, counter, suffix, numdigits):
    """prefix + I<numdigits>.<numdigits>(counter) + suffix (hydro:270-290, 712-732)."""
    if not 1 <= numdigits <= 8:
        raise ValueError("Model presently does not support numdigits of %d" % numdigits)
    return "%s%0*d%s" % (prefix, numdigits, counter, suffix)


def record_location(k, tdim, startfile):
    """(file offset from filenum, 0-based record in that file) of the k-th record the run
    consumes (k = 0, 1, 2 are initHydro's back / centre / forward).  The first file holds
    tdim + 1 records when `startfile` is set (hydro:1090-1126)."""
    n0 = tdim + (1 if startfile else 0)
    if k < n0:
        return 0, k
    return 1 + (k - n0) // tdim, (k - n0) % tdim


# ---------------------------------------------------------------------------- writers --
def write_grid_nc(path, w):
    """A ROMS grid file with the variables initGrid reads (hydro:181-261)."""
    with netcdf_file(path, "w") as f:
        f.createDimension("eta_rho", w.nj); f.createDimension("xi_rho", w.ni)
        f.createDimension("eta_u", w.nj); f.createDimension("xi_u", w.ni - 1)
        f.createDimension("eta_v", w.nj - 1); f.createDimension("xi_v", w.ni)
        for name, a, dims in (("h", w.h, ("eta_rho", "xi_rho")), ("angle", w.angle, ("eta_rho", "xi_rho")),
                              ("lon_rho", w.lon_r, ("eta_rho", "xi_rho")), ("lat_rho", w.lat_r, ("eta_rho", "xi_rho")),
                              ("lon_u", w.lon_u, ("eta_u", "xi_u")), ("lat_u", w.lat_u, ("eta_u", "xi_u")),
                              ("lon_v", w.lon_v, ("eta_v", "xi_v")), ("lat_v", w.lat_v, ("eta_v", "xi_v")),
                              ("mask_rho", w.mask_rho, ("eta_rho", "xi_rho")), ("mask_u", w.mask_u, ("eta_u", "xi_u")),
                              ("mask_v", w.mask_v, ("eta_v", "xi_v"))):
            v = f.createVariable(name, "d", dims)
            v[:] = np.asarray(a, np.float64)


_HIS = (("zeta", "zeta", ("ocean_time", "eta_rho", "xi_rho")), ("salt", "salt", ("ocean_time", "s_rho", "eta_rho", "xi_rho")),
        ("temp", "temp", ("ocean_time", "s_rho", "eta_rho", "xi_rho")), ("u", "u", ("ocean_time", "s_rho", "eta_u", "xi_u")),
        ("v", "v", ("ocean_time", "s_rho", "eta_v", "xi_v")), ("w", "w", ("ocean_time", "s_w", "eta_rho", "xi_rho")),
        ("AKs", "aks", ("ocean_time", "s_w", "eta_rho", "xi_rho")))


def write_history_nc(w, prefix, suffix, filenum, numdigits, nrec, tdim, startfile=False):
    """History files holding records 0 .. nrec-1 of `w`, tdim per file (tdim + 1 in the first
    when `startfile`), float32 like ROMS output.  Returns the file names."""
    names, k = [], 0
    while k < nrec:
        off, _ = record_location(k, tdim, startfile)
        cnt = min(nrec - k, tdim + (1 if (startfile and off == 0) else 0))
        path = history_filename(prefix, filenum + off, suffix, numdigits)
        with netcdf_file(path, "w") as f:
            f.createDimension("ocean_time", None)
            f.createDimension("s_rho", w.us); f.createDimension("s_w", w.ws)
            f.createDimension("eta_rho", w.nj); f.createDimension("xi_rho", w.ni)
            f.createDimension("eta_u", w.nj); f.createDimension("xi_u", w.ni - 1)
            f.createDimension("eta_v", w.nj - 1); f.createDimension("xi_v", w.ni)
            for name, a, dim in (("s_rho", w.sc_r, "s_rho"), ("Cs_r", w.Cs_r, "s_rho"), ("s_w", w.sc_w, "s_w"), ("Cs_w", w.Cs_w, "s_w")):
                v = f.createVariable(name, "d", (dim,)); v[:] = a
            t = f.createVariable("ocean_time", "d", ("ocean_time",))
            vs = {name: f.createVariable(name, "f", dims) for name, _, dims in _HIS}
            for q in range(cnt):
                rec = w.record(k + q)
                t[q] = (k + q) * w.dt_hydro
                for name, key, _ in _HIS:
                    vs[name][q] = rec[key]
        names.append(path)
        k += cnt
    return names


# ----------------------------------------------------------------------------- reader --
class RomsWorld(World):
    """A `World` whose grid and records come from ROMS NetCDF files, with the reference's file
    sequencing (prefix / filenum / numdigits / suffix, tdim records per file, startfile) and
    its read / const switches (readZeta, constZeta, ... hydro:757-985)."""

    def __init__(self, gridfile, prefix, suffix, filenum, numdigits, tdim, startfile=False,
                 proj=None, dt_hydro=3600.0, const=None):
        self.proj = proj or Projection()
        self.prefix, self.suffix, self.filenum, self.numdigits = prefix, suffix, filenum, numdigits
        self.tdim, self.startfile, self.dt_hydro = tdim, bool(startfile), dt_hydro
        self.const = dict(const or {})          # e.g. {"zeta": 0.0}: not read, constant (readZeta = .FALSE.)
        self.uniform = None
        with netcdf_file(gridfile, "r", mmap=False, maskandscale=False) as f:
            g = {k: np.array(f.variables[k][:], np.float64) for k in
                 ("h", "angle", "lon_rho", "lat_rho", "lon_u", "lat_u", "lon_v", "lat_v", "mask_rho", "mask_u", "mask_v")}
        self.nj, self.ni = g["h"].shape
        self.h, self.angle = g["h"], g["angle"]
        self.lon_r, self.lat_r, self.lon_u, self.lat_u = g["lon_rho"], g["lat_rho"], g["lon_u"], g["lat_u"]
        self.lon_v, self.lat_v = g["lon_v"], g["lat_v"]
        self.mask_rho = g["mask_rho"].astype(np.int32)          # hydro keeps the file's masks (:346-348)
        self.mask_u, self.mask_v = g["mask_u"].astype(np.int32), g["mask_v"].astype(np.int32)
        self.mask_bnd = self._erode(self.mask_rho)              # createBounds erodes its own copy (boundary:166-196)
        P = self.proj
        self.x_r, self.y_r = P.lon2x(self.lon_r, self.lat_r), P.lat2y(self.lat_r)      # hydro:357-379
        self.x_u, self.y_u = P.lon2x(self.lon_u, self.lat_u), P.lat2y(self.lat_u)
        self.x_v, self.y_v = P.lon2x(self.lon_v, self.lat_v), P.lat2y(self.lat_v)
        first = history_filename(prefix, filenum, suffix, numdigits)
        with netcdf_file(first, "r", mmap=False, maskandscale=False) as f:      # s-levels (hydro:295-340)
            def var(a, b):
                return np.array(f.variables[a if a in f.variables else b][:], np.float64)
            self.sc_r, self.Cs_r = var("s_rho", "sc_r"), var("Cs_r", "Cs_r")
            self.sc_w, self.Cs_w = var("s_w", "sc_w"), var("Cs_w", "Cs_w")
        self.us, self.ws = len(self.sc_r), len(self.sc_w)
        self._grid = None
        self._open = (None, None)

    def _file(self, off):
        if self._open[0] != off:
            if self._open[1] is not None:
                self._open[1].close()
            path = history_filename(self.prefix, self.filenum + off, self.suffix, self.numdigits)
            self._open = (off, netcdf_file(path, "r", mmap=False, maskandscale=False))
        return self._open[1]

    def record(self, r, dtype=np.float32):
        off, q = record_location(r, self.tdim, self.startfile)
        f = self._file(off)
        shapes = dict(zeta=(self.nj, self.ni), salt=(self.us, self.nj, self.ni), temp=(self.us, self.nj, self.ni),
                      u=(self.us, self.nj, self.ni - 1), v=(self.us, self.nj - 1, self.ni),
                      w=(self.ws, self.nj, self.ni), aks=(self.ws, self.nj, self.ni))
        out = {}
        for name, key, _ in _HIS:
            if key in self.const:
                out[key] = np.full(shapes[key], self.const[key], dtype)
            else:
                out[key] = np.ascontiguousarray(f.variables[name][q], dtype=dtype)
        return out

    def close(self):
        if self._open[1] is not None:
            self._open[1].close()
        self._open = (None, None)


# -------------------------------------------------------------------- particle output --
class ParticleNetCDF:
    """The reference's particle NetCDF output: one file `<NCOutFile>.nc`, or numbered files
    `<NCOutFile>_NNN.nc` started every `NCtime` seconds of model time (hydro:2984-3001,
    3316-3340).  Variables are (time, numpar) doubles: age, lon, lat, depth, color [, hitBottom,
    hitLand, salinity, temperature], model_time(time), and dob(numpar) in the first file."""

    def __init__(self, outpath, NCOutFile, numpar, NCtime=0, SaltTempOn=False, TrackCollisions=False, attrs=None):
        self.outpath, self.name, self.numpar, self.NCtime = outpath, NCOutFile, numpar, NCtime
        self.salt, self.hits, self.attrs = bool(SaltTempOn), bool(TrackCollisions), dict(attrs or {})
        self.NCcount, self.NCstart, self.prcount, self.f = 0, 0, 0, None

    def _path(self):
        if self.NCtime == 0:
            return os.path.join(self.outpath, self.name + ".nc")
        return os.path.join(self.outpath, "%s_%03d.nc" % (self.name, self.NCcount))

    def create(self, dob=None):
        if self.f is not None:
            self.f.close()
        self.prcount = 0
        if self.NCtime != 0:
            self.NCcount += 1
        f = self.f = netcdf_file(self._path(), "w")
        f.createDimension("time", None); f.createDimension("numpar", self.numpar)
        def var(name, dims, long_name, units, field):
            v = f.createVariable(name, "d", dims)
            v.long_name, v.field = long_name, field
            if units is not None:
                v.units = units
            return v
        var("model_time", ("time",), "time that has passed thus far in the model", "seconds", "model_time, scalar, series")
        if dob is not None:
            v = var("dob", ("numpar",), "Date of Birth of particles in seconds from model start", "seconds", "age, scalar, series")
            v[:] = np.asarray(dob, np.float64)
        var("age", ("time", "numpar"), "age of particles", "seconds", "age, scalar, series")
        var("lon", ("time", "numpar"), "longitude of particles", "decimal degrees E", "lon, scalar, series")
        var("lat", ("time", "numpar"), "latitude of particles", "decimal degrees N", "lat, scalar, series")
        var("depth", ("time", "numpar"), "depth of particles", "meters below surface", "depth, scalar, series")
        var("color", ("time", "numpar"), "identification number for particle behavior or status",
            "nondimensional, see LTRANS User Guide", "color, scalar, series")
        if self.hits:
            var("hitBottom", ("time", "numpar"), "# of times Particle Collided with Bottom", "Number of Collisions",
                "hitBottom, scalar, series")
            var("hitLand", ("time", "numpar"), "# of times Particle Collided with Land", "Number of Collisions",
                "hitLand, scalar, series")
        if self.salt:
            var("salinity", ("time", "numpar"), "Salinity at the particle's location", None, "salinity, scalar, series")
            var("temperature", ("time", "numpar"), "Temperature at the particle's location", "dg Celsius",
                "temperature, scalar, series")
        f.type = "Position and characteristics of particles"
        f.title = "LTRANS output"
        for k, v in self.attrs.items():
            setattr(f, k, v)

    def write(self, time, age, lon, lat, depth, color, hitB=None, hitL=None, salt=None, temp=None):
        if self.NCtime != 0 and time - self.NCstart >= self.NCtime:            # :3329-3333
            self.NCstart = time
            self.create()
        q, v = self.prcount, self.f.variables
        self.prcount += 1
        v["model_time"][q] = float(time)
        for name, a in (("age", age), ("lon", lon), ("lat", lat), ("depth", depth), ("color", color)):
            v[name][q] = np.asarray(a, np.float64)
        if self.hits:
            v["hitBottom"][q] = np.asarray(hitB, np.float64); v["hitLand"][q] = np.asarray(hitL, np.float64)
        if self.salt:
            v["salinity"][q] = np.asarray(salt, np.float64); v["temperature"][q] = np.asarray(temp, np.float64)
        self.f.flush()

    def close(self):
        if self.f is not None:
            self.f.close()
            self.f = None
