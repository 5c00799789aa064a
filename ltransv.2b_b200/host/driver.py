"""Host-side mirror of the reference's run loop (LTRANS.f90:141-167 run_LTRANS,
:548-614 external / internal time steps, :1617-1779 printOutput / writeOutput,
:618-658 fin_LTRANS) around the C ABI.  In production this loop is the Fortran host's
(INTEGRATION.md); here it lets the tests and examples run the whole pipeline -- namelist in,
`para*.csv` / `endfile.csv` / hit files / ErrorLog out -- on the synthetic ROMS world.

`engine` is an `LtransLib` (the CUDA library).  The tests also pass the CPU oracle, which has
the same call surface, to check the files this loop writes."""
import os

import numpy as np

from . import formats


class Run:
    def __init__(self, engine, world, prm, outdir, days, iprint, write_csv=True, write_nc=False, NCOutFile="output",
                 NCtime=0):
        self.e, self.w, self.prm, self.outdir = engine, world, prm, outdir
        self.write_nc, self.NCOutFile, self.NCtime, self.nc = write_nc, NCOutFile, NCtime, None
        self.stepT = int(int(days * 86400) / prm.dt)            # LTRANS.f90:156-157
        self.iprint, self.write_csv = iprint, write_csv
        self.prcount, self.printdt = 1, 0                       # :290, :240
        self.startpoly = None
        os.makedirs(outdir, exist_ok=True)

    # ini_LTRANS (:169-544): uploads
    def init(self, lon, lat, z, dob, startpoly=None, habitat=None, r_ele=None, u_ele=None, v_ele=None, first_id=1):
        P = self.w.proj
        x, y = P.lon2x(lon, lat), P.lat2y(lat)                  # :260-261
        self.e.create(self.prm)
        self.e.set_grid(self.w.grid())
        self.e.set_bounds(self.w.bounds())
        if self.prm.settlementon:
            self.e.set_habitat(habitat if habitat is not None else self.w.habitat())
        self.e.set_particles(x, y, z, dob, startpoly, r_ele, u_ele, v_ele, first_id=first_id)
        self.startpoly = startpoly
        self.first_id = first_id
        head = {1: "The following particles were returned to their previous locations:", 2: "The following particles were killed:",
                3: "The following particles were set out of bounds:"}.get(self.prm.ErrorFlag)      # :337-354
        if head:
            with open(os.path.join(self.outdir, "ErrorLog.txt"), "w") as f:
                f.write(" " + head + "\n  \n")
        rc, self.screened, bad = self.e.screen_initial()        # :356-452 start-up screen, on the device
        self._errors()
        if rc:
            raise RuntimeError(f"particle {bad}: bad initial location (ErrorFlag outside 1..3: the reference STOPs)")
        for k in range(3):                                      # initHydro: back, centre, forward
            self.e.push_hydro(self.w.record(k))
        if self.write_nc:                                       # :468-476 createNetCDF(par(:,pDOB)), first print at t = 0
            from .roms_io import ParticleNetCDF
            self.nc = ParticleNetCDF(self.outdir, self.NCOutFile, len(x), self.NCtime, self.prm.SaltTempOn,
                                     self.prm.TrackCollisions)
            self.nc.create(dob)
            zero = np.zeros(len(x))                             # :292-309 initial state at model time 0
            self.nc.write(0, zero, lon, lat, z, np.full(len(x), float(self.prm.Behavior)), zero, zero, zero, zero)
        if self.prm.TrackCollisions:                            # :461-466 header lines
            for name, col in (("LandHits.csv", "hitLand"), ("BottomHits.csv", "hitBottom")):
                with open(os.path.join(self.outdir, name), "w") as f:
                    f.write(" numpar,lon,lat,depth,age,time,%s\n" % col)       # list-directed: leading blank

    def run(self):
        stepIT = self.prm.dt // self.prm.idt
        for p in range(1, self.stepT + 1):                      # :159-161
            if p > 2:                                           # :559 updateHydro
                self.e.push_hydro(self.w.record(p))
                self.e.rotate_hydro()
            for it in range(1, stepIT + 1):                     # :573-577
                rc = self.e.step(p, it)
                if rc:
                    break
                self.printdt += self.prm.idt                    # :602-612
                if self.printdt >= self.iprint:
                    self.print_output((p - 1) * self.prm.dt + it * self.prm.idt)
                    self.printdt = 0
            rc, bad = self.e.sync()
            if rc:                                              # ErrorFlag = 0: the reference STOPs
                self._errors()
                raise RuntimeError(f"particle {bad} hit a STOP condition (ErrorFlag outside 1..3)")
        self._errors()
        return self.finish()

    def _errors(self):
        ev = self.e.drain_events(1 << 16, everything=True)
        if ev:
            formats.append_errorlog(os.path.join(self.outdir, "ErrorLog.txt"), ev)

    # printOutput / writeOutput (:1617-1779)
    def print_output(self, ix3):
        self.prcount += 1
        f = self.e.fetch(("z", "age", "status", "salt", "temp", "hitBottom", "hitLand"))
        lon, lat = self.e.fetch_lonlat(self.w.proj)             # :1712-1716 x2lon / y2lat, on the device
        if self.write_csv:
            st = self.prm.SaltTempOn
            formats.write_para_csv(formats.para_filename(self.prcount, self.outdir), f["z"], f["status"], lon, lat,
                                   f["salt"] if st else None, f["temp"] if st else None)
        if self.nc is not None:                                 # :1754-1775 writeNetCDF(int(ix(3)), ...)
            self.nc.write(int(ix3), f["age"], lon, lat, f["z"], f["status"].astype(np.float64), f["hitBottom"], f["hitLand"],
                          f["salt"], f["temp"])
        if self.prm.TrackCollisions:
            ids = self.first_id + np.arange(len(lon))
            formats.append_hits(os.path.join(self.outdir, "LandHits.csv"), ids, lon, lat, f["z"], f["age"], ix3, f["hitLand"])
            formats.append_hits(os.path.join(self.outdir, "BottomHits.csv"), ids, lon, lat, f["z"], f["age"], ix3, f["hitBottom"])
            self.e.reset_hits()                                 # :1662-1665
        self._errors()

    # fin_LTRANS (:618-658)
    def finish(self):
        f = self.e.fetch(("x", "y", "status", "endpoly", "lifespan"))
        lon, lat = self.e.fetch_lonlat(self.w.proj)
        if self.write_csv:
            sp = self.startpoly if self.prm.settlementon else None
            if self.prm.settlementon and sp is None:
                sp = np.zeros(len(lon), np.int32)
            formats.write_endfile(os.path.join(self.outdir, "endfile.csv"), f["status"], lat, lon, f["lifespan"],
                                  sp, f["endpoly"] if self.prm.settlementon else None)
        if self.nc is not None:
            self.nc.close()
        return f
