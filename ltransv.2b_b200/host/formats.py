"""The reference's text formats on either side of the particle loop:
`LTRANS.data` namelists, the particle / habitat CSV inputs, and the CSV outputs
(`para*.csv`, `endfile.csv`, `LandHits.csv`, `BottomHits.csv`, `ErrorLog.txt`).  The north star
keeps these formats unchanged; in production they stay with the Fortran host, this module is
the host-side mirror used by the Python driver and the tests."""
import os
import re

import numpy as np

from .binding import Params

# ----------------------------------------------------------------------------------------
# LTRANS.data: 14 namelists, old-style `$name ... $end` delimiters (also `&name ... /`), `!`
# comments, CRLF line ends (reference Model/parameter_module.f90:29-86, Model/LTRANS.h:47-269)
_GROUP = re.compile(r"^\s*[$&](\w+)\s*$")
_END = re.compile(r"^\s*([$&]end|/)\s*$", re.I)


def _value(tok):
    t = tok.strip().rstrip(",")
    if re.fullmatch(r"\.t(rue)?\.?", t, re.I):
        return True
    if re.fullmatch(r"\.f(alse)?\.?", t, re.I):
        return False
    if len(t) >= 2 and t[0] in "'\"" and t[-1] == t[0]:
        return t[1:-1]
    try:
        return int(t)
    except ValueError:
        pass
    try:
        return float(t.replace("d", "e").replace("D", "e"))
    except ValueError:
        return t


def read_namelist(path):
    """-> {group: {name: value}} with lower-cased names, in file order."""
    out, cur = {}, None
    with open(path, "r", newline="") as f:
        text = f.read().replace("\r\n", "\n").replace("\r", "\n")
    for raw in text.split("\n"):
        line, q = [], None                      # strip `!` comments outside quotes
        for ch in raw:
            if q:
                if ch == q:
                    q = None
            elif ch in "'\"":
                q = ch
            elif ch == "!":
                break
            line.append(ch)
        line = "".join(line).strip()
        if not line:
            continue
        m = _GROUP.match(line)
        if m and not _END.match(line):
            cur = out.setdefault(m.group(1).lower(), {})
            continue
        if _END.match(line):
            cur = None
            continue
        if cur is None or "=" not in line:
            continue
        k, v = line.split("=", 1)
        cur[k.strip().lower()] = _value(v)
    return out


def params_from_namelist(nml, **over):
    """ltgpu_params from the namelists the particle loop reads (LTRANS.h:45-269)."""
    flat = {}
    for grp in nml.values():
        flat.update(grp)
    p = Params.shipped()
    names = {f[0].lower(): f[0] for f in Params._fields_}
    for k, v in flat.items():
        if k in names and names[k] not in ("rng_mode", "field_dtype", "vturb_window_sigs", "vturb_fp32_walk"):
            if isinstance(v, bool):
                v = int(v)
            setattr(p, names[k], v)
    for k, v in over.items():
        setattr(p, k, v)
    return p, flat


# ----------------------------------------------------------------------------------------
# CSV inputs (list-directed reads, LTRANS.f90:252-273; settlement_module.f90:143-149, 190-197)
def read_particles_csv(path, settlementon):
    """rows: lon, lat, depth, dob[, startpoly] (ledger 25) -> arrays"""
    a = np.loadtxt(path, delimiter=",", ndmin=2)
    lon, lat, z, dob = a[:, 0], a[:, 1], a[:, 2], a[:, 3]
    startpoly = a[:, 4].astype(np.int32) if settlementon and a.shape[1] > 4 else np.zeros(len(a), np.int32)
    return lon, lat, z, dob, startpoly


def write_particles_csv(path, lon, lat, z, dob, startpoly=None):
    with open(path, "w") as f:
        for i in range(len(lon)):
            row = "%.10f,%.10f,%.6f,%.1f" % (lon[i], lat[i], z[i], dob[i])
            if startpoly is not None:
                row += ",%d" % startpoly[i]
            f.write(row + "\n")


def read_polygons_csv(path, ncol):
    """End_polygons.csv (5 columns) / End_holes.csv (6 columns), rows of one id contiguous."""
    return np.loadtxt(path, delimiter=",", ndmin=2)[:, :ncol]


# ----------------------------------------------------------------------------------------
# Fortran edit descriptors used by the writers
def _F(v, w, d):
    s = "%*.*f" % (w, d, v)
    if len(s) > w and s.lstrip().startswith(("0.", "-0.")):          # Fortran may drop the leading zero
        s = s.replace("0.", ".", 1)
    return s if len(s) <= w else "*" * w


def _I(v, w):
    s = "%*d" % (w, int(v))
    return s if len(s) <= w else "*" * w


def para_filename(prcount, outpath=""):
    """'para' + I8(prcount + 10000000) + '.csv' (LTRANS.f90:1729-1736); first file is ...002
    because prcount starts at 1 (:291) and is incremented before the write (:1624)."""
    return os.path.join(outpath, "para%8d.csv" % (prcount + 10000000))


_CHUNK = 1 << 20


def _fits(v, w, d):
    """every value prints within F<w>.<d> with the plain C format (no '*' fill, no dropped leading zero)"""
    v = np.asarray(v, np.float64)
    if not np.all(np.isfinite(v)):
        return False
    lim_pos = 10.0 ** (w - d - 1) - 1.0
    lim_neg = 10.0 ** (w - d - 2) - 1.0
    return bool(np.all((v < lim_pos) & (v > -lim_neg)))


def _write_rows(f, fmt, cols, slow_row):
    """Rows of fixed-format numbers.  Fast path: one C-level `%` per chunk of 2^20 rows (1.5 s per million
    rows instead of a Python loop per row); used when every value fits its field, which the caller has
    checked with _fits / integer ranges.  Otherwise the per-row routine with Fortran's overflow rules."""
    n = len(cols[0])
    if slow_row is not None:
        for k in range(n):
            f.write(slow_row(k))
        return
    for lo in range(0, n, _CHUNK):
        hi = min(n, lo + _CHUNK)
        flat = [None] * ((hi - lo) * len(cols))
        for c, col in enumerate(cols):
            flat[c::len(cols)] = col[lo:hi].tolist()
        f.write((fmt * (hi - lo)) % tuple(flat))


def _int_fits(v, w):
    v = np.asarray(v)
    return bool(np.all((v < 10 ** w) & (v > -(10 ** (w - 1)))))


def write_para_csv(path, z, status, lon, lat, salt=None, temp=None):
    """format 5: F10.3,',',I7,2(',',F9.4)[,2(',',F8.4)] = depth, status, lon, lat[, salt, temp]
    (LTRANS.f90:1738-1751)"""
    z, lon, lat = (np.asarray(a, np.float64) for a in (z, lon, lat))
    status = np.asarray(status).astype(np.int64)
    st = salt is not None
    ok = _fits(z, 10, 3) and _int_fits(status, 7) and _fits(lon, 9, 4) and _fits(lat, 9, 4)
    cols = [z, status, lon, lat]
    fmt = "%10.3f,%7d,%9.4f,%9.4f"
    if st:
        salt, temp = np.asarray(salt, np.float64), np.asarray(temp, np.float64)
        ok = ok and _fits(salt, 8, 4) and _fits(temp, 8, 4)
        cols += [salt, temp]
        fmt += ",%8.4f,%8.4f"

    def slow(n):
        row = _F(z[n], 10, 3) + "," + _I(status[n], 7) + "," + _F(lon[n], 9, 4) + "," + _F(lat[n], 9, 4)
        if st:
            row += "," + _F(salt[n], 8, 4) + "," + _F(temp[n], 8, 4)
        return row + "\n"
    with open(path, "w") as f:
        _write_rows(f, fmt + "\n", cols, None if ok else slow)


def write_endfile(path, status, lat, lon, lifespan, startpoly=None, endpoly=None):
    """settlement on: I7,I7,I7,F9.4,F9.4,I7 = startpoly, endpoly, status, lat, lon, lifespan;
    off: I7,F9.4,F9.4,I7 (LTRANS.f90:642-656)"""
    lat, lon = np.asarray(lat, np.float64), np.asarray(lon, np.float64)
    status = np.asarray(status).astype(np.int64)
    life = np.asarray(lifespan).astype(np.int64)                      # int(par(n,pLifespan))
    cols, fmt = [status, lat, lon, life], "%7d,%9.4f,%9.4f,%7d\n"
    ok = _int_fits(status, 7) and _fits(lat, 9, 4) and _fits(lon, 9, 4) and _int_fits(life, 7)
    if startpoly is not None:
        sp, ep = np.asarray(startpoly).astype(np.int64), np.asarray(endpoly).astype(np.int64)
        cols, fmt = [sp, ep] + cols, "%7d,%7d," + fmt
        ok = ok and _int_fits(sp, 7) and _int_fits(ep, 7)

    def slow(n):
        row = ""
        if startpoly is not None:
            row = _I(startpoly[n], 7) + "," + _I(endpoly[n], 7) + ","
        return row + _I(status[n], 7) + "," + _F(lat[n], 9, 4) + "," + _F(lon[n], 9, 4) + "," + _I(lifespan[n], 7) + "\n"
    with open(path, "w") as f:
        _write_rows(f, fmt, cols, None if ok else slow)


def append_hits(path, ids, lon, lat, z, age, time_s, hits):
    """format 101: I7,2(',',F9.4),',',F10.3,2(',',F10.5),',',I7 = id, lon, lat, depth, age(d),
    time(d), hits; only particles with hits > 0 (LTRANS.f90:1647-1661)"""
    sel = np.nonzero(np.asarray(hits) > 0)[0]
    ids_, lon_, lat_, z_ = (np.asarray(a)[sel] for a in (ids, lon, lat, z))
    aged, hit_ = np.asarray(age, np.float64)[sel] / 86400.0, np.asarray(hits)[sel].astype(np.int64)
    tday = np.full(len(sel), time_s / 86400.0)
    ok = (_int_fits(ids_, 7) and _fits(lon_, 9, 4) and _fits(lat_, 9, 4) and _fits(z_, 10, 3) and _fits(aged, 10, 5)
          and _fits(tday, 10, 5) and _int_fits(hit_, 7))

    def slow(k):
        return (_I(ids_[k], 7) + "," + _F(lon_[k], 9, 4) + "," + _F(lat_[k], 9, 4) + "," + _F(z_[k], 10, 3) + ","
                + _F(aged[k], 10, 5) + "," + _F(tday[k], 10, 5) + "," + _I(hit_[k], 7) + "\n")
    with open(path, "a") as f:
        _write_rows(f, "%7d,%9.4f,%9.4f,%10.3f,%10.5f,%10.5f,%7d\n",
                    [ids_.astype(np.int64), lon_.astype(np.float64), lat_.astype(np.float64), z_.astype(np.float64), aged, tday, hit_],
                    None if ok else slow)


_EVENT_TEXT = {
    11: "Particle %10d initially outside main bounds",          # LTRANS.f90:377, 400, 446-450: no time
    12: "Particle %10d initially inside island bounds",
    13: "Particle %10d initially not in rho element",
    14: "Particle %10d initially not in u element",
    15: "Particle %10d initially not in v element",
    21: "Particle %10d not in rho element after %10d seconds",
    22: "Particle %10d not in u element after %10d seconds",
    23: "Particle %10d not in v element after %10d seconds",
    24: "Particle %10d out after 3rd reflection after %10d seconds",
    25: "Particle %10d outside main bounds after intersect_reflect after %10d seconds",
    26: "Particle %10d inside island bounds after intersect_reflect after %10d seconds",
    27: "Particle %10d jumped over rho element after %10d seconds",
    28: "Particle %10d jumped over u element after %10d seconds",
    29: "Particle %10d jumped over v element after %10d seconds",
}


def append_errorlog(path, events):
    """ErrorLog.txt, formats 21-29 (LTRANS.f90:761-775); events = [(particle, code, time)]"""
    with open(path, "a") as f:
        for pid, code, t in events:
            f.write((_EVENT_TEXT[code] % pid if code < 20 else _EVENT_TEXT[code] % (pid, int(t))) + "\n")
