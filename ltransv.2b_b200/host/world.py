"""Synthetic ROMS-like world + the host-side setup tables the particle step consumes.

The reference's example inputs (grid / history NetCDF, CSVs) are external URLs
(reference README.md:36-55) and unavailable offline, so benchmarks and parity
tests run on this deterministic analytic world.  What is
restated here is host SETUP, not the hot path:
  * node numbering, wet-element compaction and adjacency lists
    (reference Model/hydrodynamic_module.f90:381-647),
  * lon/lat -> x/y, spherical branch (conversion_module.f90:176-197, 248-266),
  * a boundary tracer honouring the invariants of createBounds
    (boundary_module.f90:82-1215: mask erosion :166-196, vertices on U/V points
    between rho nodes of differing mask, domain edge on the outermost U/V line with
    open-ocean flags :1089-1096,1180-1183, main polygon first then islands with ids
    1001+ :1099,1175-1203) -- implemented as marching squares, not a port,
  * habitat polygon specs (settlement_module.f90:49-86, 275-472) as CSR lists.
All field values are rounded to float32 so float32 device storage is lossless
against the reference's float -> double NetCDF read.
"""
import numpy as np

PI_NML = 3.14159265358979        # LTRANS.data:154 (ledger 5)
EARTH_RADIUS = 6378000.0         # LTRANS.data:155


class Projection:
    """conversion_module.f90 double-precision spherical branch; lonmin/latmin are the
    namelist values minus 1 (parameter_module.f90:133-134)."""

    def __init__(self, lonmin_nml=-77.0, latmin_nml=36.0, pi=PI_NML, radius=EARTH_RADIUS):
        self.lonmin, self.latmin, self.pi, self.R = lonmin_nml - 1.0, latmin_nml - 1.0, pi, radius
        self.RCF = 180.0 / pi

    def lon2x(self, lon, lat):
        return (lon - self.lonmin) / 180.0 * self.R * self.pi * np.cos(lat / self.RCF)

    def lat2y(self, lat):
        return (lat - self.latmin) * self.R * self.pi / 180.0

    def y2lat(self, y):
        return y * self.RCF / self.R + self.latmin

    def x2lon(self, x, y):
        lat = y * 180.0 / (self.R * self.pi) + self.latmin
        return x * 180.0 / (self.R * self.pi * np.cos(lat / self.RCF)) + self.lonmin


def _f32(a):
    return np.asarray(a, dtype=np.float32)


def s_coordinate(N, theta_s=3.0, theta_b=0.4):
    """ROMS Song-Haidvogel stretching: (sc_r, Cs_r, sc_w, Cs_w); k=0 is the bottom."""
    def cs(s):
        return ((1 - theta_b) * np.sinh(theta_s * s) / np.sinh(theta_s)
                + theta_b * (np.tanh(theta_s * (s + 0.5)) / (2 * np.tanh(0.5 * theta_s)) - 0.5))
    sc_r = (np.arange(1, N + 1) - N - 0.5) / N
    sc_w = (np.arange(0, N + 1) - N) / N
    return sc_r, cs(sc_r), sc_w, cs(sc_w)


class World:
    def __init__(self, ni=130, nj=130, us=20, lon0=-76.2, lat0=36.9, dlon=0.004, dlat=0.0032,
                 hmin=5.0, hmax=30.0, islands=True, open_east=True, dt_hydro=3600.0, speed=0.6,
                 uniform=None):
        self.ni, self.nj, self.us, self.ws = ni, nj, us, us + 1
        self.dt_hydro, self.speed = dt_hydro, speed
        self.uniform = uniform      # (U0, V0, K0): steady uniform flow, flat bed, zeta = 0, angle = 0
        self.proj = Projection()
        i = np.arange(ni)[None, :]
        j = np.arange(nj)[:, None]
        self.lon_r = lon0 + dlon * i + 0.0 * j
        self.lat_r = lat0 + dlat * j + 0.0 * i
        self.lon_u = lon0 + dlon * (np.arange(ni - 1)[None, :] + 0.5) + 0.0 * j
        self.lat_u = lat0 + dlat * j + 0.0 * np.arange(ni - 1)[None, :]
        jv = np.arange(nj - 1)[:, None]
        self.lon_v = lon0 + dlon * i + 0.0 * jv
        self.lat_v = lat0 + dlat * (jv + 0.5) + 0.0 * i
        P = self.proj
        self.x_r, self.y_r = P.lon2x(self.lon_r, self.lat_r), P.lat2y(self.lat_r)
        self.x_u, self.y_u = P.lon2x(self.lon_u, self.lat_u), P.lat2y(self.lat_u)
        self.x_v, self.y_v = P.lon2x(self.lon_v, self.lat_v), P.lat2y(self.lat_v)
        # ---- mask: 2-cell land rim W/S/N, open ocean E, islands ------------------
        m = np.ones((nj, ni), dtype=np.int32)
        m[:, :2] = 0
        m[:2, :] = 0
        m[-2:, :] = 0
        if not open_east:
            m[:, -2:] = 0
        if islands:
            a, b = int(0.30 * nj), int(0.38 * ni)
            m[a:a + max(3, nj // 16), b:b + max(4, ni // 14)] = 0          # block island
            a, b = int(0.62 * nj), int(0.55 * ni)
            w = max(3, ni // 20)
            m[a:a + 2 * w, b:b + w] = 0                                     # L-shaped island
            m[a:a + w, b:b + 2 * w] = 0
            a, b = int(0.5 * nj), int(0.2 * ni)
            m[a, b] = 0; m[a + 1, b + 1] = 0                                # diagonal pair
            m[a + 1, b] = 0
        self.mask_rho = self._erode(m)
        self.mask_u = self.mask_rho[:, :-1] * self.mask_rho[:, 1:]
        self.mask_v = self.mask_rho[:-1, :] * self.mask_rho[1:, :]
        # ---- bathymetry / angle ---------------------------------------------------
        fi, fj = i / (ni - 1.0), j / (nj - 1.0)
        h = hmin + (hmax - hmin) * (0.35 + 0.45 * fi + 0.2 * np.sin(2 * np.pi * fj) * np.cos(3 * np.pi * fi))
        self.h = _f32(np.clip(h, hmin, hmax)).astype(np.float64)
        self.angle = 0.04 * np.sin(2 * np.pi * fi) * np.cos(np.pi * fj) + 0.01
        if uniform is not None:
            self.h = np.full((nj, ni), float(np.float32(hmax)))
            self.angle = np.zeros((nj, ni))
        self.sc_r, self.Cs_r, self.sc_w, self.Cs_w = s_coordinate(us)
        self._grid = None

    @staticmethod
    def _erode(m):
        """boundary_module.f90:166-196: water nodes with < 2 water 4-neighbours -> land."""
        m = m.copy()
        while True:
            p = np.pad(m, 1)
            nb = p[:-2, 1:-1] + p[2:, 1:-1] + p[1:-1, :-2] + p[1:-1, 2:]
            bad = (m == 1) & (nb < 2)
            if not bad.any():
                return m
            m[bad] = 0

    # ------------------------------------------------------------------ grid --
    def _elements(self, ni, nj, mask):
        """hydro:416-519 (corner order, wet compaction) and :590-647 (adjacency)."""
        ie, je = np.meshgrid(np.arange(ni - 1), np.arange(nj - 1))
        n1 = ie + je * ni + 1
        E = np.stack([n1, n1 + 1, n1 + 1 + ni, n1 + ni], axis=-1).reshape(-1, 4)   # 1-based
        mk = mask.reshape(-1)
        wet = (mk[E - 1] == 1).any(axis=1)
        comp = np.zeros((nj - 1) * (ni - 1), dtype=np.int64)
        comp[wet] = np.arange(1, wet.sum() + 1)
        comp2 = np.pad(comp.reshape(nj - 1, ni - 1), 1)
        nE = int(wet.sum())
        adj = np.zeros((10, nE), dtype=np.int32)                # Fortran (nE,10) column-major
        adj[0] = np.arange(1, nE + 1)
        cnt = np.ones(nE, dtype=np.int64)
        wj, wi = np.nonzero(wet.reshape(nj - 1, ni - 1))
        for dj, di in ((-1, -1), (-1, 0), (-1, 1), (0, -1), (0, 1), (1, -1), (1, 0), (1, 1)):
            nb = comp2[wj + 1 + dj, wi + 1 + di]
            ok = nb > 0
            adj[cnt[ok], np.nonzero(ok)[0]] = nb[ok]
            cnt[ok] += 1
        return E[wet].astype(np.int32), adj, comp.reshape(nj - 1, ni - 1)

    def grid(self):
        if self._grid is not None:
            return self._grid
        ni, nj = self.ni, self.nj
        RE, rAdj, self.rcomp = self._elements(ni, nj, self.mask_rho)
        UE, uAdj, self.ucomp = self._elements(ni - 1, nj, self.mask_u)
        VE, vAdj, self.vcomp = self._elements(ni, nj - 1, self.mask_v)
        g = dict(vi=ni, uj=nj, ui=ni - 1, vj=nj - 1,
                 rx=self.x_r.ravel(), ry=self.y_r.ravel(), ux=self.x_u.ravel(), uy=self.y_u.ravel(),
                 vx=self.x_v.ravel(), vy=self.y_v.ravel(), depth=self.h.ravel(), angle=self.angle.ravel(),
                 rho_mask=self.mask_rho.ravel(), u_mask=self.mask_u.ravel(), v_mask=self.mask_v.ravel(),
                 SC=self.sc_r, CS=self.Cs_r, SCW=self.sc_w, CSW=self.Cs_w,
                 RE=RE, UE=UE, VE=VE, nRE=len(RE), nUE=len(UE), nVE=len(VE),
                 rAdj=rAdj, uAdj=uAdj, vAdj=vAdj)
        self._grid = g
        return g

    # ---------------------------------------------------------------- fields --
    def record(self, r, dtype=np.float32):
        """One ROMS history record at t = r*dt_hydro, arrays in ROMS memory order
        (level, eta, xi) i.e. node fastest -- what one NF90_GET_VAR returns
        (hydro:1140-1364)."""
        ni, nj, us, ws = self.ni, self.nj, self.us, self.ws
        if self.uniform is not None:
            U0, V0, K0 = self.uniform
            return dict(zeta=np.zeros((nj, ni), dtype), u=np.full((us, nj, ni - 1), U0, dtype),
                        v=np.full((us, nj - 1, ni), V0, dtype), w=np.zeros((ws, nj, ni), dtype),
                        aks=np.full((ws, nj, ni), K0, dtype), salt=np.full((us, nj, ni), 20.0, dtype),
                        temp=np.full((us, nj, ni), 15.0, dtype))
        t = r * self.dt_hydro
        om = 2 * np.pi / 44714.0                                   # M2
        i = np.arange(ni)[None, :]; j = np.arange(nj)[:, None]
        zeta = 0.5 * np.sin(om * t - 2 * np.pi * i / ni) * (0.6 + 0.4 * j / nj)

        def psi_u(ii, jj):   # u = -dpsi/dy (index space), gyre modulated in time
            return -np.sin(np.pi * ii / (ni - 1)) * np.cos(np.pi * jj / (nj - 1)) * (1 + 0.3 * np.sin(om * t))

        def psi_v(ii, jj):
            return np.cos(np.pi * ii / (ni - 1)) * np.sin(np.pi * jj / (nj - 1)) * (1 + 0.3 * np.sin(om * t))
        iu = np.arange(ni - 1)[None, :] + 0.5
        jv = np.arange(nj - 1)[:, None] + 0.5
        u2 = self.speed * (0.7 * psi_u(iu, j) + 0.3 * np.cos(om * t) * np.ones_like(iu * j))
        v2 = self.speed * (0.7 * psi_v(i, jv) + 0.1 * np.sin(om * t + 1.0) * np.ones_like(i * jv))
        kr = (np.arange(us) + 0.5) / us                            # 0 bottom .. 1 surface
        kw = np.arange(ws) / us
        shear = (0.55 + 0.45 * kr ** 0.7)[:, None, None]
        u = shear * u2[None]
        v = shear * v2[None]
        w = 2e-4 * np.sin(np.pi * kw)[:, None, None] * (np.sin(2 * np.pi * i / ni) * np.cos(2 * np.pi * j / nj)
                                                        * np.cos(om * t))[None]
        kmax = 4e-3 * (1 + 0.5 * np.sin(2 * np.pi * i / ni) * np.cos(om * t) * np.cos(np.pi * j / nj))
        aks = 1e-5 + (4 * kw * (1 - kw))[:, None, None] * kmax[None] * (0.8 + 0.2 * np.sin(7 * np.pi * kw))[:, None, None]
        salt = (12 + 12 * i / ni + 0 * j)[None] + (6 * np.tanh((0.5 - kr) * 6.0))[:, None, None] \
            + 0.5 * np.sin(om * t)
        temp = (18 + 0 * i * j)[None] + (6 * kr)[:, None, None] + 0.3 * np.cos(om * t)
        return dict(zeta=zeta.astype(dtype), u=u.astype(dtype), v=v.astype(dtype), w=w.astype(dtype),
                    aks=aks.astype(dtype), salt=salt.astype(dtype), temp=temp.astype(dtype))

    # ---------------------------------------------------------------- bounds --
    def bounds(self):
        """Marching-squares boundary tracer (see module docstring)."""
        ni, nj, m = self.ni, self.nj, getattr(self, "mask_bnd", self.mask_rho)
        s = m.copy()
        s[0, :] = 0; s[-1, :] = 0; s[:, 0] = 0; s[:, -1] = 0
        ring = np.zeros_like(m, dtype=bool)
        ring[0, :] = ring[-1, :] = True; ring[:, 0] = ring[:, -1] = True

        def vert(kind, i, j):
            if kind == 'u':
                return (self.x_u[j, i], self.y_u[j, i])
            return (self.x_v[j, i], self.y_v[j, i])

        def is_open(kind, i, j):          # between an original-water ring node and a water inner node
            a, b = ((j, i), (j, i + 1)) if kind == 'u' else ((j, i), (j + 1, i))
            return bool(m[a] == 1 and m[b] == 1 and (ring[a] != ring[b]))
        segs = {}
        c0 = s[:-1, :-1]; c1 = s[:-1, 1:]; c2 = s[1:, 1:]; c3 = s[1:, :-1]
        mixed = (c0 + c1 + c2 + c3 > 0) & (c0 + c1 + c2 + c3 < 4)
        nbr = {}

        def link(a, b):
            nbr.setdefault(a, []).append(b)
            nbr.setdefault(b, []).append(a)
        for j, i in zip(*np.nonzero(mixed)):
            j, i = int(j), int(i)
            s0, s1, s2, s3 = s[j, i], s[j, i + 1], s[j + 1, i + 1], s[j + 1, i]
            e = [('u', i, j), ('v', i + 1, j), ('u', i, j + 1), ('v', i, j)]
            crossed = [s0 != s1, s1 != s2, s3 != s2, s0 != s3]
            k = [q for q in range(4) if crossed[q]]
            if len(k) == 2:
                link(e[k[0]], e[k[1]])
            elif len(k) == 4:
                if s0 == 1:
                    link(e[3], e[0]); link(e[1], e[2])
                else:
                    link(e[0], e[1]); link(e[2], e[3])
        del segs
        loops, seen = [], set()
        for start in sorted(nbr):
            if start in seen:
                continue
            loop, prev, cur = [start], None, start
            seen.add(start)
            while True:
                nx = [q for q in nbr[cur] if q != prev]
                nxt = nx[0] if nx else nbr[cur][0]
                if nxt == start:
                    break
                if nxt in seen:     # saddle vertex revisited: close
                    break
                loop.append(nxt); seen.add(nxt)
                prev, cur = cur, nxt
            loops.append(loop)

        def area(loop):
            p = np.array([vert(*q) for q in loop])
            x, y = p[:, 0], p[:, 1]
            return 0.5 * np.sum(x * np.roll(y, -1) - np.roll(x, -1) * y)
        loops = [lp if area(lp) < 0 else lp[::-1] for lp in loops]          # clockwise
        loops.sort(key=lambda lp: -abs(area(lp)))
        main, isl = loops[0], loops[1:]
        bx = np.array([vert(*q)[0] for q in main + [main[0]]])
        by = np.array([vert(*q)[1] for q in main + [main[0]]])
        opn = [is_open(*q) for q in main + [main[0]]]
        seg_x, seg_y, land = [], [], []
        for k in range(len(bx) - 1):
            seg_x.append((bx[k], bx[k + 1])); seg_y.append((by[k], by[k + 1]))
            land.append(0 if (opn[k] and opn[k + 1]) else 1)
        hx, hy, hid = [], [], []
        for n, lp in enumerate(isl):
            px = [vert(*q)[0] for q in lp + [lp[0]]]
            py = [vert(*q)[1] for q in lp + [lp[0]]]
            hx += px; hy += py; hid += [1001 + n] * len(px)
            for k in range(len(px) - 1):
                seg_x.append((px[k], px[k + 1])); seg_y.append((py[k], py[k + 1])); land.append(1)
        return dict(bnd_x=np.array(seg_x, dtype=np.float64).reshape(-1, 2),   # (nbounds,2) C == Fortran (2,nbounds)
                    bnd_y=np.array(seg_y, dtype=np.float64).reshape(-1, 2),
                    land=np.array(land, dtype=np.int32), bx=bx, by=by,
                    hx=np.array(hx, dtype=np.float64), hy=np.array(hy, dtype=np.float64),
                    hid=np.array(hid, dtype=np.int32), nislands=len(isl))

    # --------------------------------------------------------------- habitat --
    def index_to_xy(self, fi, fj):
        """bilinear map of fractional rho index (fi,fj) to (x,y) on the rho quads."""
        i0 = np.clip(np.floor(fi).astype(int), 0, self.ni - 2)
        j0 = np.clip(np.floor(fj).astype(int), 0, self.nj - 2)
        a, b = fi - i0, fj - j0
        def bl(F):
            return ((1 - a) * (1 - b) * F[j0, i0] + a * (1 - b) * F[j0, i0 + 1]
                    + a * b * F[j0 + 1, i0 + 1] + (1 - a) * b * F[j0 + 1, i0])
        return bl(self.x_r), bl(self.y_r)

    def habitat(self, npoly=8, holes=True, seed=7, nvert=12, radius_cells=1.6):
        """End_polygons.csv / End_holes.csv equivalents + createPolySpecs CSR lists."""
        rng = np.random.Generator(np.random.Philox(seed))
        g = self.grid()
        wet = np.argwhere(self._interior_water(3))
        pick = wet[rng.choice(len(wet), size=npoly, replace=False)]
        dx = float(np.mean(np.diff(self.x_r[self.nj // 2])))
        rows, hrows, ids, hids = [], [], [], []
        for k, (j, i) in enumerate(pick):
            cx, cy = self.x_r[j, i], self.y_r[j, i]
            pid = 101001 + k
            R = radius_cells * dx * (0.8 + 0.4 * rng.random())
            th = -2 * np.pi * np.arange(nvert + 1) / nvert          # closed, first point repeated
            th[-1] = th[0]
            for t in th:
                rows.append((pid, cx, cy, cx + R * np.cos(t), cy + R * np.sin(t)))
            ids.append(pid)
            if holes and k % 2 == 0:
                hid = 100201 + k
                th = -2 * np.pi * np.arange(9) / 8; th[-1] = th[0]
                for t in th:
                    hrows.append((hid, cx, cy, cx + 0.35 * R * np.cos(t), cy + 0.35 * R * np.sin(t), pid))
                hids.append(hid)
        polys = np.array(rows, dtype=np.float64)
        hol = np.array(hrows, dtype=np.float64).reshape(-1, 6)

        # polyspecs / holespecs / elepolys / polyholes exactly as createPolySpecs builds them
        # (settlement_module.f90:245-480), through the bucketed routine of host/polyspecs.py
        from .polyspecs import create_poly_specs
        return create_poly_specs(g["rx"][g["RE"] - 1], g["ry"][g["RE"] - 1], polys, hol)

    # ------------------------------------------------------------- particles --
    def _interior_water(self, margin):
        m = self.mask_rho.astype(bool)
        ok = m.copy()
        for _ in range(margin):
            p = np.pad(ok, 1)
            ok = ok & p[:-2, 1:-1] & p[2:, 1:-1] & p[1:-1, :-2] & p[1:-1, 2:]
        ok[:, -margin - 1:] = False
        return ok

    def seed_particles(self, n, seed=1234, margin=2, dob_max=0.0, locate=True):
        """Random release points (x, y, z, dob) inside water, >= margin cells off land,
        plus their rho/u/v elements (what setEle_all hydro:1536 would find); locate=False
        returns None for the elements (ltgpu_set_particles then locates on the device)."""
        rng = np.random.Generator(np.random.Philox(seed))
        cells = np.argwhere(self._interior_water(margin)[:-1, :-1])
        c = cells[rng.integers(0, len(cells), size=n)]
        fi = c[:, 1] + rng.random(n); fj = c[:, 0] + rng.random(n)
        x, y = self.index_to_xy(fi, fj)
        i0, j0 = np.floor(fi).astype(int), np.floor(fj).astype(int)
        hloc = self.h[j0, i0]
        z = -hloc * (0.08 + 0.84 * rng.random(n))
        dob = np.zeros(n) if dob_max <= 0 else np.floor(rng.random(n) * dob_max / 120.0) * 120.0
        r, u, v = self.locate(x, y, fi, fj) if locate else (None, None, None)
        return x, y, z, dob, r, u, v

    def locate(self, x, y, fi, fj):
        g = self.grid()

        def one(comp, nodes_x, nodes_y, E, gi, gj):
            nje, nie = comp.shape
            best = np.zeros(len(x), dtype=np.int32)
            for dj, di in ((0, 0), (0, -1), (0, 1), (-1, 0), (1, 0), (-1, -1), (-1, 1), (1, -1), (1, 1)):
                jj = np.clip(gj + dj, 0, nje - 1); ii = np.clip(gi + di, 0, nie - 1)
                e = comp[jj, ii]
                todo = (best == 0) & (e > 0)
                if not todo.any():
                    continue
                q = E[e[todo] - 1] - 1
                qx, qy = nodes_x[q], nodes_y[q]
                px, py = x[todo][:, None], y[todo][:, None]
                cr = (np.roll(qx, -1, 1) - qx) * (py - qy) - (np.roll(qy, -1, 1) - qy) * (px - qx)
                inside = (cr > 0).all(1) | (cr < 0).all(1)
                t = np.nonzero(todo)[0]
                best[t[inside]] = e[todo][inside]
            return best
        r = one(self.rcomp, g["rx"], g["ry"], g["RE"], np.floor(fi).astype(int), np.floor(fj).astype(int))
        u = one(self.ucomp, g["ux"], g["uy"], g["UE"], np.floor(fi - 0.5).astype(int), np.floor(fj).astype(int))
        v = one(self.vcomp, g["vx"], g["vy"], g["VE"], np.floor(fi).astype(int), np.floor(fj - 0.5).astype(int))
        if (r == 0).any() or (u == 0).any() or (v == 0).any():
            raise RuntimeError("seed_particles: a particle was not located in a rho/u/v element")
        return r, u, v
