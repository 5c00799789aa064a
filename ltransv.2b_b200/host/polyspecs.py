"""Habitat polygon tables for ltgpu_set_habitat: the reference's `createPolySpecs`
(Model/settlement_module.f90:245-480) and the reject radii of `getHabitat` (:167-236), rebuilt so
that set-up does not cost O(rho_elements x pedges) exact point-in-element tests (SURVEY.md 8f,
rank 3): 786 k elements x 10^4 edge points is 10^10 `gridcell` calls in the reference.

Same answer, different search: a polygon can only be listed for an element whose bounding box meets
the polygon's bounding box (an edge point inside the element lies in both boxes; an element corner
inside the polygon lies in the polygon's box), so the exact predicates -- restated here from
gridcell_module.f90:26-257 and point_in_polygon_module.f90:25-167, with their on-edge / on-vertex
rules -- run only on those candidates, found with vectorised box tests.

Reference criterion, kept literally (createPolySpecs :296-402), for element i and polygon q:
  (1) some edge point of q lies in element i (gridcell), or
  (2) some corner of element i is closer than maxbdis(q) to q's centre AND some corner of element i
      lies in q (inpoly, points on the outline count as inside).
Polygons are listed per element in file order, holes per polygon in file order.
"""
import numpy as np


# ---------------------------------------------------------------- exact predicates (scalar)
def gridcell(ex, ey, X, Y):
    """gridcell_module.f90:26-257 for one element: True <=> triangle /= 0"""
    if all(Y < v for v in ey) or all(Y > v for v in ey):
        return False
    if all(X < v for v in ex) or all(X > v for v in ex):
        return False
    for k in range(4):
        if X == ex[k] and Y == ey[k]:
            return True
    for a, b in ((0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)):
        if ey[a] == ey[b] and Y == ey[a]:
            return (ex[a] > ex[b] and ex[b] < X < ex[a]) or (ex[b] > ex[a] and ex[a] < X < ex[b])
    if Y in (ey[0], ey[1], ey[2], ey[3]) and (Y == max(ey) or Y == min(ey)):
        return False
    total = 0
    for p in range(4):
        bx1, by1, bx2, by2 = ex[p], ey[p], ex[(p + 1) % 4], ey[(p + 1) % 4]
        if (X <= bx1 or X <= bx2) and ((by1 > by2 and by2 <= Y <= by1) or (by2 > by1 and by1 <= Y <= by2)):
            if bx1 == bx2:
                if X == bx1:
                    return True
                total += 0 if Y == by2 else 1
            else:
                slope = (by1 - by2) / (bx1 - bx2)
                xi = (Y - by1 + (slope * bx1)) / slope
                if xi == X:
                    return True
                if xi > X and Y != by2:
                    total += 1
    return total % 2 != 0


def inpoly(x, y, px, py, onin=True):
    """point_in_polygon_module.f90:25-167 on a closed outline (first point repeated)"""
    n = len(px)
    hilo = [0] * (n + 2)
    on = False
    for i in range(1, n + 1):
        if py[i - 1] > y:
            hilo[i] = 1
        elif py[i - 1] < y:
            hilo[i] = -1
        if py[i - 1] == y and px[i - 1] > x:
            on = True
        if px[i - 1] == x and py[i - 1] == y:
            return onin
    crossed = 0
    if on:
        first, i = True, 1
        while i <= n:
            if hilo[i] == 0 and px[i - 1] > x:
                if first:
                    i += 1
                    continue
                if hilo[i - 1] == 0:
                    return onin
                j = 1
                while True:
                    if i + j == n + 1:
                        j = 2 - i
                    if hilo[i + j] != 0:
                        break
                    if px[i + j - 1] < x:
                        return onin
                    j += 1
                if hilo[i - 1] + hilo[i + j] == 0:
                    crossed += 1
                if j < 0:
                    break
                i += j
            first = False
            i += 1
    for i in range(1, n):
        ax, ay, bx, by = px[i - 1], py[i - 1], px[i], py[i]
        if (ax <= x and bx <= x) or (ay <= y and by <= y) or (ay >= y and by >= y):
            continue
        if ax > x and bx > x:
            crossed += 1
            continue
        m = (by - ay) / (bx - ax)
        ix = (y - (ay - m * ax)) / m
        if ix == x:
            return onin
        if ix > x:
            crossed += 1
    return crossed % 2 != 0


# ---------------------------------------------------------------------------- the tables
def _specs(tab):
    """first row (1-based) and row count of every id, ids in file order (createPolySpecs :275-292, :408-421)"""
    ids = tab[:, 0]
    first = np.r_[True, ids[1:] != ids[:-1]]
    start = np.nonzero(first)[0]
    size = np.diff(np.r_[start, len(ids)])
    return np.rint(ids[start]).astype(np.int32), (start + 1).astype(np.int32), size.astype(np.int32)


def _maxdis(tab, start, size):
    """getHabitat :167-172 / :230-236: 1.0, or the largest centre-to-edge-point distance"""
    d = np.sqrt((tab[:, 1] - tab[:, 3]) ** 2 + (tab[:, 2] - tab[:, 4]) ** 2)
    return np.array([max(1.0, float(d[s - 1:s - 1 + z].max())) for s, z in zip(start, size)], dtype=np.float64)


def create_poly_specs(r_ele_x, r_ele_y, polys, holes=None):
    """r_ele_x, r_ele_y: (nE, 4) corner coordinates of the rho elements (getR_ele); polys (pedges, 5) rows
    [id, centre x, centre y, edge x, edge y]; holes (hedges, 6) rows [id, cx, cy, ex, ey, parent polygon id].
    Returns the dictionary ltgpu_set_habitat takes (host.binding.LtransLib.set_habitat)."""
    ex, ey = np.asarray(r_ele_x, np.float64), np.asarray(r_ele_y, np.float64)
    polys = np.asarray(polys, np.float64).reshape(-1, 5)
    nE = len(ex)
    pid, pstart, psize = _specs(polys)
    pmax = _maxdis(polys, pstart, psize)
    exmin, exmax, eymin, eymax = ex.min(1), ex.max(1), ey.min(1), ey.max(1)
    pairs = []                                        # (element, polygon position) in the reference's listing order
    for q, (s, z) in enumerate(zip(pstart, psize)):
        px, py = polys[s - 1:s - 1 + z, 3], polys[s - 1:s - 1 + z, 4]
        cx, cy = polys[s - 1 + z - 1, 1], polys[s - 1 + z - 1, 2]        # centre as read on the polygon's LAST row (:333-347)
        cand = np.nonzero((exmax >= px.min()) & (exmin <= px.max()) & (eymax >= py.min()) & (eymin <= py.max()))[0]
        for e in cand:
            qx, qy = ex[e], ey[e]
            hit = False
            # (1) an edge point in the element: only points inside the element's box can be
            inbox = np.nonzero((px >= exmin[e]) & (px <= exmax[e]) & (py >= eymin[e]) & (py <= eymax[e]))[0]
            for k in inbox:
                if gridcell(qx, qy, float(px[k]), float(py[k])):
                    hit = True
                    break
            if not hit:
                dis = np.sqrt((qx - cx) ** 2 + (qy - cy) ** 2)
                if (dis < pmax[q]).any():
                    for k in range(4):
                        if inpoly(float(qx[k]), float(qy[k]), px, py):
                            hit = True
                            break
            if hit:
                pairs.append((int(e), q))
    pairs.sort()                                      # per element, polygons in file order
    eptr = np.zeros(nE + 1, np.int32)
    for e, _ in pairs:
        eptr[e + 1] += 1
    eptr = np.cumsum(eptr).astype(np.int32)
    eidx = np.array([q for _, q in pairs], dtype=np.int32)

    if holes is not None and len(holes):
        holes = np.asarray(holes, np.float64).reshape(-1, 6)
        hid, hstart, hsize = _specs(holes)
        hmax = _maxdis(holes, hstart, hsize)
        parent = np.rint(holes[hstart - 1, 5]).astype(np.int64)
        hptr, hidx = [0], []
        for q in pid:                                 # :427-470: holes of polygon q in file order
            hidx += [k for k in range(len(hid)) if parent[k] == q]
            hptr.append(len(hidx))
    else:
        holes = np.zeros((0, 6))
        hid = hstart = hsize = np.zeros(0, np.int32)
        hmax = np.zeros(0)
        hptr, hidx = [0] * (len(pid) + 1), []
    return dict(pedges=len(polys), polys=np.ascontiguousarray(polys.T), hedges=len(holes),
                holes=np.ascontiguousarray(holes.T) if len(holes) else np.zeros((6, 0)),
                poly_id=pid, poly_start=pstart, poly_size=psize, poly_maxdis=pmax,
                hole_id=hid, hole_start=hstart, hole_size=hsize, hole_maxdis=hmax,
                elepoly_ptr=eptr, elepoly_idx=eidx,
                polyhole_ptr=np.array(hptr, dtype=np.int32), polyhole_idx=np.array(hidx, dtype=np.int32))
