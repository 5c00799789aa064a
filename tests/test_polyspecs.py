"""createPolySpecs (settlement_module.f90:245-480): the bucketed routine of the host layer
(ltransv.2b_b200/host/polyspecs.py) against the literal element x edge-point loops restated in
oracle/np_leaf.py, incl. polygons whose vertices sit exactly on element corners and edges."""
import numpy as np

from common import SMALL, World
from ltrans_b200.host.polyspecs import create_poly_specs
from oracle import np_leaf as NL


def _lists(h):
    return [[int(h["poly_id"][k]) for k in h["elepoly_idx"][h["elepoly_ptr"][e]:h["elepoly_ptr"][e + 1]]]
            for e in range(len(h["elepoly_ptr"]) - 1)]


def _literal(ex, ey, polys, h):
    return NL.create_poly_specs_literal(ex.tolist(), ey.tolist(), polys.tolist(), {int(i): float(m) for i, m in zip(h["poly_id"], h["poly_maxdis"])})


def test_world_habitat_equals_literal_createPolySpecs():
    w = World(**SMALL)
    g = w.grid()
    ex, ey = g["rx"][g["RE"] - 1], g["ry"][g["RE"] - 1]
    for npoly, holes in ((8, True), (20, False)):
        h = w.habitat(npoly=npoly, holes=holes)
        polys = np.ascontiguousarray(h["polys"].T)
        assert _lists(h) == _literal(ex, ey, polys, h)
        assert h["elepoly_ptr"][-1] == len(h["elepoly_idx"]) > npoly
        # specs: contiguous rows per id, first row 1-based
        for k, (s, z) in enumerate(zip(h["poly_start"], h["poly_size"])):
            assert np.all(polys[s - 1:s - 1 + z, 0] == h["poly_id"][k])
        if holes:
            hol = np.ascontiguousarray(h["holes"].T)
            for q in range(npoly):
                for k in h["polyhole_idx"][h["polyhole_ptr"][q]:h["polyhole_ptr"][q + 1]]:
                    assert hol[h["hole_start"][k] - 1, 5] == h["poly_id"][q]


def test_polygons_on_element_corners_and_edges():
    """vertices exactly on rho nodes and on element edges: the reference's on-vertex / on-edge rules decide"""
    w = World(**SMALL)
    g = w.grid()
    ex, ey = g["rx"][g["RE"] - 1], g["ry"][g["RE"] - 1]
    X, Y = w.x_r, w.y_r
    rows = []
    def poly(pid, pts):
        cx, cy = float(np.mean([p[0] for p in pts])), float(np.mean([p[1] for p in pts]))
        for p in pts + [pts[0]]:
            rows.append((pid, cx, cy, p[0], p[1]))
    j, i = 12, 14
    poly(7001, [(X[j, i], Y[j, i]), (X[j, i + 2], Y[j, i + 2]), (X[j + 2, i + 2], Y[j + 2, i + 2]), (X[j + 2, i], Y[j + 2, i])])      # corners on nodes
    poly(7002, [(0.5 * (X[20, 20] + X[20, 21]), Y[20, 20]), (X[20, 23], Y[20, 23]), (X[22, 22], 0.5 * (Y[22, 22] + Y[23, 22]))])     # on edges
    poly(7003, [(X[8, 25] + 1.0, Y[8, 25] + 1.0), (X[8, 25] + 30.0, Y[8, 25] + 2.0), (X[8, 25] + 15.0, Y[8, 25] + 25.0)])           # inside one element
    big = [(X[26, 8] - 10, Y[26, 8] - 10), (X[26, 14] + 10, Y[26, 14] - 10), (X[31, 14] + 10, Y[31, 14] + 10), (X[31, 8] - 10, Y[31, 8] + 10)]
    poly(7004, big)                                                                                                                  # contains whole elements
    polys = np.array(rows, dtype=np.float64)
    h = create_poly_specs(ex, ey, polys, None)
    lit = _literal(ex, ey, polys, h)
    assert _lists(h) == lit
    n_in = sum(1 for l in lit if 7004 in l)
    assert n_in >= 30 and sum(1 for l in lit if 7003 in l) >= 1 and sum(1 for l in lit if 7001 in l) >= 4
