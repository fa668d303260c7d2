"""The reference's neighbour graph built on the GPU (dsc_delaunay_build / dsc_set_graph_delaunay) against
Qhull (scipy, the library the reference calls: Geometry.cc:333-341 "d Qbb Qt") and the oracle's cot weights: adjacency
bit-exact, weights / area to rounding."""
import numpy as np
import pytest

from oracle import scenes, edges, graph as ograph

pytestmark = pytest.mark.gpu


def _points(kind, n, seed):
    rng = np.random.default_rng(seed)
    if kind == "normal":                                   # create_data.py's distribution (config 2)
        V = np.stack([rng.normal(0, 0.03, n), rng.normal(0, 0.03, n), rng.normal(0.2, 0.01, n)], 1)
    elif kind == "uniform":
        V = np.stack([rng.uniform(-0.1, 0.1, n), rng.uniform(-0.05, 0.05, n), rng.normal(0.2, 0.01, n)], 1)
    else:                                                  # two clusters of very different density + a few far outliers
        a = rng.normal(0, 0.001, (n // 2, 2))
        b = rng.normal(0.2, 0.05, (n - n // 2 - 3, 2))
        c = np.array([[3.0, 0.1], [-2.0, 1.5], [0.5, -4.0]])
        xy = np.concatenate([a, b, c])
        V = np.concatenate([xy, rng.normal(0.2, 0.01, (n, 1))], 1)
    return V.astype(np.float32)


@pytest.mark.parametrize("kind,n,seed", [("normal", 7, 0), ("normal", 120, 1), ("normal", 5000, 2), ("uniform", 20000, 3), ("clusters", 30000, 4),
                                         ("normal", 200000, 5)])
def test_delaunay_graph_matches_qhull(pkg, ctx, kind, n, seed):
    V = _points(kind, n, seed)
    g = ograph.delaunay_graph(V.astype(np.float64))
    rowptr, col, w, area, ntri, nsecond = ctx.delaunay_graph(V)
    assert ntri == g.n_triangles
    assert np.array_equal(rowptr, g.rowptr) and np.array_equal(col, g.col)                 # the adjacency, bit for bit
    np.testing.assert_allclose(w, g.w, rtol=1e-12, atol=1e-14)
    assert area == pytest.approx(g.area, rel=1e-12)
    assert 0 < nsecond < max(64, 40 * np.sqrt(n))                                          # only the rim goes through the second pass
    # both directions of an edge carry the same bits
    row = np.repeat(np.arange(n), np.diff(rowptr))
    key = row.astype(np.int64) * n + col
    rev = col.astype(np.int64) * n + row
    assert np.array_equal(w[np.argsort(key)], w[np.argsort(rev)])


def test_delaunay_edge_cases(pkg, ctx):
    for n in (0, 1, 2):
        rowptr, col, w, area, ntri, ns = ctx.delaunay_graph(np.zeros((n, 3), np.float32) + np.arange(n)[:, None])
        assert len(col) == 0 and ntri == 0
    V = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)                            # one triangle
    rowptr, col, w, area, ntri, ns = ctx.delaunay_graph(V)
    assert ntri == 1 and list(rowptr) == [0, 2, 4, 6] and area == pytest.approx(0.5)
    with pytest.raises(pkg.DscError):                                                      # no extent
        ctx.delaunay_graph(np.zeros((5, 3), np.float32))


def test_coincident_points_get_no_vertex(pkg, ctx):
    """two points with the same (x, y) (different z): the later one is left out of the mesh, as Qhull and the host triangulator
    leave it out; everything else is the triangulation of the distinct locations"""
    V = _points("normal", 4000, 11)
    twins = [(17, 2500), (300, 301), (5, 3999)]
    for a, b in twins:
        V[b, :2] = V[a, :2]
    rowptr, col, w, area, ntri, ns = ctx.delaunay_graph(V)
    deg = np.diff(rowptr)
    keep = np.ones(len(V), bool)
    keep[[b for _, b in twins]] = False
    assert np.all(deg[~keep] == 0) and np.all(deg[keep] > 0)
    g = ograph.delaunay_graph(V[keep].astype(np.float64))
    new_of_old = np.cumsum(keep) - 1
    old_of_new = np.nonzero(keep)[0]
    assert ntri == g.n_triangles
    assert np.array_equal(deg[keep], np.diff(g.rowptr)) and np.array_equal(new_of_old[col], g.col)
    np.testing.assert_allclose(w, g.w, rtol=1e-12, atol=1e-14)
    assert np.array_equal(old_of_new[g.col], col)


def _exact_flip_verdicts(V, rp_a, col_a, only_a):
    """for every edge (p, q) of mesh A that mesh B lacks: the sign of the EXACT in-circle determinant (rational arithmetic on the
    float32 coordinates) of p, q and the two apexes A gives the edge; > 0: the edge is not locally Delaunay"""
    from fractions import Fraction as Fr
    P = lambda i: (Fr(float(V[i, 0])), Fr(float(V[i, 1])))
    nb = lambda i: set(col_a[rp_a[i]:rp_a[i + 1]].tolist())
    out = []
    for p, q in only_a:
        (px, py), (qx, qy) = P(p), P(q)
        side = lambda r: (qx - px) * (P(r)[1] - py) - (qy - py) * (P(r)[0] - px)
        com = nb(p) & nb(q)
        left, right = [r for r in com if side(r) > 0], [r for r in com if side(r) < 0]
        worst = None
        for r in left:
            for t in right:
                (dx, dy) = P(t)
                a, b, c = [(x - dx, y - dy) for x, y in (P(p), P(q), P(r))]
                det = ((a[0] ** 2 + a[1] ** 2) * (b[0] * c[1] - c[0] * b[1]) - (b[0] ** 2 + b[1] ** 2) * (a[0] * c[1] - c[0] * a[1])
                       + (c[0] ** 2 + c[1] ** 2) * (a[0] * b[1] - b[0] * a[1]))
                # p, q, r counter-clockwise  =>  det > 0 iff t is strictly inside their circle
                worst = det if worst is None else max(worst, det)
        out.append(worst)
    return out


def test_far_outliers_do_not_decide_the_grid(pkg, ctx):
    """badly triangulated matches far from everything else: the search grid follows mean +- 4 sigma and the mesh stays the
    Delaunay triangulation.  With points 3000 sigma away Qhull itself (floating-point lifting, 'Qbb' scaling) flips a handful of
    nearly co-circular quads the wrong way; wherever the two meshes differ the EXACT in-circle test decides: every edge only
    the device has is locally Delaunay, every edge only Qhull has is not."""
    V = _points("normal", 30000, 12)
    V[:5, :2] = np.array([[40.0, 3.0], [-25.0, -60.0], [0.5, 90.0], [70.0, -70.0], [-90.0, 0.1]], np.float32)
    n = len(V)
    g = ograph.delaunay_graph(V.astype(np.float64))
    rowptr, col, w, area, ntri, ns = ctx.delaunay_graph(V)
    assert ntri == g.n_triangles and len(col) == len(g.col)
    edges_of = lambda rp, c: set(zip(np.repeat(np.arange(n), np.diff(rp)).tolist(), np.asarray(c).tolist()))
    mine, qh = edges_of(rowptr, col), edges_of(g.rowptr, g.col)
    only_mine = sorted(e for e in mine - qh if e[0] < e[1])
    only_qh = sorted(e for e in qh - mine if e[0] < e[1])
    assert len(only_mine) == len(only_qh) <= 1e-3 * n                                      # diagonal flips of single quads
    assert all(d is not None and d < 0 for d in _exact_flip_verdicts(V, rowptr, col, only_mine))
    assert all(d is not None and d > 0 for d in _exact_flip_verdicts(V, g.rowptr, g.col, only_qh))
    # the same cloud without the outliers' influence on Qhull's scaling: bit-exact as everywhere else
    W = V[5:]
    g2 = ograph.delaunay_graph(W.astype(np.float64))
    rowptr2, col2, *_ = ctx.delaunay_graph(W)
    assert np.array_equal(rowptr2, g2.rowptr) and np.array_equal(col2, g2.col)


def test_refinement_on_the_gpu_built_mesh_matches_the_oracle(pkg, ctx):
    """config 1's flow with the graph built on the device: same LM trace as the oracle on its Qhull mesh"""
    from oracle import lm
    sc = scenes.sheet_scene(1500, seed=7)
    p, keep = scenes.problem_from_scene(sc, "delaunay", 8)
    w = edges.Weights(rep=1.0, arap=3.0e3, depth_sigma=0.003)
    pair = pkg.make_pair(p.cam1, p.cam2, p.T1, p.T2)
    ctx.problem_upload(pair, p.X1, p.X2, p.uv1, p.uv2, p.d1, p.d2, p.inv_sigma2_1, p.inv_sigma2_2, scale1=p.s1, scale2=p.s2, Tg7=p.Tg.as7())
    area, ntri, E = ctx.set_graph_delaunay()
    assert ntri == p.graph.n_triangles and E == len(p.graph.col) and area == pytest.approx(p.graph.area, rel=1e-12)
    ctx.compute_rotations()
    ctx.set_solver(1)
    ctx.set_pcg(rtol=1e-12, max_iters=20000, check_every=64)
    recs, st = ctx.optimize(pkg.make_weights(w.rep, w.arap, w.depth_sigma), 4)
    ost, otr = lm.optimize(p, w, 4)
    assert [r.trials for r in recs] == list(otr.trials)
    for r, c in zip(recs, otr.chi2):
        assert r.chi2_before == pytest.approx(c, rel=1e-5)
    assert st.final_chi2 == pytest.approx(otr.final_chi2, rel=1e-5)


def test_delaunay_1m_timing(pkg, ctx):
    """size of the bench workload: build time and consistency (symmetry is validated by the library's own set-up kernel)"""
    import time
    V = _points("uniform", 1_000_000, 9)
    ctx.delaunay_graph(V[:1000])
    t0 = time.perf_counter()
    rowptr, col, w, area, ntri, nsecond = ctx.delaunay_graph(V)
    ms = (time.perf_counter() - t0) * 1e3
    n = len(V)
    assert ntri > 1.9 * n and nsecond < 40 * np.sqrt(n)
    row = np.repeat(np.arange(n), np.diff(rowptr))
    key = row.astype(np.int64) * n + col
    rev = col.astype(np.int64) * n + row
    assert np.array_equal(np.sort(key), np.sort(rev))                                      # symmetric
    assert np.array_equal(w[np.argsort(key)], w[np.argsort(rev)])                          # ... with the same bits both ways
    assert ntri == len(col) // 2 - n + 1                                                   # Euler, triangulated disc: T = E - V + 1
    print(f"[delaunay 1M] {ms:.1f} ms incl. upload/download; {ntri} triangles, {len(col) // 2} edges, {nsecond} second-pass cells")
