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
    assert ntri > 1.9 * n and len(col) == 2 * (ntri + (len(col) // 2 - ntri)) and nsecond < 40 * np.sqrt(n)
    hull = len(col) // 2 - (3 * ntri - len(col) // 2)        # Euler: E = 3T + h - ... for a triangulated disc: h = 3V - 3 - E
    print(f"[delaunay 1M] {ms:.1f} ms incl. upload/download; {ntri} triangles, {len(col) // 2} edges, {nsecond} second-pass cells")
