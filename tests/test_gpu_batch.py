"""The batched path (include/dsc.h dsc_batch_*: BASELINE.json configs[4]): many frame pairs, ONE kernel launch, the whole
Levenberg-Marquardt loop of a pair inside a thread-block cluster.  Checked against the direct-solve oracle (north_star's
1e-5 bar) and against the single-pair path of the same library (same algorithm, different partial-sum partition)."""
import numpy as np
import pytest

from oracle import scenes, edges, lm

pytestmark = pytest.mark.gpu


def _problem_dict(pkg, p):
    g = p.graph
    return dict(pair=pkg.make_pair(p.cam1, p.cam2, p.T1, p.T2), X1=p.X1, X2=p.X2, uv1=p.uv1, uv2=p.uv2, d1=p.d1, d2=p.d2,
                s1=p.s1, s2=p.s2, rowptr=g.rowptr, col=g.col, w=g.w, area=g.area, ntri=g.n_triangles, Tg7=p.Tg.as7(),
                isg1=p.inv_sigma2_1, isg2=p.inv_sigma2_2)


def _single(pkg, p, w, iters, early=False):
    with pkg.Context(0) as c:
        pair = pkg.make_pair(p.cam1, p.cam2, p.T1, p.T2)
        c.problem_upload(pair, p.X1, p.X2, p.uv1, p.uv2, p.d1, p.d2, p.inv_sigma2_1, p.inv_sigma2_2, scale1=p.s1, scale2=p.s2,
                         Tg7=p.Tg.as7())
        g = p.graph
        c.set_graph(g.rowptr, g.col, g.w, g.area, g.n_triangles, 1)
        c.compute_rotations()
        c.set_solver(1)
        c.set_pcg(rtol=1e-12, max_iters=20000, check_every=64)
        if early:
            c.set_early_reject((1e-3, 1e-4), (1.0, 0.5))
        recs, st = c.optimize(w, iters)
        return recs, st, c.download()


def _mixed_problems():
    out = []
    for n, seed, kind, k, tube in ((300, 1, "knn", 6, False), (1237, 2, "knn", 8, True), (2500, 3, "knn", 8, False),
                                  (640, 4, "delaunay", 8, False), (10000, 5, "knn", 8, False)):
        if tube:
            sc = scenes.tube_scene(n, seed=seed, cam=scenes.REALCOLON_CAM, depth_sigma=0.0003, scales=(1.3, 0.8))
            p, _ = scenes.problem_from_scene(sc, kind, k, min_cos=0.99999)
            w = edges.Weights(rep=1.0, arap=1.0e7, depth_sigma=0.0003)
        else:
            sc = scenes.sheet_scene(n, seed=seed)
            p, _ = scenes.problem_from_scene(sc, kind, k)
            w = edges.Weights(rep=1.0, arap=2.0e5 if seed % 2 else 50.0, depth_sigma=0.003)
        out.append((p, w))
    return out


def test_batch_of_mixed_pairs_against_oracle_and_single_path(pkg):
    probs = _mixed_problems()
    iters = 5
    with pkg.Batch(0) as b:
        b.upload([_problem_dict(pkg, p) for p, _ in probs])
        b.set_pcg(rtol=1e-12, max_iters=20000)
        recs, stats, ms = b.optimize([pkg.make_weights(w.rep, w.arap, w.depth_sigma) for _, w in probs], iters)
        outs = b.download()
        info = b.size()
    assert info["problems"] == len(probs) and info["clusters"] >= 1 and ms > 0
    assert sum(s.kernel_launches for s in stats) == 1                       # the whole batch is ONE launch
    for k, (p, w) in enumerate(probs):
        r, st, out = recs[k], stats[k], outs[k]
        assert st.iterations == iters and st.pcg_unconverged == 0
        # the single-pair path of the same library: identical decisions, values to rounding
        r1, s1, o1 = _single(pkg, p, pkg.make_weights(w.rep, w.arap, w.depth_sigma), iters)
        assert [x.trials for x in r] == [x.trials for x in r1]
        for a, c in zip(r, r1):
            # (two PCG solves to rtol 1e-12 with different partial-sum partitions: the costs agree to ~1e-9)
            assert a.chi2_before == pytest.approx(c.chi2_before, rel=1e-7) and a.chi2_after == pytest.approx(c.chi2_after, rel=1e-7)
            assert a.lam == pytest.approx(c.lam, rel=1e-6)
        assert st.final_chi2 == pytest.approx(s1.final_chi2, rel=1e-7)
        scale = np.abs(o1["X1"]).max()
        assert np.abs(out["X1"].astype(np.float64) - o1["X1"]).max() <= 1e-6 * scale
        assert out["update"] == pytest.approx(o1["update"], rel=1e-5)
        # the direct-solve oracle (north_star bar) on the pairs it solves in seconds
        if p.n <= 2500:
            ost, otr = lm.optimize(p, w, iters)
            for a, c, q in zip(r, otr.chi2, otr.trials):
                assert a.chi2_before == pytest.approx(c, rel=1e-5) and a.trials == q
            assert st.final_chi2 == pytest.approx(otr.final_chi2, rel=1e-5)
            n1, n2, upd = lm.write_back(p, ost)
            assert np.abs(out["X1"] - n1).max() <= 1e-5 * scale and np.abs(out["X2"] - n2).max() <= 1e-5 * scale
            assert out["scales"][0] == pytest.approx(ost.s1, rel=1e-5) and out["scales"][1] == pytest.approx(ost.s2, rel=1e-5)
            np.testing.assert_allclose(out["Tg"], ost.Tg.as7(), atol=1e-5)


def test_batch_early_rejection_keeps_the_trace_and_reset_reruns(pkg):
    """config 5's setting: 10k-correspondence sheets, early rejection of bad trials inside the kernel"""
    ps = []
    for seed in range(6):
        sc = scenes.sheet_scene(4000 + 500 * seed, seed=30 + seed)
        p, _ = scenes.problem_from_scene(sc, "knn", 8)
        ps.append(p)
    w = pkg.make_weights(1.0, 2.0e5, 0.003)
    with pkg.Batch(0) as b:
        b.upload([_problem_dict(pkg, p) for p in ps])
        b.set_pcg(rtol=1e-10, max_iters=6000)
        r0, s0, ms0 = b.optimize(w, 8)
        o0 = b.download()
        b.reset_state()
        b.set_early_reject((1e-3, 1e-4), (1.0, 0.5))
        r1, s1, ms1 = b.optimize(w, 8)
        o1 = b.download()
        b.set_early_reject((), ())
        b.reset_state()
        r2, s2, ms2 = b.optimize(w, 8)
        o2 = b.download()
    assert sum(s.early_rejects for s in s1) > 0 and sum(s.total_pcg_iters for s in s1) < sum(s.total_pcg_iters for s in s0)
    for k in range(len(ps)):
        assert [x.trials for x in r1[k]] == [x.trials for x in r0[k]]
        assert [x.chi2_after for x in r1[k]] == [x.chi2_after for x in r0[k]]            # accepted steps bit-identical
        assert np.array_equal(o0[k]["X1"], o1[k]["X1"]) and np.array_equal(o0[k]["X2"], o1[k]["X2"])
        # same launch configuration, same state: the re-run is bit-identical (deterministic reductions)
        assert [x.chi2_after for x in r2[k]] == [x.chi2_after for x in r0[k]]
        assert np.array_equal(o0[k]["X1"], o2[k]["X1"])


def test_batch_edge_cases(pkg):
    """an empty batch, an empty pair inside a batch, more pairs than clusters (the queue), one shared dsc_weights"""
    sc = scenes.sheet_scene(900, seed=77)
    p, _ = scenes.problem_from_scene(sc, "knn", 8)
    w = pkg.make_weights(1.0, 50.0, 0.003)
    with pkg.Batch(0) as b:
        b.upload([])
        recs, stats, ms = b.optimize(w, 3)
        assert recs == [] and b.download() == []
        empty = dict(pair=pkg.make_pair(p.cam1, p.cam2, p.T1, p.T2), X1=np.zeros((0, 3)), X2=np.zeros((0, 3)), uv1=np.zeros((0, 2)),
                     uv2=np.zeros((0, 2)), d1=np.zeros(0), d2=np.zeros(0), s1=1.0, s2=1.0, rowptr=np.zeros(1, np.int32),
                     col=np.zeros(0, np.int32), w=np.zeros(0), area=1.0, ntri=0)
        many = [_problem_dict(pkg, p)] * 40
        many.insert(7, empty)
        for d in many:
            d.pop("isg1", None); d.pop("isg2", None)
        b.upload(many)
        b.set_pcg(rtol=1e-12, max_iters=20000)
        recs, stats, ms = b.optimize(w, 3)
        outs = b.download()
        assert stats[7].iterations == 0 and outs[7]["X1"].shape == (0, 3)
        ref = [x.chi2_after for x in recs[0]]
        for k in range(len(many)):
            if k != 7:
                assert [x.chi2_after for x in recs[k]] == ref                              # identical pairs, identical results
                assert np.array_equal(outs[k]["X1"], outs[0]["X1"])
    r1, s1, o1 = _single(pkg, p, w, 3)
    assert ref == pytest.approx([x.chi2_after for x in r1], rel=1e-7)


def test_weight_search_candidates_refined_together_match_the_oracle(pkg):
    """the outer objective (nloptOptimization.cc:5-37) of the weights a Nelder-Mead search visits, (a) by the
    oracle one refinement at a time, (b) on the device with the K replicas of the pair resident in one dsc_batch, as many of them
    refined per launch as there are candidates (dsc_batch_set_active / dsc_batch_pixel_sigma).  Same evaluations, same order."""
    from oracle import outer
    sc = scenes.sheet_scene(400, seed=7)
    p, _ = scenes.problem_from_scene(sc, "delaunay", 8)
    iters, sigma_d = 4, 0.003
    x0, lb, ub = [1.0, 50.0, 2.0e5], [1.0, 50.0, 1e-5], [1.0, 50.0, 1e7]          # Simulation.yaml:88-93: a 1-D search over arap
    fo = lambda y: outer.outer_objective(p, y[0], y[1], y[2], sigma_d, iters)
    xo, fbest_o, log_o = outer.nelder_mead(fo, x0, lb, ub, 0.15, 0.15, 7)
    K = 4
    with pkg.Batch(0) as b:
        b.upload([_problem_dict(pkg, p) for _ in range(K)])
        b.set_pcg(rtol=1e-12, max_iters=20000)

        def many(cands):
            b.set_active(len(cands))
            b.reset_state()
            b.optimize([pkg.make_weights(y[0], y[2], sigma_d, glob=y[1]) for y in cands], iters)
            s = b.pixel_sigma()
            return [float(np.log(s[k, 0]) ** 2 + np.log(s[k, 1]) ** 2) for k in range(len(cands))]
        # the weights the oracle's search visited, four at a time: one launch per group
        xs = [e[0] for e in log_o]
        fg = []
        for at in range(0, len(xs), K):
            fg += many(xs[at:at + K])
        np.testing.assert_allclose(fg, [e[1] for e in log_o], rtol=2e-5)
        # and the search itself on device values: the same walk
        xg, fbest_g, log_g = outer.nelder_mead(lambda y: many([y])[0], x0, lb, ub, 0.15, 0.15, 7)
        assert len(log_g) == len(log_o)
        np.testing.assert_allclose([e[0] for e in log_g], xs, rtol=1e-12)
        assert xg == pytest.approx(xo, rel=1e-12) and fbest_g == pytest.approx(fbest_o, rel=2e-5)
        # a subset launch leaves the other replicas alone
        b.set_active(-1)
        b.reset_state()
        b.set_active(1)
        b.optimize([pkg.make_weights(1.0, 2.0e5, sigma_d, glob=50.0)], iters)
        b.set_active(-1)
        s = b.pixel_sigma()
        assert s[1, 0] == s[2, 0] == s[3, 0] and s[0, 0] != s[1, 0]
        assert s[1, 0] == pytest.approx(scenes.pixel_sigma(p.cam1, p.T1, p.X1, p.uv1), rel=1e-9)
