#!/usr/bin/env python
"""Regenerates the fixtures under tests/golden/ (run in the build container, where /root/reference exists).

  config1_points.npz      Data/original_points.csv, Data/moved_points.csv (the 120-point input of config 1)
  debug_hessian.json      structure facts of /root/reference/debug.txt (g2o's Hessian dump of an 861-correspondence pair)
  sigma_database.npz/json three (original, moved) pairs of Data/SinteticDataBase + the "C1/C2 standard desv" their
                          Experiment.txt records
  config1_trace.json      the ORACLE's LM trace on config 1 (pins the oracle against accidental edits; the reference
                          itself cannot be run, see oracle/__init__.py)
"""
import collections
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/"

from oracle import scenes, lm, edges  # noqa: E402


def main():
    o = scenes.load_points_csv(REF + "Data/original_points.csv")
    m = scenes.load_points_csv(REF + "Data/moved_points.csv")
    np.savez(os.path.join(HERE, "config1_points.npz"), original=o, moved=m)

    # --- debug.txt
    cols = collections.Counter()
    rows_in_col = collections.defaultdict(set)
    n = 0
    for line in open(REF + "debug.txt"):
        if line.startswith("# rows:"):
            n = int(line.split(":")[1])
        if line.startswith("#") or not line.strip():
            continue
        r, c, _ = line.split()
        r, c = int(r), int(c)
        cols[c] += 1
        if c <= 8:
            rows_in_col[c].add(r)
    facts = dict(rows=n, correspondences=(n - 8) // 6, col_nnz_first8=[cols[c] for c in range(1, 9)],
                 lower_triangle=all(min(rows_in_col[c]) >= c for c in range(1, 9)),
                 t_scale_coupling=sorted(int(r) for c in range(1, 7) for r in rows_in_col[c] if r in (7, 8)),
                 scale_scale_coupling=(8 in rows_in_col[7]))
    json.dump(facts, open(os.path.join(HERE, "debug_hessian.json"), "w"), indent=1)

    # --- synthetic database sigma
    db = REF + "Data/SinteticDataBase/"
    picks = ["80cm Depth/Planar/2_5 mm rigid/3", "20cm Depth/Planar/2_5 mm rigid/2", "150cm Depth/Planar/10 mm gaussian/1"]
    arrs, meta = {}, []
    for k, pth in enumerate(picks):
        d = db + pth + "/"
        if not os.path.exists(d + "Experiment.txt"):
            continue
        txt = open(d + "Experiment.txt").read().replace(",", ".").splitlines()
        s1 = float([t for t in txt if t.startswith("C1 standard desv")][0].split(":")[1])
        s2 = float([t for t in txt if t.startswith("C2 standard desv")][0].split(":")[1])
        yaml = open(os.path.dirname(os.path.dirname(d[:-1])) + "/Test.yaml").read() if os.path.exists(os.path.dirname(os.path.dirname(d[:-1])) + "/Test.yaml") else open(os.path.dirname(d[:-1]) + "/Test.yaml").read()
        def key(name):
            return float([t for t in yaml.splitlines() if t.startswith(name)][0].split(":")[1])
        C1 = [key("Camera.FirstPose.x"), key("Camera.FirstPose.y"), key("Camera.FirstPose.z")]
        C2 = [key("Camera.SecondPose.x"), key("Camera.SecondPose.y"), key("Camera.SecondPose.z")]
        arrs[f"o{k}"] = scenes.load_points_csv(d + "original_points.csv")
        arrs[f"m{k}"] = scenes.load_points_csv(d + "moved_points.csv")
        meta.append(dict(case=pth, C1=C1, C2=C2, sigma_c1=s1, sigma_c2=s2, key=k))
    np.savez(os.path.join(HERE, "sigma_database.npz"), **arrs)
    json.dump(meta, open(os.path.join(HERE, "sigma_database.json"), "w"), indent=1)

    # --- oracle trace on config 1
    p, w, fe, keep = scenes.config1(REF + "Data/original_points.csv", REF + "Data/moved_points.csv")
    st, tr = lm.optimize(p, w, 25)
    json.dump(dict(n=p.n, n_edges=int(p.graph.n_edges), n_triangles=int(p.graph.n_triangles), area=p.graph.area,
                   s1=p.s1, s2=p.s2, chi2=tr.chi2, lam=tr.lam, trials=tr.trials, final_chi2=tr.final_chi2,
                   X1_sum=float(st.X1.sum()), X2_sum=float(st.X2.sum()), s1_final=st.s1, s2_final=st.s2,
                   Tg=st.Tg.as7().tolist(), uv1_sum=float(fe["uv1"].astype(np.float64).sum()),
                   d1_sum=float(fe["d1"].astype(np.float64).sum())),
              open(os.path.join(HERE, "config1_trace.json"), "w"), indent=1)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
