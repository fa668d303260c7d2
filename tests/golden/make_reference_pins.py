#!/usr/bin/env python
"""Builds tests/golden/reference_pins.{json,npz}: numbers HELD BY THE REFERENCE that the hot path must reproduce.

Run in the build container (needs /root/reference).  For every (original, moved) pair of Data/SinteticDataBase/** and
every seed location, the `INITIAL MEASUREMENTS` block of the matching historic log
Data/Experiments/{ARAP,ARAP_NoGlobal}/{InRays,TwoPoints,FarPoints}/<case>/Experiment.txt records, right after
triangulation and before any optimisation:
    C1 / C2 standard desv   calculatePixelsStandDev (Modules/Utils/Geometry.cc:370-498) of the triangulated points
    Av. error, RMSE         measureSimAbsoluteMapErrors (Modules/Utils/Measurements.cc:8-98), mm
and Data/SinteticDataBase/<case>/Experiment.txt records C1 / C2 standard desv with the ground-truth points inserted.
They depend on: std::default_random_engine + std::normal_distribution<float> (SLAM.cc:281-309), lookAt / setCameraPoses
(:223-235,340-351), PinHole::project / unproject (the logs predate the KannalaBrandt8 switch of Settings.cc:43-51: their
Test.yaml carries pin-hole intrinsics and only the PinHole model reproduces them), roundToDecimals, unproject,
triangulateNRSLAM with the three seed locations (Geometry.cc:103-153), isValidParallax (Mapping.cc:351-364) and the two
metrics.  The oracle reproduces them to the 6 printed digits; cases whose log was evidently written from other input
files (a third of the tree: the same numbers appear under several seed directories) are skipped.
The FarPoints logs under Experiments/ARAP/ were written by an older seed placement and are not used; ARAP_NoGlobal,
Elastic and HyperElasticOdgen hold the FarPoints numbers of the placement at HEAD.
"""
import glob
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/Data/"

from oracle import scenes, camera, metrics  # noqa: E402
from oracle.triangulate import triangulate_pairs, GATE_SIM  # noqa: E402

LOGS = (("InRays", "ARAP"), ("TwoPoints", "ARAP"), ("FarPoints", "ARAP_NoGlobal"))
PER_DEPTH = 6                 # pairs kept per depth directory


def yaml_key(txt, name):
    return float([t for t in txt.splitlines() if t.startswith(name)][0].split(":")[1])


def block(path, header):
    """C1/C2 standard desv, Av. error, RMSE of the block starting at `header` (es_ES decimal commas)."""
    t = open(path).read().replace(",", ".").splitlines()
    if header:
        i0 = [k for k, l in enumerate(t) if l.startswith(header)]
        if not i0:
            return None
        t = t[i0[0]:]

    def g(key):
        return float([l for l in t if l.startswith(key)][0].split(":")[1])
    return dict(sigma_c1=g("C1 standard desv"), sigma_c2=g("C2 standard desv"), av_error=g("Av. error"), rmse=g("RMSE"))


def oracle_numbers(o, m, C1, C2, loc):
    fe = scenes.simulation_frontend(o, m, C1, C2, model=camera.PINHOLE)
    cam = (camera.PINHOLE, fe["cam"])
    X1, X2, valid, _ = triangulate_pairs(fe["uv1"], fe["uv2"], cam, cam, fe["T1"], fe["T2"], "NRSLAM", loc, GATE_SIM, 0.9998)
    v = valid
    _, av, rmse = metrics.sim_absolute_map_errors(X1[v], X2[v], o[v], m[v])
    return dict(sigma_c1=scenes.pixel_sigma(cam, fe["T1"], X1[v], fe["uv1"][v]),
                sigma_c2=scenes.pixel_sigma(cam, fe["T2"], X2[v], fe["uv2"][v]), av_error=av, rmse=rmse, n=int(v.sum()))


def main():
    arrs, meta = {}, []
    kept = {}
    for case in sorted(glob.glob(REF + "SinteticDataBase/*/*/*/[0-9]")):
        rel = case[len(REF + "SinteticDataBase/"):]
        depth = rel.split("/")[0]
        if kept.get(depth, 0) >= PER_DEPTH:
            continue
        ytxt = open(os.path.dirname(case) + "/Test.yaml").read()
        C1 = [yaml_key(ytxt, "Camera.FirstPose." + a) for a in "xyz"]
        C2 = [yaml_key(ytxt, "Camera.SecondPose." + a) for a in "xyz"]
        o = scenes.load_points_csv(case + "/original_points.csv")
        m = scenes.load_points_csv(case + "/moved_points.csv")
        entry = dict(case=rel, C1=C1, C2=C2, key=len(meta), logs={})
        db = block(case + "/Experiment.txt", None)
        ok = True
        for loc, model in LOGS:
            lp = REF + f"Experiments/{model}/{loc}/{rel}/Experiment.txt"
            lg = block(lp, "INITIAL") if os.path.exists(lp) else None
            if lg is None:
                ok = False
                break
            mine = oracle_numbers(o, m, C1, C2, loc)
            # a log written from these very input files agrees to the printed digits; anything else is another pair
            if abs(mine["av_error"] - lg["av_error"]) > 2e-5 * lg["av_error"] or abs(mine["rmse"] - lg["rmse"]) > 2e-5 * lg["rmse"]:
                ok = False
                break
            lg["source"] = f"Data/Experiments/{model}/{loc}/{rel}/Experiment.txt (INITIAL MEASUREMENTS)"
            lg["n_mapped"] = mine["n"]
            entry["logs"][loc] = lg
        if not ok:
            continue
        entry["database"] = dict(sigma_c1=db["sigma_c1"], sigma_c2=db["sigma_c2"],
                                 source=f"Data/SinteticDataBase/{rel}/Experiment.txt")
        arrs[f"o{entry['key']}"] = o
        arrs[f"m{entry['key']}"] = m
        meta.append(entry)
        kept[depth] = kept.get(depth, 0) + 1
    np.savez_compressed(os.path.join(HERE, "reference_pins.npz"), **arrs)
    json.dump(meta, open(os.path.join(HERE, "reference_pins.json"), "w"), indent=1)
    print(f"{len(meta)} pairs x {len(LOGS)} seed locations pinned:", {k: v for k, v in kept.items()})


if __name__ == "__main__":
    main()
