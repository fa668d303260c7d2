#!/usr/bin/env python
"""Golden LM traces of the FULL-SIZE configurations (BASELINE.json configs 3 and 4, 1M correspondences), produced by the
plain-C oracle (oracle/c/dsc_oracle.c, PCG rtol 1e-12) on the CPU of the build container.  The GPU test
tests/test_gpu_fullsize.py rebuilds the same frame pair on the B200 (triangulation + k-NN graph on the device) and
compares cost, gradient, operator and the per-iteration trace against these files: the oracle needs tens of minutes for
one such solve, the GPU box must not wait for it.

  python tests/golden/make_fullsize_traces.py drunkard 1000000 8 30      -> fullsize_drunkard_1000000_k8.json
  python tests/golden/make_fullsize_traces.py realcolon 1000000 16 6     -> fullsize_realcolon_1000000_k16.json
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

import fullsize  # noqa: E402
from oracle import edges, cport  # noqa: E402


def probe_vector(m):
    """deterministic, machine-independent test vector for the operator check"""
    k = np.arange(m, dtype=np.float64)
    return ((k * 0.6180339887498949) % 1.0) - 0.5


def main():
    workload, n, k, iters = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    threads = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    wl = fullsize.workloads()
    sc = wl.make_scene(workload, n, 0)
    t0 = time.time()
    p, idx = fullsize.oracle_problem(sc, n, k)
    w = edges.Weights(**sc["weights"])
    print(f"problem built in {time.time() - t0:.1f} s: n={p.n} E={p.graph.n_edges}", flush=True)
    cp = cport.CProblem(p, rotations=p.R)
    m = 8 + 6 * p.n
    x = probe_vector(m)
    b, hd, y0, chi = cport.debug_linearize(cp, w, 0.0, x)
    lam = 1e-5 * np.abs(hd).max()
    y = y0 + lam * x
    sel = np.arange(0, p.n, max(1, p.n // 250))
    rows = np.concatenate([np.arange(8), (8 + 6 * sel[:, None] + np.arange(6)[None, :]).reshape(-1)])
    out = dict(workload=workload, n=p.n, k=k, seed=0, directed_edges=int(p.graph.n_edges), fingerprint=fullsize.fingerprint(p),
               weights=sc["weights"], s1=p.s1, s2=p.s2, pcg_rtol=1e-12,
               chi2_initial=chi, b_norm=float(np.linalg.norm(b)), b_dot_probe=float(b @ x), hdiag_sum=float(hd.sum()),
               hdiag_max=float(np.abs(hd).max()), lambda_probe=lam, y_norm=float(np.linalg.norm(y)), y_dot_probe=float(y @ x),
               sample_rows=rows.tolist(), b_sample=b[rows].tolist(), hdiag_sample=hd[rows].tolist(), y_sample=y[rows].tolist())
    t0 = time.time()
    tr = fullsize.run_c_oracle(p, w, iters, pcg_rtol=1e-12, pcg_max=200000, threads=threads)
    dt = time.time() - t0
    print(f"{len(tr['chi2'])} LM iterations, {sum(tr['pcg_iters'])} PCG iterations in {dt:.0f} s", flush=True)
    out.update(chi2=tr["chi2"], lam=tr["lam"], trials=tr["trials"], pcg_iters=tr["pcg_iters"], final_chi2=tr["final_chi2"],
               chi2_parts_initial=tr["chi2_parts"], scales=tr["scales"], Tg=tr["Tg"], sample_points=sel.tolist(),
               X1_sample=tr["X1"][sel].tolist(), X2_sample=tr["X2"][sel].tolist(),
               X1_sum=float(tr["X1"].sum()), X2_sum=float(tr["X2"].sum()), oracle_seconds=dt, oracle_threads=cport.threads() if threads == 0 else threads)
    path = os.path.join(HERE, f"fullsize_{workload}_{p.n}_k{k}.json")
    json.dump(out, open(path, "w"))
    print("wrote", path, flush=True)


if __name__ == "__main__":
    main()
