#!/usr/bin/env python
"""How far is the gain ratio rho at a loose PCG tolerance from its converged value?  Runs LM on a set of synthetic
problems with DSC_EARLY_LOG=1 and margins so large that nothing is rejected, and reports, per loose level, the smallest
loose rho over the trials whose converged rho was positive (= steps that must never be rejected early).
usage: DSC_EARLY_LOG=1 python profiles/early_reject_study.py 2> gpurun_out/early.log; python profiles/early_reject_study.py --parse gpurun_out/early.log"""
import os
import re
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

LEVELS = [3e-2, 1e-2, 1e-3, 1e-4]


def parse(path):
    pat = re.compile(r"\[dsc early\] it (\d+) trial (\d+) level (-?\d+) tol (\S+) its (\d+) rho (\S+)")
    worst = {}
    cur = {}
    n_acc = n_rej = 0
    rej_max = {}
    for line in open(path):
        if line.startswith("== "):
            print(line.strip()); continue
        m = pat.search(line)
        if not m:
            continue
        level, rho = int(m.group(3)), float(m.group(6))
        if level >= 0:
            cur[level] = rho
        else:
            if rho > 0:
                n_acc += 1
                for l, r in cur.items():
                    worst[l] = min(worst.get(l, 1e300), r)
            else:
                n_rej += 1
                for l, r in cur.items():
                    rej_max.setdefault(l, []).append((r, rho))
            cur = {}
    print(f"accepted trials {n_acc}, rejected trials {n_rej}")
    for l in sorted(worst):
        caught = {mg: sum(1 for r, _ in rej_max.get(l, []) if r < -mg) for mg in (0.5, 1.0, 2.0, 3.0)}
        print(f"level {l} (rtol {LEVELS[l]:.0e}): smallest loose rho of an accepted step {worst[l]:+.4f}; rejected trials caught at margin: {caught}")


def main():
    import numpy as np
    import bench
    import __graft_entry__ as g
    pkg = g.package()
    ctx = pkg.Context(0)
    cases = []
    for wl, n, k, seed in [("drunkard", 20000, 8, 1), ("drunkard", 100000, 8, 2), ("drunkard", 300000, 8, 3), ("drunkard", 100000, 16, 4),
                           ("realcolon", 100000, 16, 5), ("realcolon", 200000, 8, 6), ("sheet", 50000, 8, 7), ("sheet", 100000, 8, 8),
                           ("sheet", 20000, 16, 9), ("drunkard", 1000000, 8, 0)]:
        cases.append((wl, n, k, seed))
    for wl, n, k, seed in cases:
        sys.argv = [sys.argv[0], "--workload", wl, "--points", str(n), "--k", str(k)]
        args = bench.parse()
        sc = bench.make_scene(pkg, args, seed)
        prob = bench.prepare(pkg, ctx, sc, args)
        for scale in (1.0, 0.01, 100.0):
            bench.upload(ctx, prob)
            wd = dict(sc["weights"])
            wd["arap"] = wd["arap"] * scale
            w = pkg.make_weights(**wd)
            ctx.set_pcg(rtol=args.pcg_rtol, max_iters=args.pcg_max_iters, check_every=64)
            ctx.set_early_reject(LEVELS, [1e30] * len(LEVELS))
            sys.stderr.write(f"== {wl} n={n} k={k} arap x{scale}\n"); sys.stderr.flush()
            recs, st = ctx.optimize(w, 12)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--parse":
        parse(sys.argv[2])
    else:
        main()
