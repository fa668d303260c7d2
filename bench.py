#!/usr/bin/env python
"""bench.py -- the hot path on BASELINE.json's metric: non-rigid LM iterations/s (and triangulated
points/s, HBM GB/s of the dominant kernel) on a synthetic deformable two-view pair.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--points 1000000] [--k 8]

A step = one arapOptimization call (triangulated pair -> `lm_iters` Levenberg-Marquardt iterations)
on the workload named in config.workload.  `value` times the refinement with all inputs resident in
HBM (CUDA events on the library's stream); `e2e` times the same step through the C ABI from host
buffers (uv -> triangulate -> host; problem + graph upload; rotations; LM; download).  N > 1 runs one
independent frame-pair problem per rank (no data-path collective): weak scaling.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--points", dest="n", type=int, default=1_000_000, help="correspondences per frame pair")
    ap.add_argument("--k", type=int, default=8)
    ap.add_argument("--workload", default="drunkard", choices=["drunkard", "realcolon", "sheet", "batch"])
    ap.add_argument("--problems", type=int, default=0, help="batch workload: frame-pair problems per GPU (config 5 has 4096 in total)")
    ap.add_argument("--total-problems", type=int, default=0, help="batch workload: strong scaling over this many pairs in total (config 5: 4096)")
    ap.add_argument("--config5-problems", type=int, default=64, help="frame pairs per GPU of the config-5 sub-record of the default line (0 = skip)")
    ap.add_argument("--lm-iters", type=int, default=0, help="0 = the config's own count")
    ap.add_argument("--pcg-rtol", type=float, default=1e-10)
    ap.add_argument("--pcg-max-iters", type=int, default=6000)
    ap.add_argument("--early-rtol", type=float, nargs="*", default=[1e-3, 1e-4], help="loose tolerances of the early-reject check (none = off)")
    ap.add_argument("--early-margin", type=float, nargs="*", default=[1.0, 0.5])
    ap.add_argument("--cpu-budget", type=float, default=900.0, help="--impl reference: wall-time budget of the CPU step in seconds")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fp32", action="store_true", help="skip the fp32-mode sub-record")
    ap.add_argument("--no-point-sharded", action="store_true", help="N > 1: skip the one-pair-over-all-GPUs sub-record")
    ap.add_argument("--no-classic-ba", action="store_true", help="skip the classic bundle-adjustment sub-record (SURVEY 8f-4)")
    ap.add_argument("--ba-points", type=int, default=1_000_000, help="map points of the classic-BA sub-record (two views)")
    ap.add_argument("--no-shim-e2e", action="store_true", help="skip the arapOptimization-through-the-C++-shim sub-record")
    ap.add_argument("--shim-points", type=int, default=1_000_000, help="points of the shim sub-record's sheet scene")
    return ap.parse_args()


class _stdout_to_stderr:
    """keep the one-JSON-line contract: anything a library writes to fd 1 meanwhile goes to stderr"""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *a):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


# ---------------------------------------------------------------------------------------------- workload
def make_scene(pkg, args, seed):
    return pkg_workloads(pkg).make_scene(args.workload, args.n, seed)


def pkg_workloads(pkg):
    import importlib
    return importlib.import_module(pkg.__name__ + ".workloads")


def prepare(pkg, ctx, sc, args, pin=False):
    """Triangulate on the GPU, keep exactly n valid correspondences, build the k-NN graph (host, untimed)."""
    wl = pkg_workloads(pkg)
    cam = (0, sc["cam"])
    pair = pkg.make_pair(cam, cam, sc["T1"], sc["T2"])
    prm = ctx.tri_params("NRSLAM", "FarPoints", 1, sc["min_cos"])
    X1, X2, valid, cosp, nv = ctx.triangulate(pair, prm, sc["uv1"], sc["uv2"])
    idx = np.nonzero(valid)[0][: args.n]
    if len(idx) < args.n:
        raise RuntimeError(f"only {len(idx)} valid correspondences of {args.n} requested")
    prob = dict(pair=pair, prm=prm, X1=X1[idx], X2=X2[idx], uv1=sc["uv1"][idx], uv2=sc["uv2"][idx],
                d1=sc["d1"][idx].astype(np.float64), d2=sc["d2"][idx].astype(np.float64), area=sc["area"])
    t0 = time.perf_counter()
    rowptr, col, w = ctx.knn_graph(prob["X1"], args.k)          # symmetrised k-NN graph built on the GPU (dsc_knn_build)
    prob.update(rowptr=rowptr, col=col, w=w, ntri=2 * len(idx), graph_build_ms=(time.perf_counter() - t0) * 1e3)
    # initial depth scales (KeyFrame::setInitialDepthScaleInSimulationImages) from the kept points
    ctx.tri_upload(pair, prob["uv1"], prob["uv2"], prob["d1"].astype(np.float32), prob["d2"].astype(np.float32))
    ctx.tri_run(prm)
    prob["s1"], prob["s2"] = ctx.depth_scale_init(1), ctx.depth_scale_init(2)
    if pin:
        # The caller's buffers are page-locked ONCE (a front end keeps its key-point / map-point arrays across frames);
        # every step's uploads and downloads are then DMA straight from / into them -- inside the timed region.
        for key in ("X1", "X2", "uv1", "uv2", "d1", "d2", "rowptr", "col", "w"):
            prob[key] = np.ascontiguousarray(prob[key])
        pkg.pin_host(*[prob[key] for key in ("X1", "X2", "uv1", "uv2", "d1", "d2", "rowptr", "col", "w")])
        n = len(idx)
        prob["out_tri"] = dict(X1=np.empty((n, 3), np.float32), X2=np.empty((n, 3), np.float32), valid=np.empty(n, np.uint8),
                               cosp=np.empty(n, np.float32))
        prob["out_opt"] = dict(X1=np.empty((n, 3), np.float32), X2=np.empty((n, 3), np.float32))
        pkg.pin_host(*prob["out_tri"].values(), *prob["out_opt"].values())
    return prob


def upload(ctx, prob, flags=1):
    ctx.problem_upload(prob["pair"], prob["X1"], prob["X2"], prob["uv1"], prob["uv2"], prob["d1"], prob["d2"],
                       scale1=prob["s1"], scale2=prob["s2"])
    ctx.set_graph(prob["rowptr"], prob["col"], prob["w"], prob["area"], prob["ntri"], flags)
    ctx.compute_rotations()


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


# ---------------------------------------------------------------------------------------------- CPU arm
def reference_arm(args):
    """`--impl reference`: the reference's CPU algorithm (oracle port: the reference itself needs g2o / Eigen / Sophus /
    Qhull / Open3D and cannot be compiled here) on the SAME frame pair, size, iteration count and PCG tolerance as the
    GPU arm, with g2o's policy (every trial solved to the tolerance), all host threads.  A CPU step takes minutes, so
    ONE step is timed whatever --steps / --warmup say (echoed in the line; config.steps_timed tells)."""
    from oracle import bench_cpu
    res = bench_cpu.run_reference(args.workload, args.n, args.k, args.lm_iters, pcg_rtol=args.pcg_rtol, budget_s=args.cpu_budget)
    cfg = res["config"]
    cfg.update(parallelism="host cores only (rank 0)", trace=res["trace"])
    return dict(metric="non-rigid LM iterations/s at 1M correspondences", value=res["value"], unit="LM it/s",
                n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=res["ms_per_step"],
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64", data="synthetic",
                impl="reference", config=cfg, cpu_baseline=res["cpu_baseline"],
                e2e=dict(value=res["value"], unit="LM it/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)


# ---------------------------------------------------------------------------------------------- main
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload == "batch" and args.impl == "ours":
        return main_batch(args, rank, world, local)
    if args.workload == "batch":                       # the CPU arm of config 5 = one of its pairs
        args.workload, args.n = "sheet", (args.n if args.n != 1_000_000 else 10_000)

    if args.impl == "reference":
        if rank != 0:
            return 0
        with _stdout_to_stderr():
            line = reference_arm(args)
        print(json.dumps(line))
        return 0

    # torchrun pins OMP_NUM_THREADS=1; the library's host-side graph renumbering is OpenMP-parallel, so give every
    # rank its share of the host cores (must be set before libgomp is loaded with the library)
    os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 8) // max(1, world)))
    import __graft_entry__ as g
    pkg = g.package()
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        with _stdout_to_stderr():                      # NCCL prints its version banner on stdout
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
    ctx = pkg.Context(local)
    sc = make_scene(pkg, args, seed=rank)
    prob = prepare(pkg, ctx, sc, args, pin=True)
    lm_iters = args.lm_iters or sc["lm_iters"]
    w = pkg.make_weights(**sc["weights"])
    ctx.set_pcg(rtol=args.pcg_rtol, max_iters=args.pcg_max_iters, check_every=64)
    upload(ctx, prob)
    n, E = ctx.problem_size()

    def barrier():
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()
        ctx.synchronize()

    # ---- device-resident arm: inputs already in HBM, one arapOptimization per step.
    # The first warm-up runs every linear solve to the tight tolerance; the others (and the timed steps) reject
    # clearly bad LM trials at a loose tolerance (dsc_set_early_reject).  The two traces are compared and the
    # shortcut is only kept if they are identical, iteration by iteration.
    early = dict(rtol_loose=args.early_rtol, rho_margin=args.early_margin, used=False, trace_identical=None)
    ref_trace = None
    full_counts = None
    for k in range(args.warmup):
        ctx.reset_state()
        recs, st = ctx.optimize(w, lm_iters)
        tr = [(r.chi2_before, r.chi2_after, r.trials, r.accepted) for r in recs]
        if k == 0:
            ref_trace = tr
            early["lm_it_per_s_full_solves"] = st.iterations / (st.device_ms * 1e-3)   # first warm-up step, no shortcut
            full_counts = dict(lm_iters=st.iterations, trials=st.total_trials, pcg_iters=st.total_pcg_iters)
            if len(args.early_rtol) > 0 and args.warmup > 1:
                ctx.set_early_reject(args.early_rtol, args.early_margin)
                early["used"] = True
        elif early["used"] and early["trace_identical"] is None:
            early["trace_identical"] = (tr == ref_trace)
            early["early_rejects_per_step"] = st.early_rejects
            if not early["trace_identical"]:
                ctx.set_early_reject((), ())
                early["used"] = False
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = ctx.launch_count()
    barrier()
    t0 = time.perf_counter()
    dev_ms, its, pcg_its, trials, stats_last = 0.0, 0, 0, 0, None
    for _ in range(args.steps):
        ctx.reset_state()
        recs, st = ctx.optimize(w, lm_iters)
        dev_ms += st.device_ms
        its += st.iterations
        pcg_its += st.total_pcg_iters
        trials += st.total_trials
        stats_last = st
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    launches = ctx.launch_count() - launches0
    clocks = sampler.stop() if sampler else None

    # ---- end-to-end arm: host buffers in, host buffers out, every step
    h2d = (prob["uv1"].nbytes + prob["uv2"].nbytes) * 2 + prob["X1"].nbytes + prob["X2"].nbytes + prob["d1"].nbytes + \
        prob["d2"].nbytes + prob["rowptr"].nbytes + prob["col"].nbytes + prob["w"].nbytes
    d2h = 2 * 12 * n * 2 + 5 * n + 2 * 12 * n
    e2e_ms = 0.0
    for s in range(1 + args.steps):                       # first pass untimed
        barrier()
        t1 = time.perf_counter()
        ctx.triangulate(prob["pair"], prob["prm"], prob["uv1"], prob["uv2"], out=prob["out_tri"])
        upload(ctx, prob)
        ctx.optimize(w, lm_iters)
        out = ctx.download(doubles=False, out=prob["out_opt"])
        barrier()
        if s > 0:
            e2e_ms += (time.perf_counter() - t1) * 1e3

    # ---- fp32 mode (dsc_set_precision): float storage of what the PCG streams + fp64 iterative refinement; same steps,
    # device-resident, reported next to the fp64 headline with its distance from the fp64 result
    f32rec = None
    if not args.no_fp32:
        try:
            out64 = ctx.download(doubles=True)
            sig64 = ctx.pixel_sigma()
            chi64 = stats_last.final_chi2
            ctx.set_precision("f32")
            ms32, its32, pcg32, tr32 = 0.0, 0, 0, None
            for s in range(1 + args.steps):               # first pass untimed
                ctx.reset_state()
                recs, st = ctx.optimize(w, lm_iters)
                if s == 0:
                    tr32 = [(r.trials, r.accepted) for r in recs]
                else:
                    ms32 += st.device_ms; its32 += st.iterations; pcg32 += st.total_pcg_iters
            out32 = ctx.download(doubles=True)
            sig32 = ctx.pixel_sigma()
            k32 = ctx.profile_kernels(w, warm=3, reps=20)
            scale = float(np.abs(out64["X1d"]).max())
            f32rec = dict(lm_it_per_s=its32 / (ms32 * 1e-3), ms_per_step=ms32 / args.steps, pcg_iters_per_lm_iter=pcg32 / max(1, its32),
                          final_chi2_rel_diff=abs(st.final_chi2 - chi64) / chi64,
                          points_max_rel_diff=float(max(np.abs(out32["X1d"] - out64["X1d"]).max(), np.abs(out32["X2d"] - out64["X2d"]).max()) / scale),
                          pixel_sigma_abs_diff_px=[abs(a - b) for a, b in zip(sig32, sig64)],
                          lm_decisions_identical=(tr32 == [(t[2], t[3]) for t in ref_trace]) if ref_trace else None,
                          pcg_unconverged=st.pcg_unconverged,
                          kernels={k: dict(ms=v["ms"], gbs=v["bytes"] / (v["ms"] * 1e-3) / 1e9, bytes=v["bytes"]) for k, v in k32.items()
                                   if k in ("cg_spmv", "cg_update", "precond", "linearize")},
                          storage="Je/U/Minv/PCG vectors float; sums, state, cost double; fp64 residual refinement")
        except Exception as ex:
            f32rec = dict(error=str(ex))
        finally:
            ctx.set_precision("f64")

    # ---- kernel roofline (CUDA events on the library's stream) and triangulation throughput
    kern = ctx.profile_kernels(w, warm=3, reps=20)
    ctx.tri_upload(prob["pair"], sc["uv1"], sc["uv2"])
    tri = ctx.profile_triangulate(prob["prm"], warm=3, reps=20)
    n_tri = len(sc["uv1"])
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    cg_ms = kern["cg_spmv"]["ms"] + kern["cg_update"]["ms"]
    dom = "cg_spmv" if kern["cg_spmv"]["ms"] >= kern["cg_update"]["ms"] else "cg_update"
    ach = kern[dom]["bytes"] / (kern[dom]["ms"] * 1e-3) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(f"{args.workload}_{n}_k{args.k}", {}).get(dom)
    except Exception:
        pass
    roof = dict(bound="hbm", kernel=dom, achieved=ach, peak=peak, unit="GB/s", frac=ach / peak, traffic=traffic,
                algorithmic_bytes=kern[dom]["bytes"], peak_source=peak_src,
                kernels={k: dict(ms=v["ms"], gbs=v["bytes"] / (v["ms"] * 1e-3) / 1e9, frac=v["bytes"] / (v["ms"] * 1e-3) / 1e9 / peak)
                         for k, v in kern.items()},
                triangulate=dict(ms=tri["ms"], gbs=tri["bytes"] / (tri["ms"] * 1e-3) / 1e9, frac=tri["bytes"] / (tri["ms"] * 1e-3) / 1e9 / peak,
                                 points_per_s=2.0 * n_tri / (tri["ms"] * 1e-3)))

    # ---- the reference-facing call: arapOptimization(Map*, ...) through the C++ shim (Map gather, Delaunay mesh on the GPU,
    # LM, Map write-back), the flow of Execution/simulation.cc on a sheet scene of the headline's size
    shim = None
    if world == 1 and not args.no_shim_e2e:
        try:
            sys.path.insert(0, os.path.join(ROOT, "profiles"))
            import shim_e2e
            with _stdout_to_stderr():
                shim = shim_e2e.run(min(n, args.shim_points), reps=2)      # the second run: driver and page cache warm
        except Exception as ex:
            shim = dict(error=str(ex)[-400:])

    # ---- classic bundle adjustment: two views, --ba-points map points, the reference's bundleAdjustment settings
    cba = None
    if world == 1 and not args.no_classic_ba:
        try:
            cba = run_classic_ba(pkg, args, peak)
        except Exception as ex:
            cba = dict(error=str(ex)[-400:])

    # ---- config 5 (batch of independent 10k pairs, sharded by problem index) as a sub-record of the same line at every N
    c5 = None
    if args.config5_problems > 0:
        try:
            c5 = run_config5(pkg, args, rank, world, local, dist, args.config5_problems, 10_000, 3, 2)
        except Exception as ex:                       # a sub-record never takes the headline down (all ranks fail alike)
            c5 = dict(error=str(ex))

    # ---- ONE pair over all N GPUs (point-sharded): rank 0's frame pair refined by all ranks together
    ps = None
    if world > 1 and not args.no_point_sharded:
        try:
            ps = run_point_sharded(pkg, args, rank, world, local, dist, max(2, args.steps // 2), lm_iters, w)
        except Exception as ex:
            ps = dict(error=str(ex))

    # ---- max over ranks, aggregate
    if dist is not None:
        import importlib
        sh = importlib.import_module(pkg.__name__ + ".sharding")
        (dev_ms, wall_ms, e2e_ms), (its, pcg_its, launches) = sh.fold(dist, f"cuda:{local}", [dev_ms, wall_ms, e2e_ms],
                                                                        [its, pcg_its, launches])
        its, pcg_its, launches = int(its), int(pcg_its), int(launches)
    if rank == 0:
        step_ms = max(dev_ms, 0.0) / args.steps
        value = its / (dev_ms * 1e-3)
        e2e_val = (lm_iters * args.steps * world) / (e2e_ms * 1e-3)
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            try:
                from oracle import bench_cpu
                with _stdout_to_stderr():
                    counts = full_counts or dict(lm_iters=its // args.steps, trials=trials // args.steps, pcg_iters=pcg_its // args.steps)
                    cpu = bench_cpu.components(sc, prob, sc["weights"], counts, args.pcg_rtol)
            except Exception as ex:      # the baseline is a report, never a gate
                cpu = dict(value=None, unit="LM it/s", cores=1, kind="port", sample=f"failed: {ex}")
        line = dict(metric="non-rigid LM iterations/s at 1M correspondences", value=value, unit="LM it/s", n_gpus=world,
                    steps=args.steps, warmup=args.warmup, ms_per_step=step_ms, higher_is_better=True, scaling="weak",
                    vs_baseline=None, dtype="f64", data="synthetic",
                    config=dict(workload=sc["name"], correspondences=n, directed_edges=E, k=args.k, lm_iters_per_step=lm_iters,
                                pcg_rtol=args.pcg_rtol, pcg_iters_per_lm_iter=pcg_its / max(1, its),
                                lm_trials_per_step=trials / args.steps, early_reject=early, l2="working set ~1.1 GB/GPU, larger than the 126 MB L2",
                                frame_pairs=world, parallelism=f"{world} independent frame pairs, one per GPU"),
                    e2e=dict(value=e2e_val, unit="LM it/s", h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h),
                             ms_per_step=e2e_ms / args.steps),
                    gpu_launches=int(launches), clocks=clocks, roofline=roof, cpu_baseline=cpu,
                    triangulated_points_per_s=roof["triangulate"]["points_per_s"], knn_graph_build_ms=prob["graph_build_ms"],
                    cg_iteration_ms=cg_ms, wall_ms_per_step=wall_ms / args.steps, fp32_mode=f32rec, reference_api_e2e=shim, classic_ba=cba, config5=c5)
        if ps is not None:
            if "error" not in ps:
                ps["speedup_vs_one_gpu"] = ps["lm_it_per_s"] / (value / world)      # the same pair, same LM trace, on ONE of these GPUs
                ps["strong_scaling_efficiency"] = ps["speedup_vs_one_gpu"] / world
            line["point_sharded"] = ps
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    ctx.close()
    return 0


def run_point_sharded(pkg, args, rank, world, local, dist, steps, lm_iters, w):
    """Strong scaling of ONE frame pair (the workload's pair of seed 0, i.e. rank 0's pair of the headline) over all N GPUs:
    dsc_shard_* -- every rank uploads the same pair and refines its own range of tiles; halo rows and PCG scalars cross
    NVLink inside the kernels.  Timed like the headline: device-resident, barrier + synchronize, max over ranks."""
    import importlib
    import torch
    sh = importlib.import_module(pkg.__name__ + ".sharding")
    sc = make_scene(pkg, args, seed=0)
    c0 = pkg.Context(local)
    prob = prepare(pkg, c0, sc, args)
    c0.close()
    sp = sh.ShardedPair(pkg, local, len(prob["X1"]), dist, world, rank)
    ctx = sp.ctx
    ctx.set_pcg(rtol=args.pcg_rtol, max_iters=args.pcg_max_iters, check_every=64)
    upload(ctx, prob)
    ref = None
    early_ok = None
    for k in range(2):                                   # warm-up 0: every solve to the tolerance; 1: with the early rejection
        ctx.reset_state()
        recs, st = ctx.optimize(w, lm_iters)
        tr = [(r.chi2_before, r.chi2_after, r.trials, r.accepted) for r in recs]
        if k == 0:
            ref = tr
            if len(args.early_rtol) > 0:
                ctx.set_early_reject(args.early_rtol, args.early_margin)
        else:
            early_ok = (tr == ref)
            if not early_ok:
                ctx.set_early_reject((), ())
    dist.barrier(); torch.cuda.synchronize(); ctx.synchronize()
    ms, its, pcg = 0.0, 0, 0
    for _ in range(steps):
        ctx.reset_state()
        recs, st = ctx.optimize(w, lm_iters)
        ms += st.device_ms; its += st.iterations; pcg += st.total_pcg_iters
    dist.barrier(); torch.cuda.synchronize()
    info = ctx.shard_info()
    (ms_max,), (halo, rows) = sh.fold(dist, f"cuda:{local}", [ms], [info["halo_rows"], info["row_end"] - info["row_begin"]])
    finals = [None] * world
    dist.all_gather_object(finals, (st.final_chi2, [r.trials for r in recs]))
    sp.close(dist)
    return dict(workload=sc["name"] + ", ONE pair over all GPUs", n_gpus=world, steps=steps, lm_iters_per_step=lm_iters, scaling="strong",
                lm_it_per_s=its / (ms_max * 1e-3), ms_per_step=ms_max / steps, pcg_iters_per_lm_iter=pcg / max(1, its),
                halo_rows_all_ranks=int(halo), rows_all_ranks=int(rows),
                nvlink_bytes_per_pcg_iteration=int(halo) * 48 + world * world * 80,
                nvlink_bytes_per_trial=int(halo) * 64 + world * world * (8 + 24),
                exchange="halo rows pushed by the producing kernel into peer-mapped memory; scalars + 8 global rows by its last block; no collective call",
                all_ranks_same_bits=all(f == finals[0] for f in finals), early_reject_trace_identical=early_ok,
                final_chi2=st.final_chi2)


def batch_bytes(prob_sizes, recs, stats):
    """algorithmic bytes one lm_batch_kernel launch moves (per-phase formulas of the 1M path summed with
    every pair's own counts): linearise 488 N + 12 S + 72 E per LM iteration; per trial 480 N (preconditioner) + 224 N (trial
    state) + 136 N + 12 S (cost); per PCG iteration 936 N + 76 S (operator + update)."""
    total = 0.0
    for (n, e), st in zip(prob_sizes, stats):
        s_ = 1.04 * e
        total += st.iterations * (488.0 * n + 12.0 * s_ + 72.0 * e)
        total += st.total_trials * (480.0 * n + 224.0 * n + 136.0 * n + 12.0 * s_)
        total += st.total_pcg_iters * (936.0 * n + 76.0 * s_)
    return total


def run_config5(pkg, args, rank, world, local, dist, per_gpu, n_points, steps, warmup, total=None):
    """Config 5: independent ~10k-correspondence frame pairs (config-2 generator, seed = problem index), sharded over the
    ranks by problem index, every rank's share refined by ONE launch of lm_batch_kernel (thread-block clusters take pairs
    from a device-side queue and run the whole LM loop of a pair).  No data-path collective; one gather of the per-problem
    results.  total = None: weak scaling (per_gpu pairs on every GPU); total = N: strong scaling (N pairs over all GPUs)."""
    import importlib
    sh = importlib.import_module(pkg.__name__ + ".sharding")
    wl = pkg_workloads(pkg)
    n_problems = total if total else per_gpu * world
    sb = sh.ShardedBatch(pkg, local, n_problems, world, rank)
    ctx = pkg.Context(local)
    a2 = argparse.Namespace(**vars(args))
    a2.n, a2.workload = n_points, "sheet"
    t_build = time.perf_counter()

    def problem_of(pidx):
        sc = wl.make_scene("sheet", n_points, pidx)
        pr = prepare(pkg, ctx, sc, a2)
        pr["_sc"] = sc
        return pr
    sb.build(problem_of)
    build_ms = (time.perf_counter() - t_build) * 1e3
    ctx.close()
    sc0 = sb.problems[0]["_sc"] if sb.problems else wl.make_scene("sheet", n_points, 0)
    lm_iters = args.lm_iters or sc0["lm_iters"]
    w = pkg.make_weights(**sc0["weights"])
    b = sb.batch
    b.set_pcg(rtol=args.pcg_rtol, max_iters=args.pcg_max_iters)
    h2d = sb.upload()
    sizes = [(len(p["X1"]), len(p["col"])) for p in sb.problems]
    early = dict(rtol_loose=args.early_rtol, rho_margin=args.early_margin, used=False, trace_identical=None)
    ref = None
    for k in range(max(2, warmup)):                      # warm-up 0: every solve to the tolerance; then with early rejection
        b.reset_state()
        recs, stats, ms = b.optimize(w, lm_iters)
        tr = [[(r.chi2_before, r.chi2_after, r.trials, r.accepted) for r in rr] for rr in recs]
        if k == 0:
            ref = tr
            early["lm_it_per_s_full_solves_this_rank"] = sum(s_.iterations for s_ in stats) / (ms * 1e-3)
            if len(args.early_rtol) > 0:
                b.set_early_reject(args.early_rtol, args.early_margin)
                early["used"] = True
        elif early["used"] and early["trace_identical"] is None:
            early["trace_identical"] = (tr == ref)
            if not early["trace_identical"]:
                b.set_early_reject((), ())
                early["used"] = False

    def barrier():
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()
    barrier()
    dev_ms, its, pcg, trials, rejects, nbytes = 0.0, 0, 0, 0, 0, 0.0
    for _ in range(steps):
        b.reset_state()
        recs, stats, ms = b.optimize(w, lm_iters)
        dev_ms += ms
        its += sum(s_.iterations for s_ in stats)
        pcg += sum(s_.total_pcg_iters for s_ in stats)
        trials += sum(s_.total_trials for s_ in stats)
        rejects += sum(s_.early_rejects for s_ in stats)
        nbytes += batch_bytes(sizes, recs, stats)
    barrier()
    # end to end: host arrays in (upload + device-side graph set-up), one launch, host arrays out
    e2e_ms = 0.0
    for s_ in range(1 + steps):
        barrier()
        t1 = time.perf_counter()
        sb.upload()
        recs, stats, ms = b.optimize(w, lm_iters)
        outs = b.download()
        barrier()
        if s_ > 0:
            e2e_ms += (time.perf_counter() - t1) * 1e3
    d2h = sum(o["X1"].nbytes + o["X2"].nbytes for o in outs)
    rows = sb.result_rows(stats)
    dev = f"cuda:{local}" if dist is not None else "cpu"
    table = sh.gather_by_problem(dist, dev, n_problems, world, rank, rows)          # the one collective: results by problem
    (dev_ms_max, e2e_ms_max), (its_all, pcg_all, trials_all, bytes_all) = sh.fold(dist, dev, [dev_ms, e2e_ms], [its, pcg, trials, nbytes])
    info = b.size()
    sb.close()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    ach = nbytes / (dev_ms * 1e-3) / 1e9 if dev_ms > 0 else 0.0
    rate = its_all / (dev_ms_max * 1e-3)
    return dict(workload="config5: independent config-2 frame pairs (seed = problem index), sharded by problem index, one launch per GPU",
                problems=n_problems, problems_per_gpu=len(sb.mine), correspondences_per_problem=n_points,
                directed_edges_per_problem=int(np.mean([e for _, e in sizes])) if sizes else 0, lm_iters_per_problem=lm_iters,
                scaling="strong" if total else "weak", steps=steps,
                lm_it_per_s=rate, lm_it_per_s_per_gpu=rate / world, problems_per_s=rate / lm_iters,
                e2e_lm_it_per_s=(lm_iters * n_problems * steps) / (e2e_ms_max * 1e-3), ms_per_step=dev_ms_max / steps,
                e2e_ms_per_step=e2e_ms_max / steps, h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h),
                pcg_iters_per_lm_iter=pcg_all / max(1.0, its_all), lm_trials_per_problem=trials_all / max(1, n_problems * steps),
                early_reject=early, cluster_ctas=info["cluster_ctas"], clusters_per_gpu=info["clusters"], gpu_launches_per_step=1,
                roofline=dict(bound="hbm", kernel="lm_batch_kernel", achieved=ach, peak=peak, unit="GB/s", frac=ach / peak,
                              note="algorithmic bytes of all phases of all pairs of this rank / launch time; a pair's working set "
                                   "(~18 MB) is L2-resident while its cluster refines it, so the fraction of the HBM peak can exceed 1"),
                gathered=dict(problems=int(table.shape[0]), final_chi2_sum=float(table[:, 0].sum()), lm_iterations=int(table[:, 1].sum())),
                host_build_ms_this_rank=build_ms)


def run_classic_ba(pkg, args, peak):
    """bundleAdjustment's optimisation (g2oBundleAdjustment.cc:38-141: Huber sqrt(5.99), 20 LM iterations) on two views of
    --ba-points map points: LM iterations/s resident and from host arrays, bytes of the run against the measured HBM peak, and the
    numpy oracle (full-system sparse solve, one core) on a bounded sample of the same scene."""
    import importlib
    wl = importlib.import_module(pkg.__name__ + ".workloads")
    n = int(args.ba_points)
    sc = wl.ba_scene(n, 2, seed=0)
    huber = float(np.float32(np.sqrt(5.99)))
    iters = 20
    with pkg.BundleAdjuster(0) as b:
        def upload():
            b.upload(sc["poses7"], sc["pose_fixed"], sc["cams"], sc["X"], sc["obs_pose"], sc["obs_point"], sc["obs_uv"], sc["obs_isg"])
        upload()
        b.optimize(iters, huber)                                   # warm-up (kernel loading)
        res, e2e = [], []
        for _ in range(3):
            t0 = time.perf_counter()
            upload()
            recs, st = b.optimize(iters, huber)                    # (device_ms: CUDA events around the LM run, inputs resident)
            p7, X = b.download()
            e2e.append((time.perf_counter() - t0) * 1e3)
            res.append((st.device_ms, st.iterations, st.total_trials, st.kernel_launches, recs[0].chi2_before, st.final_chi2))
        ms, its, trials, launches, chi0, chi1 = res[-1]
        ms = float(np.median([r[0] for r in res]))
    O = len(sc["obs_pose"])
    # algorithmic bytes: linearise once per iteration (X 32 + Hll 48 + bl 24 per point, 17 in + 360 out per observation), per trial the
    # Schur pass (444 per diagonal entry: one per point seen from the free view), the back-substitution (136 per point + 148 per
    # observation) and the cost (32 per point + 17 per observation)
    by = its * (104.0 * n + 377.0 * O) + trials * (444.0 * n + 136.0 * n + 148.0 * O + 32.0 * n + 17.0 * O)
    rec = dict(workload=f"bundleAdjustment: 2 key frames x {n} map points ({O} observations), key frame 0 fixed, Huber sqrt(5.99), {iters} LM iterations, "
                        "points marginalised (Schur complement on the device)",
               lm_it_per_s=its / (ms * 1e-3), ms_per_run=ms, lm_iterations=int(its), trials=int(trials), gpu_launches_per_run=int(launches),
               e2e=dict(lm_it_per_s=its / (float(np.median(e2e)) * 1e-3), ms_per_run=float(np.median(e2e)),
                        h2d_bytes=int(sc["X"].nbytes + sc["obs_uv"].nbytes + sc["obs_pose"].nbytes + sc["obs_point"].nbytes + sc["obs_isg"].nbytes),
                        d2h_bytes=int(sc["X"].nbytes), note="upload (observation sort and Schur entry list on the host) + optimize + download"),
               chi2=[float(chi0), float(chi1)],
               roofline=dict(bound="hbm", achieved=by / (ms * 1e-3) / 1e9, peak=peak, unit="GB/s", frac=by / (ms * 1e-3) / 1e9 / peak,
                             note="algorithmic bytes of every kernel of the run / device time of the run (host round trips of the LM loop included)"))
    if not args.no_cpu_baseline:
        try:
            sys.path.insert(0, ROOT)
            from oracle import ba as oba
            from oracle.se3 import SE3
            m = min(n, 20_000)
            small = wl.ba_scene(m, 2, seed=0)
            prob = oba.BaProblem(poses=[SE3.from7(q) for q in small["poses7"]], pose_fixed=small["pose_fixed"], cams=small["cams"], X=small["X"],
                                 obs_pose=small["obs_pose"], obs_point=small["obs_point"], obs_uv=small["obs_uv"], obs_isg=small["obs_isg"].astype(np.float64))
            t0 = time.perf_counter()
            _, _, tr = oba.optimize(prob, 3, robust=True)
            dt = time.perf_counter() - t0
            done = len(tr["trials"])
            rec["cpu_baseline"] = dict(value=done / dt * (m / n), unit="LM it/s", cores=1, kind="port",
                                       sample=f"oracle/ba.py (numpy + scipy sparse direct solve of the full system) on {m} of the {n} points, {done} LM "
                                              f"iterations in {dt:.1f} s, scaled linearly to {n} points")
        except Exception as ex:
            rec["cpu_baseline"] = dict(error=str(ex)[-200:])
    return rec


def main_batch(args, rank, world, local):
    """`--workload batch`: config 5 as the headline of the line (value = LM iterations/s over all pairs of all GPUs)."""
    os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 8) // max(1, world)))
    import __graft_entry__ as g
    pkg = g.package()
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        with _stdout_to_stderr():                      # NCCL prints its version banner on stdout
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
    n = args.n if args.n != 1_000_000 else 10_000
    sampler = ClockSampler(local) if rank == 0 else None
    c5 = run_config5(pkg, args, rank, world, local, dist, args.problems or 128, n, args.steps, args.warmup, total=args.total_problems or None)
    clocks = sampler.stop() if sampler else None
    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            try:
                from oracle import bench_cpu
                with _stdout_to_stderr():
                    r = bench_cpu.run_reference("sheet", n, args.k, args.lm_iters, pcg_rtol=args.pcg_rtol, budget_s=120.0)
                cpu = r["cpu_baseline"]
                cpu["sample"] = "ONE pair of the batch (the CPU refines the pairs one after the other, all threads on each): " + cpu["sample"]
            except Exception as ex:
                cpu = dict(value=None, unit="LM it/s", cores=1, kind="port", sample=f"failed: {ex}")
        line = dict(metric="non-rigid LM iterations/s, batch of independent 10k-correspondence frame pairs", value=c5["lm_it_per_s"],
                    unit="LM it/s", n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=c5["ms_per_step"],
                    higher_is_better=True, scaling=c5["scaling"], vs_baseline=None, dtype="f64", data="synthetic",
                    config=dict(workload=c5["workload"], problems=c5["problems"], problems_per_gpu=c5["problems_per_gpu"],
                                correspondences_per_problem=c5["correspondences_per_problem"], k=args.k,
                                lm_iters_per_problem=c5["lm_iters_per_problem"], pcg_rtol=args.pcg_rtol, early_reject=c5["early_reject"],
                                cluster_ctas=c5["cluster_ctas"], clusters_per_gpu=c5["clusters_per_gpu"],
                                l2="a pair's working set (~18 MB) is L2-resident by design; the batch of a GPU (~2.3 GB at 128 pairs) is not",
                                parallelism=f"{world} GPUs x {c5['problems_per_gpu']} pairs, no data-path collective"),
                    e2e=dict(value=c5["e2e_lm_it_per_s"], unit="LM it/s", h2d_bytes_per_step=c5["h2d_bytes_per_step"],
                             d2h_bytes_per_step=c5["d2h_bytes_per_step"], ms_per_step=c5["e2e_ms_per_step"]),
                    gpu_launches=int(args.steps * world), clocks=clocks, roofline=c5["roofline"], cpu_baseline=cpu, config5=c5)
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
