"""ctypes binding of the C ABI in include/dsc.h (libdsc_b200.so).

This is the thin Python face used by tests/ and bench.py; the product is the shared library.
The binding never falls back to a CPU implementation: if the library is missing or no CUDA
device is present, it raises.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libdsc_b200.so")

CAM_KB8, CAM_PINHOLE = 0, 1
TRI_CLASSIC, TRI_NRSLAM, TRI_ORBSLAM, TRI_DEPTH = 0, 1, 2, 3
LOC_INRAYS, LOC_TWOPOINTS, LOC_FARPOINTS = 0, 1, 2
GATE_NONE, GATE_SIM, GATE_REAL = 0, 1, 2
METHODS = {"Classic": 0, "NRSLAM": 1, "ORBSLAM": 2, "DepthMeasurement": 3}
LOCATIONS = {"InRays": 0, "TwoPoints": 1, "FarPoints": 2}

EXPORTS = [
    "dsc_create", "dsc_destroy", "dsc_last_error", "dsc_status_string", "dsc_version", "dsc_synchronize",
    "dsc_timer_start", "dsc_timer_stop", "dsc_launch_count", "dsc_pin_host", "dsc_unpin_host",
    "dsc_triangulate", "dsc_triangulate_rays", "dsc_tri_upload", "dsc_tri_run", "dsc_tri_download", "dsc_depth_scale_init",
    "dsc_problem_upload", "dsc_set_graph", "dsc_compute_rotations", "dsc_get_rotations", "dsc_set_rotations",
    "dsc_reset_state", "dsc_set_pcg", "dsc_set_solver", "dsc_set_precision", "dsc_set_early_reject", "dsc_cost", "dsc_optimize", "dsc_download", "dsc_pixel_sigma",
    "dsc_batch_create", "dsc_batch_destroy", "dsc_batch_last_error", "dsc_batch_upload", "dsc_batch_set_pcg", "dsc_batch_set_early_reject",
    "dsc_batch_reset_state", "dsc_batch_set_active", "dsc_batch_pixel_sigma", "dsc_batch_optimize", "dsc_batch_download", "dsc_batch_size",
    "dsc_shard_init", "dsc_shard_attach", "dsc_shard_partition", "dsc_shard_info",
    "dsc_debug_linearize", "dsc_debug_matvec", "dsc_profile_kernels", "dsc_profile_triangulate", "dsc_problem_size", "dsc_knn_build", "dsc_knn_download",
    "dsc_delaunay_build", "dsc_delaunay_download", "dsc_set_graph_delaunay",
    "dsc_ba_create", "dsc_ba_destroy", "dsc_ba_last_error", "dsc_ba_upload", "dsc_ba_set_poses", "dsc_ba_set_levels", "dsc_ba_optimize",
    "dsc_ba_edge_chi2", "dsc_ba_download", "dsc_ba_launch_count",
]
KERNEL_NAMES = ["cg_spmv", "cg_update", "linearize", "cost", "precond", "apply_update", "rotations"]


class Camera(C.Structure):
    _fields_ = [("model", C.c_int), ("params", C.c_float * 8)]


class Pair(C.Structure):
    _fields_ = [("cam1", Camera), ("cam2", Camera), ("T1w", C.c_float * 12), ("T2w", C.c_float * 12)]


class TriParams(C.Structure):
    _fields_ = [("method", C.c_int), ("location", C.c_int), ("gate", C.c_int), ("min_cos", C.c_float),
                ("depth_limit", C.c_float), ("check_reproj", C.c_int)]


class Weights(C.Structure):
    _fields_ = [("rep", C.c_double), ("glob", C.c_double), ("arap", C.c_double), ("alpha", C.c_double),
                ("beta", C.c_double), ("depth_sigma", C.c_float)]


class PcgParams(C.Structure):
    _fields_ = [("rtol", C.c_double), ("max_iters", C.c_int), ("check_every", C.c_int)]


class IterRecord(C.Structure):
    _fields_ = [("chi2_before", C.c_double), ("chi2_after", C.c_double), ("lam", C.c_double), ("trials", C.c_int),
                ("accepted", C.c_int), ("pcg_iters", C.c_int)]


class OptStats(C.Structure):
    _fields_ = [("iterations", C.c_int), ("total_trials", C.c_int), ("total_pcg_iters", C.c_int), ("terminated", C.c_int),
                ("final_chi2", C.c_double), ("device_ms", C.c_double), ("linearize_ms", C.c_double), ("pcg_ms", C.c_double),
                ("trial_ms", C.c_double), ("kernel_launches", C.c_int), ("early_rejects", C.c_int), ("pcg_unconverged", C.c_int)]


class BatchPair(C.Structure):
    _fields_ = [("pair", Pair), ("scale1", C.c_double), ("scale2", C.c_double), ("Tg7", C.c_double * 7), ("area", C.c_double),
                ("n_triangles", C.c_longlong)]


class DscError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"dsc status {status}: {msg}")
        self.status = status


_lib = None


def load_library(path=None):
    """dlopen libdsc_b200.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or os.environ.get("DSC_B200_LIB") or LIB_PATH      # the override serves kernel experiments
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                f"(the hot path has no CPU fallback)")
    lib = C.CDLL(path)
    lib.dsc_last_error.restype = C.c_char_p
    lib.dsc_status_string.restype = C.c_char_p
    lib.dsc_destroy.restype = None
    lib.dsc_batch_last_error.restype = C.c_char_p
    lib.dsc_batch_destroy.restype = None
    lib.dsc_ba_last_error.restype = C.c_char_p
    lib.dsc_ba_destroy.restype = None
    _lib = lib
    return lib


def _fp(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f32(a, shape=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float32)
    if shape is not None:
        a = a.reshape(shape)
    return a


def _f64(a, shape=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


def pin_host(*arrays):
    """page-lock numpy arrays in place (dsc_pin_host): uploads from them are then plain DMA, no host-side staging"""
    lib = load_library()
    for a in arrays:
        if a is not None and a.nbytes > 0:
            if not a.flags["C_CONTIGUOUS"]:
                raise ValueError("pin_host needs C-contiguous arrays")
            st = lib.dsc_pin_host(C.c_void_p(a.ctypes.data), C.c_size_t(a.nbytes))
            if st != 0:
                raise DscError(st, "cudaHostRegister failed")


def unpin_host(*arrays):
    lib = load_library()
    for a in arrays:
        if a is not None and a.nbytes > 0:
            lib.dsc_unpin_host(C.c_void_p(a.ctypes.data))


def make_pair(cam1, cam2, T1w, T2w):
    """cam = (model, params[8]); Tcw = 3x4 float32 [R|t] (or an object with .as34())."""
    p = Pair()
    for dst, cam in ((p.cam1, cam1), (p.cam2, cam2)):
        dst.model = int(cam[0])
        prm = np.asarray(cam[1], np.float32).reshape(8)
        for k in range(8):
            dst.params[k] = float(prm[k])
    for dst, T in ((p.T1w, T1w), (p.T2w, T2w)):
        M = T.as34() if hasattr(T, "as34") else np.asarray(T, np.float32)
        M = np.asarray(M, np.float32).reshape(12)
        for k in range(12):
            dst[k] = float(M[k])
    return p


def make_weights(rep=1.0, arap=1.0, depth_sigma=1.0, glob=0.0, alpha=1.0, beta=1.0):
    return Weights(float(rep), float(glob), float(arap), float(alpha), float(beta), float(depth_sigma))


class Context:
    """One dsc_ctx: one CUDA stream, driven by one host thread."""

    def __init__(self, device=0):
        self.lib = load_library()
        self.h = C.c_void_p()
        st = self.lib.dsc_create(int(device), C.byref(self.h))
        if st != 0:
            raise DscError(st, self.lib.dsc_status_string(st).decode())
        self.n = 0
        self.tn = 0

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.lib.dsc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, st):
        if st != 0:
            raise DscError(st, self.lib.dsc_last_error(self.h).decode())

    # ---- timing / bookkeeping
    def synchronize(self):
        self._ck(self.lib.dsc_synchronize(self.h))

    def timer_start(self):
        self._ck(self.lib.dsc_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_double()
        self._ck(self.lib.dsc_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def launch_count(self):
        c = C.c_longlong()
        self._ck(self.lib.dsc_launch_count(self.h, C.byref(c)))
        return c.value

    # ---- K1
    @staticmethod
    def tri_params(method="NRSLAM", location="FarPoints", gate=GATE_SIM, min_cos=0.9998, depth_limit=np.inf,
                   check_reproj=False):
        m = METHODS.get(method, 1) if isinstance(method, str) else int(method)
        loc = LOCATIONS.get(location, 0) if isinstance(location, str) else int(location)
        return TriParams(m, loc, int(gate), float(min_cos), float(min(depth_limit, 3.0e38)), int(bool(check_reproj)))

    def triangulate(self, pair, prm, uv1, uv2, d1=None, d2=None, out=None):
        """out: optional dict of preallocated (e.g. page-locked) result arrays X1, X2 [n][3] f32, valid [n] u8, cosp [n] f32"""
        uv1, uv2 = _f32(uv1, (-1, 2)), _f32(uv2, (-1, 2))
        n = uv1.shape[0]
        d1, d2 = _f32(d1), _f32(d2)
        out = out or {}
        X1 = out.get("X1") if out.get("X1") is not None else np.empty((n, 3), np.float32)
        X2 = out.get("X2") if out.get("X2") is not None else np.empty((n, 3), np.float32)
        valid = out.get("valid") if out.get("valid") is not None else np.empty(n, np.uint8)
        cosp = out.get("cosp") if out.get("cosp") is not None else np.empty(n, np.float32)
        nv = C.c_int()
        self._ck(self.lib.dsc_triangulate(self.h, C.byref(pair), C.byref(prm), n, _fp(uv1), _fp(uv2), _fp(d1), _fp(d2),
                                          _fp(X1), _fp(X2), _fp(valid), _fp(cosp), C.byref(nv)))
        self.tn = n
        return X1, X2, valid.astype(bool), cosp, nv.value

    def triangulate_rays(self, T1w, T2w, method, location, xn1, xn2):
        """useTriangulationMethod with the reference's argument list: rays in, world points out (no gates)"""
        xn1, xn2 = _f32(xn1, (-1, 3)), _f32(xn2, (-1, 3))
        n = xn1.shape[0]
        T1 = _f32(T1w.as34() if hasattr(T1w, "as34") else T1w, (12,))
        T2 = _f32(T2w.as34() if hasattr(T2w, "as34") else T2w, (12,))
        m = METHODS.get(method, 1) if isinstance(method, str) else int(method)
        loc = LOCATIONS.get(location, 0) if isinstance(location, str) else int(location)
        X1 = np.empty((n, 3), np.float32)
        X2 = np.empty((n, 3), np.float32)
        self._ck(self.lib.dsc_triangulate_rays(self.h, _fp(T1), _fp(T2), m, loc, n, _fp(xn1), _fp(xn2), _fp(X1), _fp(X2)))
        return X1, X2

    def tri_upload(self, pair, uv1, uv2, d1=None, d2=None):
        uv1, uv2 = _f32(uv1, (-1, 2)), _f32(uv2, (-1, 2))
        self.tn = uv1.shape[0]
        self._ck(self.lib.dsc_tri_upload(self.h, C.byref(pair), self.tn, _fp(uv1), _fp(uv2), _fp(_f32(d1)), _fp(_f32(d2))))

    def tri_run(self, prm):
        self._ck(self.lib.dsc_tri_run(self.h, C.byref(prm)))

    def tri_download(self):
        n = self.tn
        X1 = np.empty((n, 3), np.float32)
        X2 = np.empty((n, 3), np.float32)
        valid = np.empty(n, np.uint8)
        cosp = np.empty(n, np.float32)
        nv = C.c_int()
        self._ck(self.lib.dsc_tri_download(self.h, _fp(X1), _fp(X2), _fp(valid), _fp(cosp), C.byref(nv)))
        return X1, X2, valid.astype(bool), cosp, nv.value

    def depth_scale_init(self, which):
        s = C.c_double()
        self._ck(self.lib.dsc_depth_scale_init(self.h, int(which), C.byref(s)))
        return s.value

    # ---- refinement problem
    def problem_upload(self, pair, X1, X2, uv1, uv2, d1, d2, inv_sigma2_1=None, inv_sigma2_2=None,
                       scale1=1.0, scale2=1.0, Tg7=None):
        X1, X2 = _f32(X1, (-1, 3)), _f32(X2, (-1, 3))
        n = X1.shape[0]
        uv1, uv2 = _f32(uv1, (-1, 2)), _f32(uv2, (-1, 2))
        d1, d2 = _f64(d1), _f64(d2)
        s1, s2 = _f32(inv_sigma2_1), _f32(inv_sigma2_2)
        tg = _f64(Tg7)
        self._ck(self.lib.dsc_problem_upload(self.h, C.byref(pair), n, _fp(X1), _fp(X2), _fp(uv1), _fp(uv2), _fp(d1), _fp(d2),
                                             _fp(s1), _fp(s2), C.c_double(scale1), C.c_double(scale2), _fp(tg)))
        self.n = n

    def set_graph(self, rowptr, col, w, area, n_triangles, reorder=1):
        rowptr = np.ascontiguousarray(rowptr, np.int32)
        col = np.ascontiguousarray(col, np.int32)
        w = _f64(w)
        self._ck(self.lib.dsc_set_graph(self.h, len(rowptr) - 1, _fp(rowptr), _fp(col), _fp(w), C.c_double(area),
                                        C.c_longlong(int(n_triangles)), int(reorder)))

    def compute_rotations(self):
        self._ck(self.lib.dsc_compute_rotations(self.h))

    def get_rotations(self):
        q = np.empty((self.n, 4), np.float64)
        self._ck(self.lib.dsc_get_rotations(self.h, _fp(q)))
        return q

    def set_rotations(self, quat):
        q = _f64(quat, (-1, 4))
        self._ck(self.lib.dsc_set_rotations(self.h, _fp(q)))

    def reset_state(self):
        self._ck(self.lib.dsc_reset_state(self.h))

    def set_solver(self, solver=0):
        """0 auto (dense Cholesky for small problems, PCG above), 1 PCG, 2 dense."""
        self._ck(self.lib.dsc_set_solver(self.h, int(solver)))

    def set_pcg(self, rtol=1e-10, max_iters=4000, check_every=32):
        p = PcgParams(float(rtol), int(max_iters), int(check_every))
        self._ck(self.lib.dsc_set_pcg(self.h, C.byref(p)))

    # ---- point-sharded pair (dsc.h: dsc_shard_*): this context is one rank of ONE frame pair
    def shard_init(self, rank, world, max_points):
        """-> the 64-byte inter-process handle of this rank's exchange arena"""
        h = np.zeros(64, np.uint8)
        self._ck(self.lib.dsc_shard_init(self.h, int(rank), int(world), int(max_points), _fp(h)))
        return h

    def shard_attach(self, handles):
        """handles: [world][64] uint8, rank order (every rank's shard_init result)"""
        hs = np.ascontiguousarray(handles, np.uint8)
        self._ck(self.lib.dsc_shard_attach(self.h, _fp(hs)))

    def shard_info(self):
        r, w, b, e, hl = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_longlong()
        self._ck(self.lib.dsc_shard_info(self.h, C.byref(r), C.byref(w), C.byref(b), C.byref(e), C.byref(hl)))
        return dict(rank=r.value, world=w.value, row_begin=b.value, row_end=e.value, halo_rows=hl.value)

    def set_precision(self, precision="f64"):
        """"f64" (default) or "f32": storage of the data the PCG streams (dsc.h, dsc_set_precision)"""
        self._ck(self.lib.dsc_set_precision(self.h, {"f64": 0, "f32": 1}.get(precision, precision)))

    def set_early_reject(self, rtol_loose=(1e-3, 1e-4), rho_margin=(1.0, 0.5)):
        """levels of (loose tolerance, rho margin); empty sequences switch the shortcut off"""
        r = np.ascontiguousarray(np.atleast_1d(rtol_loose), np.float64)
        m = np.ascontiguousarray(np.atleast_1d(rho_margin), np.float64)
        if len(r) != len(m):
            raise ValueError("one margin per tolerance")
        self._ck(self.lib.dsc_set_early_reject(self.h, len(r), _fp(r), _fp(m)))

    def cost(self, w):
        chi = C.c_double()
        parts = (C.c_double * 3)()
        self._ck(self.lib.dsc_cost(self.h, C.byref(w), C.byref(chi), parts))
        return chi.value, tuple(parts)

    def optimize(self, w, n_iters):
        recs = (IterRecord * max(1, n_iters))()
        st = OptStats()
        self._ck(self.lib.dsc_optimize(self.h, C.byref(w), int(n_iters), recs, C.byref(st)))
        return [recs[i] for i in range(st.iterations)], st

    def download(self, doubles=True, out=None):
        """out: optional dict of preallocated (e.g. page-locked) X1, X2 [n][3] f32 result arrays"""
        n = self.n
        out = out or {}
        X1 = out.get("X1") if out.get("X1") is not None else np.empty((n, 3), np.float32)
        X2 = out.get("X2") if out.get("X2") is not None else np.empty((n, 3), np.float32)
        X1d = np.empty((n, 3), np.float64) if doubles else None
        X2d = np.empty((n, 3), np.float64) if doubles else None
        scales = np.empty(2, np.float64)
        tg = np.empty(7, np.float64)
        upd = C.c_double()
        self._ck(self.lib.dsc_download(self.h, _fp(X1), _fp(X2), _fp(X1d), _fp(X2d), _fp(scales), _fp(tg), C.byref(upd)))
        return dict(X1=X1, X2=X2, X1d=X1d, X2d=X2d, scales=scales, Tg=tg, update=upd.value)

    def pixel_sigma(self):
        s = np.empty(2, np.float64)
        self._ck(self.lib.dsc_pixel_sigma(self.h, _fp(s)))
        return s

    def debug_linearize(self, w):
        nu = 8 + 6 * self.n
        b = np.empty(nu)
        hd = np.empty(nu)
        chi = C.c_double()
        self._ck(self.lib.dsc_debug_linearize(self.h, C.byref(w), _fp(b), _fp(hd), C.byref(chi)))
        return b, hd, chi.value

    def debug_matvec(self, w, lam, x):
        x = _f64(x)
        y = np.empty_like(x)
        self._ck(self.lib.dsc_debug_matvec(self.h, C.byref(w), C.c_double(lam), _fp(x), _fp(y)))
        return y

    def profile_kernels(self, w, warm=3, reps=20):
        ms = (C.c_double * 7)()
        by = (C.c_double * 7)()
        self._ck(self.lib.dsc_profile_kernels(self.h, C.byref(w), int(warm), int(reps), ms, by))
        return {k: dict(ms=ms[i], bytes=by[i]) for i, k in enumerate(KERNEL_NAMES)}

    def profile_triangulate(self, prm, warm=3, reps=20):
        ms = C.c_double()
        by = C.c_double()
        self._ck(self.lib.dsc_profile_triangulate(self.h, C.byref(prm), int(warm), int(reps), C.byref(ms), C.byref(by)))
        return dict(ms=ms.value, bytes=by.value)

    def problem_size(self):
        n = C.c_longlong()
        e = C.c_longlong()
        self._ck(self.lib.dsc_problem_size(self.h, C.byref(n), C.byref(e)))
        return n.value, e.value

    def knn_graph(self, X, k):
        """symmetrised k-NN graph in (x, y) built on the GPU -> (rowptr, col, unit weights)"""
        X = _f32(X, (-1, 3))
        n = X.shape[0]
        E = C.c_longlong()
        self._ck(self.lib.dsc_knn_build(self.h, n, _fp(X), int(k), C.byref(E)))
        rowptr = np.empty(n + 1, np.int32)
        col = np.empty(max(E.value, 1), np.int32)
        self._ck(self.lib.dsc_knn_download(self.h, _fp(rowptr), _fp(col)))
        col = col[:E.value]
        return rowptr, col, np.ones(E.value, np.float64)


    def delaunay_graph(self, X, min_weight=0.0):
        """the reference's neighbour graph built on the GPU (dsc_delaunay_build): 2-D Delaunay of (x, y), cot weights
        -> (rowptr, col, w, area, n_triangles, cells finished in the second pass)"""
        X = _f32(X, (-1, 3))
        n = X.shape[0]
        E, T, U = C.c_longlong(), C.c_longlong(), C.c_longlong()
        area = C.c_double()
        self._ck(self.lib.dsc_delaunay_build(self.h, n, _fp(X), C.c_double(min_weight), C.byref(E), C.byref(T), C.byref(area), C.byref(U)))
        rowptr = np.empty(n + 1, np.int32)
        col = np.empty(max(E.value, 1), np.int32)
        w = np.empty(max(E.value, 1), np.float64)
        self._ck(self.lib.dsc_delaunay_download(self.h, _fp(rowptr), _fp(col), _fp(w)))
        return rowptr, col[:E.value], w[:E.value], area.value, T.value, U.value

    def set_graph_delaunay(self, min_weight=0.0, reorder=1):
        """Delaunay graph of the uploaded problem's KF1 points, built and installed on the device -> (area, n_triangles, E)"""
        area = C.c_double()
        T, E = C.c_longlong(), C.c_longlong()
        self._ck(self.lib.dsc_set_graph_delaunay(self.h, C.c_double(min_weight), int(reorder), C.byref(area), C.byref(T), C.byref(E)))
        return area.value, T.value, E.value


def shard_partition(sliceptr, world):
    """dsc_shard_partition: row_begin[world + 1] of the library's row partition for a sliced ELL (pure host function)"""
    sp = np.ascontiguousarray(sliceptr, np.int32)
    rb = np.zeros(world + 1, np.int32)
    st = load_library().dsc_shard_partition(_fp(sp), len(sp) - 1, int(world), _fp(rb))
    if st != 0:
        raise DscError(st, "dsc_shard_partition")
    return rb


class Batch:
    """One dsc_batch: many independent frame pairs refined by ONE kernel launch (include/dsc.h, dsc_batch_*)."""

    def __init__(self, device=0):
        self.lib = load_library()
        self.h = C.c_void_p()
        st = self.lib.dsc_batch_create(int(device), C.byref(self.h))
        if st != 0:
            raise DscError(st, self.lib.dsc_status_string(st).decode())
        self.np, self.point_offset = 0, np.zeros(1, np.int64)

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.lib.dsc_batch_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, st):
        if st != 0:
            raise DscError(st, self.lib.dsc_batch_last_error(self.h).decode())

    def upload(self, problems, reorder=1):
        """problems: list of dicts with pair, X1, X2, uv1, uv2, d1, d2, s1, s2, rowptr, col, w, area, ntri (+ optional Tg7,
        isg1, isg2); packed into the concatenated arrays + prefix offsets the C ABI takes."""
        n_p = len(problems)
        pairs = (BatchPair * max(1, n_p))()
        po, eo = np.zeros(n_p + 1, np.int64), np.zeros(n_p + 1, np.int64)
        for k, pr in enumerate(problems):
            pairs[k].pair = pr["pair"]
            pairs[k].scale1, pairs[k].scale2 = float(pr["s1"]), float(pr["s2"])
            tg = pr.get("Tg7")
            for q in range(7):
                pairs[k].Tg7[q] = float(tg[q]) if tg is not None else 0.0
            pairs[k].area, pairs[k].n_triangles = float(pr["area"]), int(pr["ntri"])
            po[k + 1] = po[k] + len(pr["X1"])
            eo[k + 1] = eo[k] + len(pr["col"])

        def cat(key, dt, shape):
            if n_p == 0:
                return np.zeros(tuple(0 if d < 0 else d for d in shape), dt)
            return np.ascontiguousarray(np.concatenate([np.asarray(pr[key], dt).reshape(shape) for pr in problems]))
        X1, X2 = cat("X1", np.float32, (-1, 3)), cat("X2", np.float32, (-1, 3))
        uv1, uv2 = cat("uv1", np.float32, (-1, 2)), cat("uv2", np.float32, (-1, 2))
        d1, d2 = cat("d1", np.float64, (-1,)), cat("d2", np.float64, (-1,))
        has_isg = n_p > 0 and all(pr.get("isg1") is not None for pr in problems)
        s1 = cat("isg1", np.float32, (-1,)) if has_isg else None
        s2 = cat("isg2", np.float32, (-1,)) if has_isg else None
        rowptr, col, w = cat("rowptr", np.int32, (-1,)), cat("col", np.int32, (-1,)), cat("w", np.float64, (-1,))
        self._ck(self.lib.dsc_batch_upload(self.h, n_p, pairs, _fp(po), _fp(X1), _fp(X2), _fp(uv1), _fp(uv2), _fp(d1), _fp(d2),
                                           _fp(s1), _fp(s2), _fp(eo), _fp(rowptr), _fp(col), _fp(w), int(reorder)))
        self.np, self.point_offset = n_p, po
        return int(X1.nbytes + X2.nbytes + uv1.nbytes + uv2.nbytes + d1.nbytes + d2.nbytes + rowptr.nbytes + col.nbytes + w.nbytes)

    def set_pcg(self, rtol=1e-10, max_iters=4000, check_every=32):
        p = PcgParams(float(rtol), int(max_iters), int(check_every))
        self._ck(self.lib.dsc_batch_set_pcg(self.h, C.byref(p)))

    def set_early_reject(self, rtol_loose=(1e-3, 1e-4), rho_margin=(1.0, 0.5)):
        r = np.ascontiguousarray(np.atleast_1d(rtol_loose), np.float64)
        m = np.ascontiguousarray(np.atleast_1d(rho_margin), np.float64)
        self._ck(self.lib.dsc_batch_set_early_reject(self.h, len(r), _fp(r), _fp(m)))

    def reset_state(self):
        self._ck(self.lib.dsc_batch_reset_state(self.h))

    def set_active(self, n_active=-1):
        """the next optimize / reset_state / pixel_sigma work on pairs [0, n_active) only (-1: all)"""
        self._ck(self.lib.dsc_batch_set_active(self.h, int(n_active)))
        self.active = int(n_active)

    def pixel_sigma(self):
        k = self.np if getattr(self, "active", -1) < 0 else min(self.active, self.np)
        out = np.empty((max(1, k), 2))
        self._ck(self.lib.dsc_batch_pixel_sigma(self.h, _fp(out)))
        return out[:k]

    def optimize(self, weights, n_iters):
        """weights: one Weights (shared) or a list with one per pair -> (records[pair][iteration], stats[pair], device ms)"""
        ws = weights if isinstance(weights, (list, tuple)) else [weights]
        warr = (Weights * len(ws))(*ws)
        recs = (IterRecord * max(1, self.np * max(1, n_iters)))()
        stats = (OptStats * max(1, self.np))()
        ms = C.c_double()
        self._ck(self.lib.dsc_batch_optimize(self.h, warr, len(ws), int(n_iters), recs, stats, C.byref(ms)))
        out = [[recs[p * n_iters + i] for i in range(stats[p].iterations)] for p in range(self.np)]
        return out, [stats[p] for p in range(self.np)], ms.value

    def download(self):
        n = int(self.point_offset[-1])
        X1, X2 = np.empty((n, 3), np.float32), np.empty((n, 3), np.float32)
        scales, tg, upd = np.empty((self.np, 2)), np.empty((self.np, 7)), np.empty(self.np)
        self._ck(self.lib.dsc_batch_download(self.h, _fp(X1), _fp(X2), _fp(scales), _fp(tg), _fp(upd)))
        po = self.point_offset
        return [dict(X1=X1[po[p]:po[p + 1]], X2=X2[po[p]:po[p + 1]], scales=scales[p], Tg=tg[p], update=upd[p]) for p in range(self.np)]

    def size(self):
        a, b, c, d = C.c_int(), C.c_longlong(), C.c_int(), C.c_int()
        self._ck(self.lib.dsc_batch_size(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return dict(problems=a.value, points=b.value, cluster_ctas=c.value, clusters=d.value)


class BundleAdjuster:
    """dsc_ba_*: the classic bundle-adjustment paths (g2oBundleAdjustment.cc:38-444) -- poses, points, reprojection edges."""

    def __init__(self, device=0):
        self.lib = load_library()
        self.h = C.c_void_p()
        st = self.lib.dsc_ba_create(int(device), C.byref(self.h))
        if st != 0:
            raise DscError(st, self.lib.dsc_status_string(st).decode())
        self.K = self.M = self.O = 0

    def close(self):
        if self.h:
            self.lib.dsc_ba_destroy(self.h)
            self.h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, st):
        if st != 0:
            raise DscError(st, self.lib.dsc_ba_last_error(self.h).decode())

    def upload(self, poses7, pose_fixed, cams, X, obs_pose, obs_point, obs_uv, obs_isg=None, points_fixed=False):
        """poses7 (K,7) q xyzw + t of Tcw; cams: list of (model, params[8]); X (M,3); observations (pose, point, uv, invSigma2)"""
        p7 = np.ascontiguousarray(poses7, np.float64).reshape(-1, 7)
        self.K, X = len(p7), np.ascontiguousarray(X, np.float64).reshape(-1, 3)
        self.M = len(X)
        fx = np.ascontiguousarray(np.asarray(pose_fixed, bool).astype(np.uint8))
        carr = (Camera * self.K)()
        for k, (model, params) in enumerate(cams):
            carr[k].model = int(model)
            for q in range(8):
                carr[k].params[q] = float(params[q])
        op, oj = np.ascontiguousarray(obs_pose, np.int32), np.ascontiguousarray(obs_point, np.int32)
        uv = np.ascontiguousarray(obs_uv, np.float32).reshape(-1, 2)
        isg = None if obs_isg is None else np.ascontiguousarray(obs_isg, np.float32)
        self.O = len(op)
        self._ck(self.lib.dsc_ba_upload(self.h, self.K, _fp(p7), _fp(fx), carr, self.M, _fp(X), int(bool(points_fixed)), C.c_longlong(self.O),
                                        _fp(op), _fp(oj), _fp(uv), _fp(isg)))

    def set_poses(self, poses7):
        p7 = np.ascontiguousarray(poses7, np.float64).reshape(-1, 7)
        self._ck(self.lib.dsc_ba_set_poses(self.h, _fp(p7)))

    def set_levels(self, active=None):
        a = None if active is None else np.ascontiguousarray(np.asarray(active, bool).astype(np.uint8))
        self._ck(self.lib.dsc_ba_set_levels(self.h, _fp(a)))

    def optimize(self, n_iters, huber_delta=0.0):
        recs = (IterRecord * max(1, n_iters))()
        st = OptStats()
        self._ck(self.lib.dsc_ba_optimize(self.h, int(n_iters), C.c_double(huber_delta), recs, C.byref(st)))
        return [recs[i] for i in range(st.iterations)], st

    def edge_chi2(self):
        chi2, pos = np.empty(max(1, self.O)), np.empty(max(1, self.O), np.uint8)
        self._ck(self.lib.dsc_ba_edge_chi2(self.h, _fp(chi2), _fp(pos)))
        return chi2[:self.O], pos[:self.O].astype(bool)

    def download(self):
        p7, X = np.empty((self.K, 7)), np.empty((max(1, self.M), 3))
        self._ck(self.lib.dsc_ba_download(self.h, _fp(p7), _fp(X)))
        return p7, X[:self.M]

    def launch_count(self):
        c = C.c_longlong()
        self._ck(self.lib.dsc_ba_launch_count(self.h, C.byref(c)))
        return c.value
