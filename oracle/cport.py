"""ctypes front end of the plain-C oracle oracle/c/dsc_oracle.c (oracle = test infrastructure).

The C file restates the same reference code as the numpy modules next to it (citations in its header); it exists
so that (a) the CUDA path is checked against two independently written restatements and (b) bench.py's CPU legs can
time a compiled, multi-threaded CPU implementation at sizes the numpy port cannot reach.  PARITY UNPINNED, like the
rest of oracle/: the reference itself cannot be built in this image.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libdsc_oracle.so")
_lib = None


class _Problem(C.Structure):
    _fields_ = [("n", C.c_int), ("cam_model", C.c_int * 2), ("cam", (C.c_float * 8) * 2), ("T", (C.c_float * 12) * 2),
                ("uv", C.c_void_p * 2), ("isg", C.c_void_p * 2), ("d", C.c_void_p * 2),
                ("rowptr", C.c_void_p), ("col", C.c_void_p), ("w", C.c_void_p), ("area", C.c_double), ("ntri", C.c_int),
                ("R", C.c_void_p), ("X", C.c_void_p * 2), ("Tg", C.c_double * 7), ("s", C.c_double * 2)]


class _Weights(C.Structure):
    _fields_ = [("rep", C.c_double), ("arap", C.c_double), ("depth_sigma", C.c_double)]


class _Options(C.Structure):
    _fields_ = [("fd", C.c_int), ("threads", C.c_int), ("pcg_rtol", C.c_double), ("pcg_max", C.c_int), ("budget_s", C.c_double)]


def build(force=False):
    src = os.path.join(_HERE, "c", "dsc_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", os.path.join(_HERE, "c")] + (["-B"] if force else []), check=True,
                       stdout=subprocess.DEVNULL)
    return LIB_PATH


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        lib = C.CDLL(LIB_PATH)
        lib.dso_cost.restype = C.c_double
        lib.dso_cost.argtypes = [C.POINTER(_Problem), C.POINTER(_Weights), C.c_void_p]
        lib.dso_compute_rotations.argtypes = [C.POINTER(_Problem)]
        lib.dso_optimize.argtypes = [C.POINTER(_Problem), C.POINTER(_Weights), C.POINTER(_Options), C.c_int, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
        lib.dso_time_components.argtypes = [C.POINTER(_Problem), C.POINTER(_Weights), C.c_int, C.c_int, C.c_void_p]
        lib.dso_debug_linearize.argtypes = [C.POINTER(_Problem), C.POINTER(_Weights), C.c_int, C.c_double, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
        _lib = lib
    return _lib


class CProblem:
    """Owns contiguous copies of an oracle.edges.Problem in the C layout."""

    def __init__(self, p, rotations=None):
        self.n = p.n
        g = p.graph
        self.keep = dict(
            uv1=np.ascontiguousarray(p.uv1, np.float32), uv2=np.ascontiguousarray(p.uv2, np.float32),
            isg1=np.ascontiguousarray(p.inv_sigma2_1, np.float64), isg2=np.ascontiguousarray(p.inv_sigma2_2, np.float64),
            d1=np.ascontiguousarray(p.d1, np.float64), d2=np.ascontiguousarray(p.d2, np.float64),
            rowptr=np.ascontiguousarray(g.rowptr, np.int32), col=np.ascontiguousarray(g.col, np.int32),
            w=np.ascontiguousarray(g.w, np.float64),
            R=np.ascontiguousarray(np.zeros((p.n, 9)) if rotations is None else np.asarray(rotations, np.float64).reshape(p.n, 9)),
            X1=np.ascontiguousarray(p.X1, np.float64).copy(), X2=np.ascontiguousarray(p.X2, np.float64).copy())
        k = self.keep
        c = _Problem()
        c.n = p.n
        for idx, (cam, T) in enumerate(((p.cam1, p.T1), (p.cam2, p.T2))):
            c.cam_model[idx] = int(cam[0])
            prm = np.asarray(cam[1], np.float32)
            for q in range(8):
                c.cam[idx][q] = float(prm[q])
            T34 = T.as34().reshape(-1)
            for q in range(12):
                c.T[idx][q] = float(T34[q])
        c.uv[0], c.uv[1] = k["uv1"].ctypes.data, k["uv2"].ctypes.data
        c.isg[0], c.isg[1] = k["isg1"].ctypes.data, k["isg2"].ctypes.data
        c.d[0], c.d[1] = k["d1"].ctypes.data, k["d2"].ctypes.data
        c.rowptr, c.col, c.w = k["rowptr"].ctypes.data, k["col"].ctypes.data, k["w"].ctypes.data
        c.area, c.ntri = float(g.area), int(g.n_triangles)
        c.R = k["R"].ctypes.data
        c.X[0], c.X[1] = k["X1"].ctypes.data, k["X2"].ctypes.data
        tg = p.Tg.as7()
        for q in range(7):
            c.Tg[q] = float(tg[q])
        c.s[0], c.s[1] = float(p.s1), float(p.s2)
        self.c = c

    @property
    def X1(self):
        return self.keep["X1"]

    @property
    def X2(self):
        return self.keep["X2"]

    @property
    def R(self):
        return self.keep["R"].reshape(self.n, 3, 3)

    def scales(self):
        return float(self.c.s[0]), float(self.c.s[1])

    def Tg7(self):
        return np.array([self.c.Tg[q] for q in range(7)])


def _w(w):
    return _Weights(float(w.rep), float(w.arap), float(w.depth_sigma))


def compute_rotations(cp):
    """computeR in C; returns the number of rank-deficient vertices (left as identity)."""
    return load().dso_compute_rotations(C.byref(cp.c))


def cost(cp, w):
    parts = np.zeros(3)
    v = load().dso_cost(C.byref(cp.c), C.byref(_w(w)), parts.ctypes.data)
    return v, parts


def debug_linearize(cp, w, lam, x, fd=False):
    m = 8 + 6 * cp.n
    b, hd, y = np.zeros(m), np.zeros(m), np.zeros(m)
    chi = C.c_double(0)
    x = np.ascontiguousarray(x, np.float64)
    rc = load().dso_debug_linearize(C.byref(cp.c), C.byref(_w(w)), int(fd), float(lam), x.ctypes.data, b.ctypes.data,
                                    hd.ctypes.data, y.ctypes.data, C.byref(chi))
    if rc:
        raise MemoryError("dso_debug_linearize")
    return b, hd, y, chi.value


def optimize(cp, w, iters, fd=False, threads=0, pcg_rtol=1e-12, pcg_max=20000, budget_s=0.0):
    """Runs LM in place on cp; returns dict(chi2, lam, trials, pcg_iters, final_chi2)."""
    chi2 = np.zeros(iters + 1)
    lam = np.zeros(max(iters, 1))
    trials = np.zeros(max(iters, 1), np.int32)
    its = np.zeros(max(iters, 1), np.int32)
    done = C.c_int(0)
    opt = _Options(int(fd), int(threads), float(pcg_rtol), int(pcg_max), float(budget_s))
    rc = load().dso_optimize(C.byref(cp.c), C.byref(_w(w)), C.byref(opt), int(iters), chi2.ctypes.data, lam.ctypes.data,
                             trials.ctypes.data, its.ctypes.data, C.byref(done))
    if rc:
        raise MemoryError("dso_optimize")
    k = done.value
    return dict(chi2=chi2[:k].tolist(), lam=lam[:k].tolist(), trials=trials[:k].tolist(), pcg_iters=its[:k].tolist(),
                final_chi2=float(chi2[k]))


def time_components(cp, w, pcg_reps=20, threads=0):
    """seconds of (linearisation + cost, one PCG iteration, one cost evaluation, the preconditioner set-up of a solve)"""
    out = np.zeros(4)
    rc = load().dso_time_components(C.byref(cp.c), C.byref(_w(w)), int(threads), int(pcg_reps), out.ctypes.data)
    if rc < 0:
        raise MemoryError("dso_time_components")
    return dict(linearize_s=float(out[0]), pcg_iter_s=float(out[1]), cost_s=float(out[2]), precond_s=float(out[3]))


def threads():
    lib = load()
    lib.dso_threads.restype = C.c_int
    return int(lib.dso_threads())
