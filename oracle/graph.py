"""Neighbour graph, cotangent weights and per-vertex rotations (oracle = test infrastructure).

Follows /root/reference/Modules/Utils/Geometry.cc:258-368 (extractPositions,
ComputeEdgeWeightsCot, createVectorMap, ComputeDelaunayTriangulation3D) and
:549-604 (computeR), as called from g2oBundleAdjustment.cc:653-688.

Qhull / Open3D are absent from /root/reference; scipy.spatial.Delaunay is the
same Qhull library and is called with the reference's options ("Qbb Qt";
"d" is implied).  Open3D's ComputeAdjacencyList / GetEdgeToVerticesMap /
GetSurfaceArea are restated from their published behaviour: adjacency = union of
triangle edges, edge->opposite-vertex lists, area = sum of 3-D triangle areas.

Graph container (what the C ABI's dsc_set_graph takes):
  rowptr (N+1) int32, col (E) int32 ascending inside a row (the reference iterates
  an unordered_set, i.e. an unspecified order), w (E) float64 symmetric,
  area float64, n_triangles int.
"""
import numpy as np
from dataclasses import dataclass


@dataclass
class Graph:
    rowptr: np.ndarray
    col: np.ndarray
    w: np.ndarray
    area: float
    n_triangles: int
    triangles: np.ndarray = None

    @property
    def n(self):
        return len(self.rowptr) - 1

    @property
    def n_edges(self):
        return len(self.col)

    def rows(self):
        return np.repeat(np.arange(self.n, dtype=np.int32), np.diff(self.rowptr))


def _csr_from_pairs(n, i, j):
    """Directed pairs (both directions present) -> CSR with ascending columns, no duplicates."""
    key = np.unique(i.astype(np.int64) * n + j.astype(np.int64))
    ii = (key // n).astype(np.int32)
    jj = (key % n).astype(np.int32)
    rowptr = np.zeros(n + 1, np.int32)
    np.add.at(rowptr, ii + 1, 1)
    rowptr = np.cumsum(rowptr).astype(np.int32)
    return rowptr, jj, ii


def triangle_area_sum(V, tri):
    a = V[tri[:, 1]] - V[tri[:, 0]]
    b = V[tri[:, 2]] - V[tri[:, 0]]
    return float(0.5 * np.linalg.norm(np.cross(a, b), axis=1).sum())


def cot_weights(V, tri, min_weight=0.0):
    """ComputeEdgeWeightsCot (Geometry.cc:272-298): per undirected edge, mean over the
    adjacent triangles of a.b/|a x b| at the opposite vertex, clamped to >= min_weight.
    Returns dict-free arrays: (edge_lo, edge_hi, weight)."""
    e0 = np.concatenate([tri[:, 0], tri[:, 1], tri[:, 2]])
    e1 = np.concatenate([tri[:, 1], tri[:, 2], tri[:, 0]])
    opp = np.concatenate([tri[:, 2], tri[:, 0], tri[:, 1]])
    a = V[e0] - V[opp]
    b = V[e1] - V[opp]
    with np.errstate(all="ignore"):
        cot = np.einsum("ij,ij->i", a, b) / np.linalg.norm(np.cross(a, b), axis=1)
    lo = np.minimum(e0, e1).astype(np.int64)
    hi = np.maximum(e0, e1).astype(np.int64)
    n = V.shape[0]
    key = lo * n + hi
    uniq, inv = np.unique(key, return_inverse=True)
    s = np.zeros(len(uniq))
    c = np.zeros(len(uniq))
    np.add.at(s, inv, cot)
    np.add.at(c, inv, 1.0)
    w = s / c
    w = np.where(w < min_weight, min_weight, w)
    return (uniq // n).astype(np.int32), (uniq % n).astype(np.int32), w


def delaunay_graph(X1):
    """g2oBundleAdjustment.cc:657-662: 2-D Delaunay of world (x,y) of KF1's points,
    adjacency list, cot weights (min_weight 0), mesh area, triangle count."""
    from scipy.spatial import Delaunay
    V = np.asarray(X1, np.float64)
    tri = Delaunay(V[:, :2], qhull_options="Qbb Qt").simplices.astype(np.int32)
    n = V.shape[0]
    lo, hi, w = cot_weights(V, tri, 0.0)
    rowptr, col, row = _csr_from_pairs(n, np.concatenate([lo, hi]), np.concatenate([hi, lo]))
    key_e = lo.astype(np.int64) * n + hi
    key_d = np.minimum(row, col).astype(np.int64) * n + np.maximum(row, col)
    wd = w[np.searchsorted(key_e, key_d)]
    return Graph(rowptr, col, wd.astype(np.float64), triangle_area_sum(V, tri), int(tri.shape[0]), tri)


def knn_graph(X1, k, area, n_triangles=None, weight=1.0):
    """Symmetrised k-nearest-neighbour graph in world (x,y).
    Not in the reference (its graph is Delaunay); unit weights; `area` and
    `n_triangles` (default 2N, the planar-triangulation count) are supplied because
    the ARAP information is arapW * n_triangles^2 and the residual divides by area
    (g2oBundleAdjustment.cc:942-946)."""
    from scipy.spatial import cKDTree
    V = np.asarray(X1, np.float64)
    n = V.shape[0]
    _, idx = cKDTree(V[:, :2]).query(V[:, :2], k=k + 1)
    i = np.repeat(np.arange(n), k + 1)
    j = idx.reshape(-1)
    keep = i != j
    i, j = i[keep], j[keep]
    rowptr, col, _ = _csr_from_pairs(n, np.concatenate([i, j]), np.concatenate([j, i]))
    w = np.full(len(col), float(weight))
    return Graph(rowptr, col, w, float(area), int(2 * n if n_triangles is None else n_triangles))


def compute_rotations(g, X1, X2):
    """computeR (Geometry.cc:549-604): S_i = sum_j w_ij (p1i-p1j)(p2i-p2j)^T,
    S = U S V^T, R_i = V U^T with the det<0 fix on U.col(2).  Vertices without
    neighbours keep the identity (Rs is initialised to identity, g2oBundleAdjustment.cc:687)."""
    X1 = np.asarray(X1, np.float64)
    X2 = np.asarray(X2, np.float64)
    n = g.n
    row = g.rows()
    d1 = X1[row] - X1[g.col]
    d2 = X2[row] - X2[g.col]
    outer = g.w[:, None, None] * d1[:, :, None] * d2[:, None, :]
    S = np.zeros((n, 3, 3))
    np.add.at(S, row, outer)
    U, _, Vh = np.linalg.svd(S)
    V = np.swapaxes(Vh, 1, 2)
    R = V @ np.swapaxes(U, 1, 2)
    neg = np.linalg.det(R) < 0
    U2 = U.copy()
    U2[neg, :, 2] *= -1
    R[neg] = V[neg] @ np.swapaxes(U2[neg], 1, 2)
    deg = np.diff(g.rowptr)
    R[deg == 0] = np.eye(3)
    # Degenerate covariances.  S == 0 keeps the identity.  Rank 1 (e.g. a hull vertex whose cot weights
    # are clamped to 0 on all but one edge): V U^T depends on an arbitrary completion of the SVD bases in
    # the reference; defined here (and in the CUDA kernel) as the minimal rotation taking u1 to v1.
    sv = np.linalg.svd(S, compute_uv=False)
    for i in np.nonzero(~(sv[:, 0] > 0.0))[0]:
        R[i] = np.eye(3)
    for i in np.nonzero((sv[:, 0] > 0.0) & ~(sv[:, 1] > 1e-12 * sv[:, 0]))[0]:
        u, v = U[i, :, 0], V[i, :, 0]
        d = float(u @ v)
        if d > -1.0 + 1e-12:
            c = np.cross(u, v)
            K = np.array([[0, -c[2], c[1]], [c[2], 0, -c[0]], [-c[1], c[0], 0]])
            R[i] = np.eye(3) + K + (K @ K) / (1.0 + d)
        else:
            k = int(np.argmin(np.abs(u)))
            a = np.cross(u, np.eye(3)[k])
            a /= np.linalg.norm(a)
            R[i] = 2.0 * np.outer(a, a) - np.eye(3)
    return R
