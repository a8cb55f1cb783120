"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): numpy restatement of pointops.knn_query
(libs/pointops/src/knn_query/knn_query_cuda_kernel.cu:60-104; python wrapper libs/pointops/functions/query.py:8-29).

Pinned on the GPU box against the reference's OWN kernel: oracle/Makefile compiles knn_query_cuda_kernel.cu from where
it lies under /root/reference into oracle/_ref/libref_knn.so and tests/test_gpu_knn.py runs both on the same inputs.
(On the CPU the reference kernel cannot run; this restatement is then the checker.)
"""
import numpy as np


def knn_query(nsample, xyz, offset, new_xyz=None, new_offset=None, chunk=2048):
    """-> idx int32 [m, nsample] (-1 placeholder), dist2 fp32 [m, nsample] (1e10 placeholder); squared distances in fp32
    with the kernel's expression (without FMA contraction: last-ulp differences against the GPU are possible);
    ties -> lower index first (the kernel's strict `<` for nsample = 1)"""
    if new_xyz is None:
        new_xyz, new_offset = xyz, offset
    xyz, new_xyz = np.asarray(xyz, np.float32), np.asarray(new_xyz, np.float32)
    m = len(new_xyz)
    idx = np.full((m, nsample), -1, dtype=np.int32)
    d2o = np.full((m, nsample), 1e10, dtype=np.float32)
    s = ns = 0
    for e, ne in zip(np.asarray(offset).tolist(), np.asarray(new_offset).tolist()):
        pts = xyz[s:e]
        for q0 in range(ns, ne, chunk):
            q = new_xyz[q0:min(q0 + chunk, ne)]
            if len(pts):
                dx = q[:, None, 0] - pts[None, :, 0]
                dy = q[:, None, 1] - pts[None, :, 1]
                dz = q[:, None, 2] - pts[None, :, 2]
                d2 = (dx * dx + dy * dy) + dz * dz
                k = min(nsample, len(pts))
                order = np.lexsort((np.broadcast_to(np.arange(len(pts)), d2.shape), d2), axis=1)[:, :k]
                idx[q0:q0 + len(q), :k] = order + s
                d2o[q0:q0 + len(q), :k] = np.take_along_axis(d2, order, 1)
        s, ns = e, ne
    return idx, d2o
