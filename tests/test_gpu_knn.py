"""GPU parity (-m gpu) of pointops.knn_query (§8(f) rank 4): the cell-grid kernel against the numpy oracle AND against the
reference's own CUDA kernel (oracle/_ref/libref_knn.so, compiled by oracle/Makefile from /root/reference/libs/pointops)."""
import ctypes
import os

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle import knn_np as K

pytestmark = pytest.mark.gpu
DEV = "cuda"
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_knn.so")


def _surface(n, seed, ext=(8.0, 6.0, 3.0)):
    rng = np.random.default_rng(seed)
    p = rng.random((n, 3)) * np.array(ext)
    face = rng.integers(0, 3, n)
    p[face == 0, 2] = 0; p[face == 1, 0] = 0; p[face == 2, 1] = ext[1]
    return (p + rng.normal(0, 0.003, p.shape)).astype(np.float32)


def _ref_kernel(nsample, xyz, off, new_xyz, new_off):
    """the reference's launcher on raw device pointers (legacy default stream, caller-allocated zeroed outputs: query.py:19-24)"""
    # the reference header declares an at::Tensor overload, so the object references a few c10 / ATen symbols: make torch's
    # libraries global first
    for so in ("libc10.so", "libtorch_cpu.so"):
        ctypes.CDLL(os.path.join(os.path.dirname(torch.__file__), "lib", so), mode=ctypes.RTLD_GLOBAL)
    lib = ctypes.CDLL(REF_SO)
    P = ctypes.c_void_p
    lib.knn_query_cuda_launcher.argtypes = [ctypes.c_int, ctypes.c_int, P, P, P, P, P, P]
    lib.knn_query_cuda_launcher.restype = None
    m = new_xyz.shape[0]
    idx = torch.zeros((m, nsample), dtype=torch.int32, device=DEV)
    d2 = torch.zeros((m, nsample), dtype=torch.float32, device=DEV)
    torch.cuda.synchronize()
    lib.knn_query_cuda_launcher(m, nsample, xyz.data_ptr(), new_xyz.data_ptr(), off.data_ptr(), new_off.data_ptr(), idx.data_ptr(), d2.data_ptr())
    torch.cuda.synchronize()
    return idx, d2


CASES = [((3000,), (5000,), 1), ((3000, 1200), (2500, 4000), 1), ((3000, 1200), (2500, 4000), 8), ((40,), (100,), 16), ((5,), (7,), 8),
         ((2000,), (300,), 33)]


@pytest.mark.parametrize("ns,ms,k", CASES)
def test_knn_vs_oracle_and_reference_kernel(lib, ns, ms, k):
    from cdsegnet_b200 import pointops
    xyz = np.concatenate([_surface(n, 10 + i) for i, n in enumerate(ns)])
    q = np.concatenate([_surface(m, 50 + i) + np.float32(0.01) for i, m in enumerate(ms)])
    q[:3] += 5.0                                                   # a few queries far outside the data's bounding box
    off, noff = np.cumsum(ns).astype(np.int32), np.cumsum(ms).astype(np.int32)
    t = lambda a: torch.from_numpy(a).to(DEV)
    idx, dist = pointops.knn_query(k, t(xyz), t(off), t(q), t(noff))
    assert idx.dtype == torch.int32 and dist.dtype == torch.float32 and idx.shape == (len(q), k)
    oi, od2 = K.knn_query(k, xyz, off, q, noff)
    got_i, got_d = idx.cpu().numpy(), dist.cpu().numpy()
    assert np.allclose(got_d, np.sqrt(od2), rtol=1e-5, atol=1e-7)
    agree = got_i == oi
    if not agree.all():                                            # only where two candidates are within rounding of each other
        bad = ~agree
        assert np.allclose(np.sqrt(od2)[bad], got_d[bad], rtol=1e-5, atol=1e-7) and bad.mean() < 1e-3
    if os.path.exists(REF_SO):
        ri, rd2 = _ref_kernel(k, t(xyz), t(off), t(q), t(noff))
        ri, rd2 = ri.cpu().numpy(), rd2.cpu().numpy()
        assert np.array_equal(np.sqrt(rd2), got_d)               # same expression, same compiler: bit-exact distances
        if k == 1:
            assert np.array_equal(ri, got_i)                       # strict `<` keeps the lowest index on ties, like ours
        else:
            assert (ri == got_i).mean() > 0.999 and np.array_equal(np.sort(ri, 1)[rd2[:, -1] < 1e9], np.sort(got_i, 1)[rd2[:, -1] < 1e9])


def test_knn_duplicates_and_self_query(lib):
    """duplicated points (exact ties) resolve to the lower index; self-query returns the point itself at distance 0"""
    from cdsegnet_b200 import pointops
    base = _surface(4000, 3)
    xyz = np.concatenate([base, base[:500]])                       # rows 4000.. duplicate rows 0..499
    off = np.array([len(xyz)], dtype=np.int32)
    t = lambda a: torch.from_numpy(a).to(DEV)
    idx, dist = pointops.knn_query(1, t(xyz), t(off))
    got = idx.cpu().numpy()[:, 0]
    exp = np.arange(len(xyz)); exp[4000:] = np.arange(500)
    assert np.array_equal(got, exp) and float(dist.abs().max()) == 0.0
    if os.path.exists(REF_SO):
        ri, _ = _ref_kernel(1, t(xyz), t(off), t(xyz), t(off))
        assert np.array_equal(ri.cpu().numpy()[:, 0], got)


def test_knn_evaluator_sized(lib):
    """the evaluator's call (evaluator.py:132-141): 120k voxel centres, 250k original points, k = 1 -- against the reference
    kernel when it is available, else the oracle on a slice"""
    from cdsegnet_b200 import pointops
    xyz, q = _surface(120000, 1), _surface(250000, 2)
    off, noff = np.array([120000], dtype=np.int32), np.array([250000], dtype=np.int32)
    t = lambda a: torch.from_numpy(a).to(DEV)
    idx, dist = pointops.knn_query(1, t(xyz), t(off), t(q), t(noff))
    if os.path.exists(REF_SO):
        ri, rd2 = _ref_kernel(1, t(xyz), t(off), t(q), t(noff))
        assert torch.equal(ri, idx) and torch.equal(torch.sqrt(rd2), dist)
    oi, od2 = K.knn_query(1, xyz, off, q[:4000], np.array([4000], dtype=np.int32))
    assert np.allclose(dist.cpu().numpy()[:4000], np.sqrt(od2), rtol=1e-5, atol=1e-7)
    assert (idx.cpu().numpy()[:4000] == oi).mean() > 0.999
