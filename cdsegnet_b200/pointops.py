"""Drop-in for the one `pointops` entry point the CDSegNet path reaches (SURVEY.md §8b "import-time obligations", §8(f)
rank 4): `pointcept/engines/hooks/evaluator.py:11` imports `pointops` unconditionally and calls
`pointops.knn_query(1, coord, offset.int(), origin_coord, origin_offset.int())` (evaluator.py:132-141) to map voxel
predictions back to the original points.

    import sys, cdsegnet_b200.pointops as pointops; sys.modules["pointops"] = pointops       

Same signature and return values as libs/pointops/functions/query.py:8-29 (KNNQuery.apply): idx int32 [m, nsample]
(-1 = placeholder), dist fp32 [m, nsample] = sqrt of the squared distances.
"""
import torch

from . import _lib
from .ops import _p, _stream, _ws, check


def knn_query(nsample, xyz, offset, new_xyz=None, new_offset=None):
    if new_xyz is None or new_offset is None:
        new_xyz, new_offset = xyz, offset
    assert xyz.is_contiguous() and new_xyz.is_contiguous()
    lib = _lib.load()
    n, m = xyz.shape[0], new_xyz.shape[0]
    idx = torch.zeros((m, nsample), dtype=torch.int32, device=xyz.device)
    dist2 = torch.zeros((m, nsample), dtype=torch.float32, device=xyz.device)
    off, noff = offset.int().contiguous(), new_offset.int().contiguous()
    nb = lib.cdseg_knn_workspace_bytes(n)
    ws = _ws(nb, xyz.device)
    check(lib.cdseg_knn_query(m, nsample, _p(xyz, torch.float32), _p(new_xyz, torch.float32), _p(off), _p(noff), off.numel(), n,
                              _p(idx), _p(dist2), _p(ws), nb, _stream()), "knn_query")
    return idx, torch.sqrt(dist2)
