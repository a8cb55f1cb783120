"""CPU suite for the fragment pipeline (§8(f) rank 1): oracle/fragments_np.py against the outputs of the reference's own
GridSample / collate_fn (tests/golden/fragments.npz, made by tests/golden/make_golden_fragments.py)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import fragments_np as F

Z = np.load(os.path.join(GOLDEN, "fragments.npz"))


def test_fnv_known_answers():
    """FNV64-1A as the reference writes it (multiply, then xor): offset basis for an empty row, hand-computed single column"""
    assert F.fnv_hash_vec(np.zeros((1, 0), dtype=np.int64))[0] == np.uint64(14695981039346656037)
    h = (14695981039346656037 * 1099511628211) % 2 ** 64 ^ 5
    assert F.fnv_hash_vec(np.array([[5]]))[0] == np.uint64(h)


def check_plan_against_reference(i, plan, index):
    """everything GridSample determines; `index` rows may differ from the reference only by WHICH point of a voxel comes when"""
    n = len(Z[f"c{i}_coord"])
    assert np.array_equal(np.asarray(plan["grid_coord"]), Z[f"c{i}_grid"])
    assert np.array_equal(np.asarray(plan["key"]).astype(np.uint64), Z[f"c{i}_key"])
    assert np.array_equal(np.asarray(plan["inverse"]), Z[f"c{i}_inverse"])
    ref_index = Z[f"c{i}_index"]
    assert index.shape == ref_index.shape
    inv = Z[f"c{i}_inverse"]
    V = inv.max() + 1
    for f in range(len(index)):
        assert np.array_equal(inv[index[f]], np.arange(V))              # fragment f holds exactly one point of every voxel, in voxel order
        assert np.array_equal(inv[ref_index[f]], np.arange(V))
        assert np.array_equal(Z[f"c{i}_grid"][index[f]], Z[f"c{i}_part_grid"][f])   # its grid_coord is therefore identical
    for arr in (index, ref_index):
        assert np.array_equal(np.unique(arr), np.arange(n))              # all fragments together cover every point
    # per voxel, the points visited over the fragments form the same multiset up to the cyclic start the tie order implies
    cnt = np.bincount(inv)
    for v in np.random.default_rng(0).choice(V, size=min(V, 200), replace=False):
        assert sorted(set(index[:, v])) == sorted(set(ref_index[:, v])) == sorted(np.flatnonzero(inv == v))
        assert len(set(index[:cnt[v], v])) == cnt[v]


@pytest.mark.parametrize("i", range(4))
def test_oracle_plan_vs_reference(i):
    plan = F.grid_sample_plan(Z[f"c{i}_coord"], float(Z[f"c{i}_grid_size"]), str(Z[f"c{i}_hash"]))
    check_plan_against_reference(i, plan, F.fragment_index(plan))


def test_oracle_vote():
    rng = np.random.default_rng(1)
    frags = [(rng.permutation(50)[:30], rng.standard_normal((30, 5)).astype(np.float32)) for _ in range(4)]
    pred, labels = F.vote(50, 5, frags)
    ref = np.zeros((50, 5))
    for idx, lg in frags:
        e = np.exp(lg.astype(np.float64)); ref[idx] += e / e.sum(1, keepdims=True)
    assert np.allclose(pred, ref) and np.array_equal(labels, ref.argmax(1))


def test_knn_oracle_small_known_answers():
    from oracle import knn_np as K
    xyz = np.array([[0, 0, 0], [1, 0, 0], [0, 2, 0], [5, 5, 5], [5, 5, 6]], dtype=np.float32)
    off = np.array([3, 5], dtype=np.int32)
    q = np.array([[0.9, 0, 0], [5, 5, 5.4]], dtype=np.float32)
    idx, d2 = K.knn_query(2, xyz, off, q, np.array([1, 2], dtype=np.int32))
    assert idx.tolist() == [[1, 0], [3, 4]]                       # neighbours never cross the batch boundary
    assert np.allclose(d2, [[0.01, 0.81], [0.16, 0.36]], atol=1e-6)
    idx, d2 = K.knn_query(4, xyz, off, q, np.array([1, 2], dtype=np.int32))
    assert idx.tolist() == [[1, 0, 2, -1], [3, 4, -1, -1]] and d2[1, 2] == np.float32(1e10)


def test_collect_fragment_host_logic_vs_reference():
    """CenterShift(apply_z=False) + Collect(keys=(coord, grid_coord, index), feat_keys=(color, normal)) + collate_fn of one fragment
    (transform.py:142-155, 26-50; datasets/utils.py:15-41) -- the host-side mirror in cdsegnet_b200.fragments on CPU tensors, fed with the
    REFERENCE's own fragment 0 (its tie order), against the reference's model input"""
    import torch
    from cdsegnet_b200.fragments import collect_fragment, rotate_z
    for i in range(4):
        idx = Z[f"c{i}_index"][0]
        part = dict(index=torch.from_numpy(idx), grid_coord=torch.from_numpy(Z[f"c{i}_grid"][idx]),
                    coord=torch.from_numpy(Z[f"c{i}_coord"][idx]), color=torch.from_numpy(Z[f"c{i}_color"][idx]),
                    normal=torch.from_numpy(Z[f"c{i}_normal"][idx]))
        inp = collect_fragment(part)
        assert np.array_equal(inp["coord"].numpy(), Z[f"c{i}_in_coord"])           # float32 arithmetic like numpy's: bit-exact
        assert np.array_equal(inp["feat"].numpy(), Z[f"c{i}_in_feat"])
        assert np.array_equal(inp["grid_coord"].numpy(), Z[f"c{i}_in_grid_coord"])
        assert np.array_equal(inp["index"].numpy(), Z[f"c{i}_in_index"])
        assert np.array_equal(inp["offset"].numpy(), Z[f"c{i}_in_offset"])
    # the TTA rotation: float64 result like np.dot(float32, float64) in RandomRotateTargetAngle (transform.py:259-294)
    c = torch.from_numpy(Z["c0_coord"])
    out = rotate_z(0.5)(dict(coord=c, normal=torch.from_numpy(Z["c0_normal"])))
    a = 0.5 * np.pi
    R = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
    assert out["coord"].dtype == torch.float64 and np.allclose(out["coord"].numpy(), np.dot(Z["c0_coord"], R.T), rtol=0, atol=1e-12)
    assert np.allclose(out["normal"].numpy(), np.dot(Z["c0_normal"], R.T), atol=1e-12)
