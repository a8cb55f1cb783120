"""GPU parity (-m gpu) of the test-time fragment pipeline (§8(f) rank 1): plan kernels bit-exact against the reference's
GridSample outputs (tests/golden/fragments.npz) and the numpy oracle, vote kernels against the oracle, and the whole
voting loop through the model."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import fragments_np as F
from test_cpu_fragments import Z, check_plan_against_reference

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def ops(lib):
    from cdsegnet_b200 import ops as _ops
    return _ops


def _np(plan):
    return {k: (v.cpu().numpy() if torch.is_tensor(v) else v) for k, v in plan.items()}


@pytest.mark.parametrize("i", range(4))
def test_plan_vs_reference(ops, i):
    from cdsegnet_b200.fragments import GridSample
    gs = GridSample(grid_size=float(Z[f"c{i}_grid_size"]), hash_type=str(Z[f"c{i}_hash"]), keys=("coord", "color", "normal"), return_grid_coord=True)
    plan = _np(gs.plan(Z[f"c{i}_coord"]))
    assert plan["n_fragments"] == int(Z[f"c{i}_n_fragments"])
    check_plan_against_reference(i, plan, plan["index"].astype(np.int64))
    # and bit-exact against the stable-sort oracle, tie order included
    o = F.grid_sample_plan(Z[f"c{i}_coord"], float(Z[f"c{i}_grid_size"]), str(Z[f"c{i}_hash"]))
    assert np.array_equal(plan["order"], o["order"]) and np.array_equal(plan["count"], o["count"]) and np.array_equal(plan["start"], o["start"])
    assert np.array_equal(plan["index"], F.fragment_index(o))


@pytest.mark.parametrize("n,f64,legacy", [(250000, False, False), (250000, True, False), (40000, False, True), (1, False, False)])
def test_plan_vs_oracle_full_size(ops, n, f64, legacy):
    """raw-scan sizes (ScanNet scenes hold 50k-400k raw points), float64 coordinates (after a TTA rotation), NumPy-1 float32 division"""
    rng = np.random.default_rng(n)
    coord = (rng.random((n, 3)) * np.array([8.0, 6.0, 3.0]) - np.array([4.0, 3.0, 0.2]))
    coord[:, 2] = np.round(coord[:, 2] * 3) / 3 + rng.normal(0, 0.01, n)            # a few dense layers -> several points per voxel
    coord = coord.astype(np.float64 if f64 else np.float32)
    plan = _np(ops.grid_sample_plan(torch.from_numpy(coord).to(DEV), 0.02, "fnv", legacy))
    o = F.grid_sample_plan(coord, 0.02, "fnv", legacy)
    for k in ("grid_coord", "order", "inverse", "count", "start"):
        assert np.array_equal(plan[k], o[k]), k
    assert np.array_equal(plan["key"].astype(np.uint64), o["key"])
    assert plan["n_voxels"] == len(o["count"]) and plan["n_fragments"] == int(o["count"].max())
    idx = ops.fragment_index(torch.from_numpy(plan["order"]).to(DEV), torch.from_numpy(plan["start"]).to(DEV), plan["n_voxels"], plan["n_fragments"])
    assert np.array_equal(idx.cpu().numpy(), F.fragment_index(o))


def test_return_inverse_travels_with_every_part(ops):
    """transform.py:873-875 writes `inverse` into the scene dict inside the fragment loop, so every part (and the caller's dict)
    carries the full-length point -> voxel map"""
    from cdsegnet_b200.fragments import GridSample
    gs = GridSample(grid_size=float(Z["c0_grid_size"]), hash_type=str(Z["c0_hash"]), keys=("coord",), return_inverse=True)
    d = dict(coord=Z["c0_coord"])
    parts = gs(d)
    o = F.grid_sample_plan(Z["c0_coord"], float(Z["c0_grid_size"]), str(Z["c0_hash"]))
    assert np.array_equal(d["inverse"].cpu().numpy(), o["inverse"])
    for part in parts:
        assert part["inverse"].shape[0] == Z["c0_coord"].shape[0] and np.array_equal(part["inverse"].cpu().numpy(), o["inverse"])


def test_fragments_and_collect_vs_reference(ops):
    """fragment 0 through CenterShift(apply_z=False) + Collect + collate_fn == the reference's model input (its own tie order aside:
    compared on the rows both agree on)"""
    from cdsegnet_b200.fragments import GridSample, collect_fragment
    i = 0
    gs = GridSample(grid_size=float(Z[f"c{i}_grid_size"]), hash_type="fnv", keys=("coord", "color", "normal"), return_grid_coord=True)
    parts = gs(dict(coord=Z[f"c{i}_coord"], color=Z[f"c{i}_color"], normal=Z[f"c{i}_normal"], name="scene"))
    assert parts[0]["name"] == "scene" and len(parts) == int(Z[f"c{i}_n_fragments"])
    inp = collect_fragment(parts[0])
    assert np.array_equal(inp["grid_coord"].cpu().numpy(), Z[f"c{i}_in_grid_coord"])
    assert np.array_equal(inp["offset"].cpu().numpy(), Z[f"c{i}_in_offset"])
    same = inp["index"].cpu().numpy() == Z[f"c{i}_in_index"]
    assert same.mean() > 0.9                                        # voxels with one point (and ties that happen to agree)
    assert np.array_equal(inp["feat"].cpu().numpy()[same], Z[f"c{i}_in_feat"][same])
    # CenterShift uses the fragment's own min / max, which can move by a tie: compare up to that common shift
    d = inp["coord"].cpu().numpy()[same] - Z[f"c{i}_in_coord"][same]
    assert np.abs(d - d[0]).max() < 1e-6


@pytest.mark.parametrize("C", (20, 200))
def test_vote_kernels_vs_oracle(ops, C):
    rng = np.random.default_rng(C)
    n = 30000
    frags = [(rng.permutation(n)[:20000].astype(np.int32), (3 * rng.standard_normal((20000, C))).astype(np.float32)) for _ in range(5)]
    pred = torch.zeros((n, C), dtype=torch.float32, device=DEV)
    for idx, lg in frags:
        ops.vote_softmax_add_(pred, torch.from_numpy(lg).to(DEV), torch.from_numpy(idx).to(DEV))
    ref, labels = F.vote(n, C, frags)
    assert np.abs(pred.cpu().numpy() - ref).max() < 1e-5
    got = ops.argmax_rows(pred).cpu().numpy()
    clear = np.sort(ref, 1)[:, -1] - np.sort(ref, 1)[:, -2] > 1e-4         # rows whose winner is not a rounding matter
    assert np.array_equal(got[clear], labels[clear])
    assert np.array_equal(got, pred.cpu().numpy().argmax(1))             # and exactly the argmax of what was accumulated
    x = torch.zeros((4, C), device=DEV); x[1, 3] = 1; x[2, 3] = 1; x[2, 1] = 1
    assert ops.argmax_rows(x).tolist() == [0, 3, 1, 0]                      # lowest index on ties


def test_fragment_voter_end_to_end(ops):
    """the tester's loop (voxelize -> fragments -> model -> softmax votes -> argmax) with TTA rotations on the small dual
    network == the same loop driven by the numpy oracle's plan around the same model calls"""
    import json
    import cdsegnet_b200 as cb
    from cdsegnet_b200.fragments import FragmentVoter, GridSample, collect_fragment, rotate_z
    from oracle.weights import synth_state_dict
    J = json.load(open(os.path.join(GOLDEN, "wrapper.json")))
    seg = cb.build_model(dict(type="DefaultSegmentorV2", backbone=dict(type="PT-v3m1", **dict(J["cfg"], enable_flash=False)), **J["wrapper"]))
    seg.backbone.load_state_dict(synth_state_dict(J["shapes"]), strict=True)
    seg = seg.to(DEV).eval()
    i = 3
    scene = dict(coord=Z[f"c{i}_coord"], color=Z[f"c{i}_color"], normal=Z[f"c{i}_normal"])
    vox = GridSample(grid_size=0.02, hash_type="fnv", keys=("coord", "color", "normal"), return_grid_coord=True)
    voter = FragmentVoter(seg, 20, vox)
    augs = [None, rotate_z(0.5), rotate_z(1.0, 0.95)]
    torch.manual_seed(7)
    pred = voter.votes(scene, augs)
    assert torch.isfinite(pred).all()
    n = len(scene["coord"])
    # every point is voted for the same number of times per augmentation: sum over classes == visits of the point
    visits = np.zeros(n)
    for aug in augs:
        c = scene["coord"] if aug is None else aug(dict(coord=torch.from_numpy(scene["coord"])))["coord"].numpy()
        o = F.grid_sample_plan(c, 0.02, "fnv")
        np.add.at(visits, F.fragment_index(o).ravel(), 1)
    assert np.abs(pred.sum(1).cpu().numpy() - visits).max() < 1e-3
    # replay with the oracle's fragments around the same model (same seed -> same noise / shuffles)
    torch.manual_seed(7)
    ref = np.zeros((n, 20))
    for aug in augs:
        d = {k: torch.from_numpy(v).to(DEV) for k, v in scene.items()}
        if aug is not None:
            d = aug(d)
        o = F.grid_sample_plan(d["coord"].cpu().numpy(), 0.02, "fnv")
        for idx in F.fragment_index(o):
            il = torch.from_numpy(idx).to(DEV)
            part = dict(index=il, grid_coord=torch.from_numpy(o["grid_coord"][idx]).int().to(DEV), coord=d["coord"][il], color=d["color"][il], normal=d["normal"][il])
            lg = seg.inference(collect_fragment(part), eval=False)["seg_logits"].cpu().numpy()
            ref[idx] += F.softmax(lg.astype(np.float64))
    assert np.abs(pred.cpu().numpy() - ref).max() < 1e-4
    labels = voter(scene, augs=[None])
    assert labels.shape == (n,) and labels.dtype == torch.int64
