"""CPU suite (-m "not gpu"): the oracle against the golden vectors produced by the reference's
own sources, the host-side logic, and the C-ABI library surface (no compute without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT
from helpers import load_case, oracle_forward
from oracle import serialization_np as S

ORDERS = ("z", "z-trans", "hilbert", "hilbert-trans")
DEPTHS = (1, 2, 3, 5, 8, 9, 11, 16)


def test_codes_known_answers():
    # SURVEY.md §8c known-answer vectors extracted from the reference sources
    assert S.encode(np.array([[1, 0, 0]]), None, 1, "z")[0] == 4
    assert S.encode(np.array([[0, 0, 1]]), None, 1, "z")[0] == 1
    cube = np.array([(0, 0, 0), (0, 0, 1), (0, 1, 1), (0, 1, 0), (1, 1, 0), (1, 1, 1), (1, 0, 1), (1, 0, 0)])
    assert S.encode(cube, None, 1, "hilbert").tolist() == list(range(8))
    z = np.load(os.path.join(GOLDEN, "codes.npz"))
    g = torch.randint(0, 512, (120000, 3), generator=torch.Generator().manual_seed(0), dtype=torch.int32).numpy()
    assert S.encode(g[:3], None, 9, "z").tolist() == [8887763, 45912603, 43582135] == z["kat_first3_z"].tolist()
    assert S.encode(g[:3], None, 9, "hilbert").tolist() == [8320106, 50715426, 66766583] == z["kat_first3_hilbert"].tolist()


@pytest.mark.parametrize("depth", DEPTHS)
def test_codes_vs_reference(depth):
    z = np.load(os.path.join(GOLDEN, "codes.npz"))
    for o in ORDERS:
        got = S.encode(z[f"grid_{depth}"], z[f"batch_{depth}"], depth, o)
        assert np.array_equal(got, z[f"code_{depth}_{o}"]), (depth, o)


def test_serialization_vs_reference():
    z = np.load(os.path.join(GOLDEN, "serialization.npz"))
    batch = S.offset2batch(z["offset"])
    assert np.array_equal(batch, z["batch"])
    code, order, inverse, depth = S.serialization(z["grid_coord"], batch)
    assert depth == int(z["depth"])
    assert np.array_equal(code, z["code"]) and np.array_equal(order, z["order"]) and np.array_equal(inverse, z["inverse"])


def test_padding_vs_reference():
    z = np.load(os.path.join(GOLDEN, "padding.npz"))
    for i in range(int(z["n_cases"])):
        pad, unpad, cu = S.patch_maps(z[f"offset_{i}"], int(z[f"K_{i}"]))
        assert np.array_equal(pad, z[f"pad_{i}"]) and np.array_equal(unpad, z[f"unpad_{i}"]) and np.array_equal(cu, z[f"cu_{i}"]), i
    # SURVEY.md §8c vectors
    pad, unpad, cu = S.patch_maps([5, 12], 4)
    assert pad.tolist() == [0, 1, 2, 3, 4, 1, 2, 3, 5, 6, 7, 8, 9, 10, 11, 8]
    assert unpad.tolist() == [0, 1, 2, 3, 4, 8, 9, 10, 11, 12, 13, 14] and cu.tolist() == [0, 4, 8, 12, 16]


def test_pool_plan_properties():
    from cdsegnet_b200 import synth
    sc = synth.collate([synth.small_room(1500, 1), synth.small_room(900, 2)])
    batch = S.offset2batch(sc["offset"])
    code, order, inverse, depth = S.serialization(sc["grid_coord"], batch)
    pl = S.pool_plan(code, 2, depth)
    m = len(pl["counts"])
    assert pl["idx_ptr"][-1] == len(batch) and pl["counts"].sum() == len(batch)
    # every member of a cluster shares the shifted code on EVERY curve (hierarchical curves)
    sh = code >> 3
    for r in range(4):
        assert np.array_equal(sh[r], pl["code"][r][pl["cluster"]])
        assert np.array_equal(np.sort(pl["order"][r]), np.arange(m))
        assert (np.diff(pl["code"][r][pl["order"][r]]) > 0).all()


@pytest.mark.parametrize("name", ["case1_single", "case2_batch2", "case3_cn_only", "case4_nuscenes"])
def test_oracle_vs_reference_forward(name):
    """the functional fp32 oracle reproduces the reference network's outputs (dense branch)"""
    z, cfg, shapes = load_case(name)
    c, n = oracle_forward(z, cfg, shapes, "dense")
    assert np.abs(n - z["n_feat"]).max() < 1e-4
    if cfg["condition"]:
        assert np.abs(c - z["c_feat"]).max() < 1e-4


def test_oracle_rng_replay():
    """drawing the shuffles from torch's CPU generator with the recorded seed reproduces the reference"""
    z, cfg, shapes = load_case("case3_cn_only")
    from helpers import t
    from oracle import ptv3_oracle as O
    from oracle.weights import synth_state_dict
    torch.manual_seed(int(z["seed"]))
    n = O.forward(synth_state_dict(shapes), cfg, n_in=dict(coord=t(z["coord"]), grid_coord=t(z["grid_coord"]).long(),
                                                            offset=t(z["offset"]), feat=t(z["feat"])))
    assert np.abs(n["feat"].numpy() - z["n_feat"]).max() < 1e-4


def test_flash16_emulation_close_to_dense():
    z, cfg, shapes = load_case("case3_cn_only")
    _, n16 = oracle_forward(z, cfg, shapes, "flash16")
    assert np.abs(n16 - z["n_feat"]).max() < 2e-2      # fp16 attention operands vs fp32: the reference's own flash/dense gap


# ------------------------------------------------------------------ host logic / boundary
def test_state_dict_contract():
    import cdsegnet_b200 as cb
    for name in ("case1_single", "case3_cn_only"):
        _, cfg, shapes = load_case(name)
        sd = cb.PointTransformerV3(**cfg).state_dict()
        assert {k: list(v.shape) for k, v in sd.items()} == shapes


def test_registry_keys():
    import cdsegnet_b200 as cb
    _, cfg, _ = load_case("case3_cn_only")
    m = cb.build_model(dict(type="DefaultSegmentorV2", backbone=dict(type="PT-v3m1", **cfg), condition=False))
    assert isinstance(m.backbone, cb.PointTransformerV3)


def test_cabi_exports_every_declared_symbol(lib):
    from cdsegnet_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "cdseg_b200.h")).read()
    declared = set(re.findall(r"\b(cdseg_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name)
    assert lib.cdseg_abi_version() == 3


@pytest.mark.parametrize("depth", DEPTHS)
def test_device_bit_routines_on_host(lib, depth):
    """the __host__ __device__ Morton/Hilbert routines of serialize.cu, evaluated on the host"""
    z = np.load(os.path.join(GOLDEN, "codes.npz"))
    g = np.ascontiguousarray(z[f"grid_{depth}"].astype(np.int32))
    b = np.ascontiguousarray(z[f"batch_{depth}"].astype(np.int32))
    for oid, o in enumerate(ORDERS):
        out = np.zeros(len(g), np.int64)
        st = lib.cdseg_debug_encode_host(g.ctypes.data, b.ctypes.data, len(g), depth, oid, out.ctypes.data)
        assert st == 0 and np.array_equal(out, z[f"code_{depth}_{o}"])


def test_product_fails_loudly_without_cuda():
    import cdsegnet_b200 as cb
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    z, cfg, _ = load_case("case3_cn_only")
    from helpers import t
    m = cb.PointTransformerV3(**cfg).eval()
    with pytest.raises(Exception):
        m(n_point=dict(coord=t(z["coord"]), grid_coord=t(z["grid_coord"]), offset=t(z["offset"]), feat=t(z["feat"])))


def test_ctypes_struct_mirrors_match_the_c_layouts(lib):
    """every struct of the C ABI mirrored in Python (ctypes) has the size the compiler gave it"""
    import ctypes
    from cdsegnet_b200 import _lib, netexec
    out = (ctypes.c_size_t * 16)()
    n = lib.cdseg_struct_sizes(out, 16)
    mirrors = [_lib.BlockArgs, _lib.PatchMap, _lib.PlanLevel, netexec.LinW, netexec.LnW, netexec.BlockW, netexec.PoolW, netexec.UnpoolW,
               netexec.StageW, netexec.StemW, netexec.CrossW, netexec.NetW, netexec.ForwardArgs]
    assert n == len(mirrors)
    for i, m in enumerate(mirrors):
        assert ctypes.sizeof(m) == out[i], (m.__name__, ctypes.sizeof(m), out[i])


def test_split_heuristic_is_the_same_on_all_three_launch_paths(lib):
    """net_exec.cu, block_exec.cu and the per-module Python path each pick the split-K / tap-split of a launch; the summation order --
    and with it bit-identity between the paths -- depends on all three agreeing"""
    from cdsegnet_b200 import ops
    for T in (1, 4, 8, 16, 27, 32):
        for tiles in list(range(1, 130)) + [200, 938]:
            a, b, c = lib.cdseg_debug_pick_split(tiles, T), lib.cdseg_debug_pick_split_block(tiles, T), ops.pick_split(tiles, T)
            assert a == b == c, (tiles, T, a, b, c)
            assert 1 <= a <= T
