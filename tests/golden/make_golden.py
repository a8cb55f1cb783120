"""Generate the golden fixtures in this directory by EXECUTING THE REFERENCE'S OWN
PYTHON SOURCES (read from /root/reference, never copied) on CPU.

Run once in the authoring container:   python tests/golden/make_golden.py
(/root/reference does not exist on the GPU box; the tests only read the fixtures.)

What runs unmodified from the reference:
  * pointcept/models/utils/serialization/{z_order,hilbert,default}.py
  * pointcept/models/utils/{structure,misc}.py           (Point.serialization ...)
  * pointcept/models/modules.py                          (PointSequential)
  * pointcept/models/point_transformer_v3/point_transformer_v3m1_base.py (whole dual network)
  * pointcept/utils/{registry,comm}.py                   (Registry, calc_t_emb)
What is shimmed because the package is not installed / not vendored:
  addict.Dict, timm DropPath (identity in eval), spconv.SubMConv3d +
  SparseConvTensor, torch_scatter.segment_csr.  The spconv/torch_scatter shims
  implement the published semantics of those packages -> "parity unpinned" for
  the conv tap order only.

Fixtures written:
  codes.npz       grid coords + reference codes for the 4 curves at depths 1..16
  padding.npz     reference get_padding_and_inverse() outputs for several offsets/K
  ptv3_case*.npz  inputs + reference outputs (+ serialization/pool traces) of the
                  full dual network on small clouds; weights = oracle/weights.py
  shapes_*.json   parameter name -> shape tables of the reference model
"""
import os, sys, types, json, importlib
import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle import ptv3_oracle as O          # only for the third-party op restatements used by the shims
from oracle.weights import synth_state_dict
from cdsegnet_b200 import synth


# ----------------------------------------------------------------------------
# shims for packages that are not installed here
# ----------------------------------------------------------------------------
def install_shims():
    # namespace packages so that pointcept/models/__init__.py (imports every backbone) never runs
    for name, path in (("pointcept", f"{REF}/pointcept"), ("pointcept.models", f"{REF}/pointcept/models"),
                       ("pointcept.utils", f"{REF}/pointcept/utils")):
        m = types.ModuleType(name); m.__path__ = [path]; sys.modules[name] = m

    addict = types.ModuleType("addict")

    class Dict(dict):                      # addict.Dict semantics used by Point: attr access + recursive hook
        def __init__(self, *args, **kwargs):
            super().__init__()
            for a in args:
                if a is None:
                    continue
                for k, v in (a.items() if isinstance(a, dict) else a):
                    self[k] = self._hook(v)
            for k, v in kwargs.items():
                self[k] = self._hook(v)

        @classmethod
        def _hook(cls, item):
            if isinstance(item, dict):
                return cls(item)
            if isinstance(item, (list, tuple)):
                return type(item)(cls._hook(e) for e in item)
            return item

        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError:
                raise AttributeError(k)

        def __setattr__(self, k, v):
            self[k] = v
    addict.Dict = Dict
    sys.modules["addict"] = addict

    timm = types.ModuleType("timm"); tm = types.ModuleType("timm.models"); tl = types.ModuleType("timm.models.layers")

    class DropPath(nn.Module):
        def __init__(self, p=0.0):
            super().__init__(); self.p = p

        def forward(self, x):
            assert not self.training
            return x
    tl.DropPath = DropPath
    sys.modules.update({"timm": timm, "timm.models": tm, "timm.models.layers": tl})

    ts = types.ModuleType("torch_scatter")

    def segment_csr(src, indptr, reduce="sum"):
        counts = (indptr[1:] - indptr[:-1])
        seg = torch.repeat_interleave(torch.arange(len(counts)), counts)
        if reduce == "max":
            return O.segment_max(src, seg, len(counts))
        if reduce == "mean":
            return O.segment_mean(src, seg, counts)
        raise NotImplementedError(reduce)
    ts.segment_csr = segment_csr
    sys.modules["torch_scatter"] = ts

    sp = types.ModuleType("spconv"); spp = types.ModuleType("spconv.pytorch"); spm = types.ModuleType("spconv.pytorch.modules")

    class SparseConvTensor:
        def __init__(self, features, indices, spatial_shape, batch_size):
            self.features, self.indices, self.spatial_shape, self.batch_size = features, indices, spatial_shape, batch_size

        def replace_feature(self, f):
            return SparseConvTensor(f, self.indices, self.spatial_shape, self.batch_size)

    class SubMConv3d(nn.Module):
        def __init__(self, in_channels, out_channels, kernel_size, padding=0, bias=True, indice_key=None):
            super().__init__()
            k = kernel_size
            self.weight = nn.Parameter(torch.zeros(out_channels, k, k, k, in_channels))
            self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None

        def forward(self, x):
            idx = x.indices.long()
            return x.replace_feature(O.subm_conv3d(x.features, idx[:, 0], idx[:, 1:], self.weight, self.bias))
    spm.is_spconv_module = lambda m: isinstance(m, SubMConv3d)
    spp.SubMConv3d = SubMConv3d; spp.SparseConvTensor = SparseConvTensor; spp.modules = spm
    sp.pytorch = spp
    sys.modules.update({"spconv": sp, "spconv.pytorch": spp, "spconv.pytorch.modules": spm})

    ppt = types.ModuleType("pointcept.models.point_prompt_training")
    ppt.PDNorm = type("PDNorm", (nn.Module,), {})
    sys.modules["pointcept.models.point_prompt_training"] = ppt


def load_reference():
    install_shims()
    ptv3 = importlib.import_module("pointcept.models.point_transformer_v3.point_transformer_v3m1_base")
    comm = importlib.import_module("pointcept.utils.comm")
    ser = importlib.import_module("pointcept.models.utils.serialization.default")
    return ptv3, comm, ser


# ----------------------------------------------------------------------------
SMALL_CFG = dict(
    c_in_channels=6, n_in_channels=6, order=("z", "z-trans", "hilbert", "hilbert-trans"),
    c_stride=(4, 4), c_enc_depths=(2, 1, 2), c_enc_channels=(16, 32, 64), c_enc_num_head=(1, 2, 4),
    c_dec_depths=(1, 2), c_dec_channels=(32, 32), c_dec_num_head=(2, 2),
    n_stride=(2, 2, 2, 2), n_enc_depths=(2, 1, 2, 1, 2), n_enc_channels=(16, 32, 64, 96, 128),
    n_enc_num_head=(1, 2, 4, 6, 8), n_dec_depths=(1, 1, 2, 1), n_dec_channels=(32, 32, 64, 96),
    n_dec_num_head=(2, 2, 4, 6),
    mlp_ratio=4, qkv_bias=True, qk_scale=None, attn_drop=0.0, proj_drop=0.0, drop_path=0.3, pre_norm=True,
    shuffle_orders=False, enable_rpe=False, enable_flash=False, upcast_attention=False, upcast_softmax=False,
    num_classes=20, T_dim=128, tm_bidirectional=False, tm_feat=1.0, tm_restomer=False, condition=True,
    skip_connection_mode="cat", b_factor=[1.0] * 4, s_factor=[1.0] * 4,
    skip_connection_scale=True, skip_connection_scale_i=False,
)


def stage_sizes(scene, n_levels=4):
    from oracle import serialization_np as S
    g = scene["grid_coord"]; off = scene["offset"]
    b = S.offset2batch(off)
    code, _, _, depth = S.serialization(g, b)
    sizes = [np.diff(off, prepend=0)]
    for _ in range(n_levels):
        pl = S.pool_plan(code, 2, depth)
        bb = code[0][pl["head_indices"]] >> (3 * depth)
        code, depth = pl["code"], pl["depth"]
        sizes.append(np.bincount(bb, minlength=len(off)))
    return sizes


def patch_sizes_for(scene, cap):
    """largest power of two <= min_b N_s(b) (capped): the reference's dense branch
    (used here because flash_attn has no CPU path) then partitions exactly like its
    flash branch, which is what the CUDA kernel implements."""
    out = []
    for cnt in stage_sizes(scene):
        m = int(cnt.min())
        k = 1 << (m.bit_length() - 1)
        out.append(min(cap, k) if len(cnt) > 1 else cap)
    return out


def run_case(ptv3, comm, name, scene, cap, cfg_over=None):
    cfg = dict(SMALL_CFG)
    ps = patch_sizes_for(scene, cap)
    cfg.update(n_enc_patch_size=tuple(ps), n_dec_patch_size=tuple(ps[:4]),
               c_enc_patch_size=(ps[0], ps[2], ps[4]), c_dec_patch_size=(ps[0], ps[2]))
    if cfg_over:
        cfg.update(cfg_over)
    torch.manual_seed(0)
    model = ptv3.PointTransformerV3(**cfg).eval()
    shapes = {k: list(v.shape) for k, v in model.state_dict().items()}
    model.load_state_dict(synth_state_dict(shapes), strict=True)
    N = len(scene["coord"])
    rng = np.random.default_rng(1234)
    noise = rng.standard_normal((N, cfg["c_in_channels"])).astype(np.float32)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    base = dict(coord=t(scene["coord"]), grid_coord=t(scene["grid_coord"]).long(), offset=t(scene["offset"]))
    out = {}
    # every SerializedPooling draws torch.randperm(4) on the CPU generator (ptv3.py:501-505);
    # record the draws so the oracle / CUDA path can replay them
    perms = []
    real_randperm = torch.randperm
    def rec_randperm(n, *a, **k):
        p = real_randperm(n, *a, **k); perms.append(p.numpy().copy()); return p
    torch.manual_seed(20260925)
    torch.randperm = rec_randperm
    try:
      with torch.no_grad():
        if cfg["condition"]:
            ts = 999 * torch.ones((N, 1), dtype=torch.int64)
            c_point = dict(base, feat=t(noise), t_emb=comm.calc_t_emb(ts, cfg["T_dim"]))
            n_point = dict(base, feat=t(scene["feat"]))
            c_out, n_out = model(c_point, n_point)
            out["c_feat"] = c_out["feat"].numpy()
        else:
            n_out = model(n_point=dict(base, feat=t(scene["feat"])))
        out["n_feat"] = n_out["feat"].numpy()
        out["perms"] = np.stack(perms)
        out["seed"] = 20260925
        out["serialized_code"] = n_out["serialized_code"].numpy()
        out["serialized_order"] = n_out["serialized_order"].numpy()
        out["serialized_inverse"] = n_out["serialized_inverse"].numpy()
    finally:
        torch.randperm = real_randperm
    np.savez_compressed(os.path.join(HERE, f"ptv3_{name}.npz"), noise=noise, **{k: scene[k] for k in scene}, **out)
    jcfg = {k: (list(v) if isinstance(v, tuple) else v) for k, v in cfg.items()}
    with open(os.path.join(HERE, f"ptv3_{name}.json"), "w") as f:
        json.dump(dict(cfg=jcfg, shapes=shapes), f, indent=0)
    print(name, "N", N, "patch", ps, "n_feat", out["n_feat"].shape, float(np.abs(out["n_feat"]).mean()))


def main():
    ptv3, comm, ser = load_reference()

    # ---- codes ---------------------------------------------------------------
    g = torch.Generator().manual_seed(0)
    rec = {}
    for depth in (1, 2, 3, 5, 8, 9, 11, 16):
        n = 4096 if depth > 3 else 512
        gc = torch.randint(0, 2 ** depth, (n, 3), generator=g, dtype=torch.int32)
        gc[0] = 0; gc[1] = 2 ** depth - 1                     # extremes
        batch = torch.randint(0, 5, (n,), generator=g).sort().values
        rec[f"grid_{depth}"] = gc.numpy()
        rec[f"batch_{depth}"] = batch.numpy()
        for o in ("z", "z-trans", "hilbert", "hilbert-trans"):
            rec[f"code_{depth}_{o}"] = ser.encode(gc, batch, depth, order=o).numpy()
    gc = torch.randint(0, 512, (120000, 3), generator=torch.Generator().manual_seed(0), dtype=torch.int32)
    rec["kat_first3_z"] = ser.encode(gc, None, 9, "z")[:3].numpy()
    rec["kat_first3_hilbert"] = ser.encode(gc, None, 9, "hilbert")[:3].numpy()
    np.savez_compressed(os.path.join(HERE, "codes.npz"), **rec)

    # ---- Point.serialization (order / inverse) --------------------------------
    structure = importlib.import_module("pointcept.models.utils.structure")
    sc = synth.collate([synth.small_room(1500, 3), synth.small_room(700, 4)])
    p = structure.Point(coord=torch.from_numpy(sc["coord"]), grid_coord=torch.from_numpy(sc["grid_coord"]),
                        offset=torch.from_numpy(sc["offset"]), feat=torch.from_numpy(sc["feat"]))
    p.serialization(order=("z", "z-trans", "hilbert", "hilbert-trans"), shuffle_orders=False)
    np.savez_compressed(os.path.join(HERE, "serialization.npz"), grid_coord=sc["grid_coord"], offset=sc["offset"],
                        depth=p.serialized_depth, code=p.serialized_code.numpy(), order=p.serialized_order.numpy(),
                        inverse=p.serialized_inverse.numpy(), batch=p.batch.numpy())

    # ---- padding maps ----------------------------------------------------------
    attn = ptv3.SerializedAttention(channels=16, num_heads=1, patch_size=4, enable_flash=True,
                                    upcast_attention=False, upcast_softmax=False)
    rec = {}; cases = [([5, 12], 4), ([10], 4), ([3, 11], 4), ([4, 8], 4), ([1, 2, 40], 16), ([1024, 3000, 3001], 1024),
                       ([130], 128), ([128], 128), ([127, 300], 128)]
    for i, (off, K) in enumerate(cases):
        attn.patch_size = K
        pt = structure.Point(offset=torch.tensor(off, dtype=torch.int64))
        pad, unpad, cu = attn.get_padding_and_inverse(pt)
        rec[f"offset_{i}"] = np.array(off); rec[f"K_{i}"] = K
        rec[f"pad_{i}"] = pad.numpy(); rec[f"unpad_{i}"] = unpad.numpy(); rec[f"cu_{i}"] = cu.numpy()
    rec["n_cases"] = len(cases)
    np.savez_compressed(os.path.join(HERE, "padding.npz"), **rec)

    # ---- full dual network ------------------------------------------------------
    one = synth.collate([synth.scannet_scene(3000, 5, room_m=(3.0, 2.4, 1.6), n_boxes=3)])
    run_case(ptv3, comm, "case1_single", one, cap=64)
    two = synth.collate([synth.scannet_scene(2600, 6, room_m=(3.0, 2.4, 1.6), n_boxes=3),
                         synth.scannet_scene(1900, 7, room_m=(2.4, 2.4, 1.6), n_boxes=2)])
    run_case(ptv3, comm, "case2_batch2", two, cap=64, cfg_over=dict(shuffle_orders=True))
    # BASELINE config 1: 2k-point room, CN only (condition=False), patch 128
    room = synth.collate([synth.small_room(2000, 0)])
    run_case(ptv3, comm, "case3_cn_only", room, cap=128, cfg_over=dict(condition=False))
    # BASELINE config 4 shape: nuScenes-like sweeps, 4 input channels (coord + strength), 16 classes, grid 0.05 m => depth 11
    if "case4" in sys.argv[1:] or len(sys.argv) == 1:
        sweeps = synth.collate([synth.nuscenes_sweep(3000, 0), synth.nuscenes_sweep(2400, 1)])
        run_case(ptv3, comm, "case4_nuscenes", sweeps, cap=64, cfg_over=dict(c_in_channels=4, n_in_channels=4, num_classes=16))


if __name__ == "__main__":
    main()
