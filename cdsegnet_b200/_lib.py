"""Build + ctypes loader for libcdseg_b200.so (the C-ABI CUDA library, include/cdseg_b200.h).

The library is built IN-TREE (cdsegnet_b200/libcdseg_b200.so) with nvcc for sm_100a only.
There is deliberately no CPU fallback: if the library is missing or a call fails the
product path raises.
"""
import ctypes
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.environ.get("CDSEG_LIB") or os.path.join(_HERE, "libcdseg_b200.so")      # CDSEG_LIB: an experimental build (profiles/)
SOURCES = ["serialize.cu", "pool.cu", "conv.cu", "pointwise.cu", "attn_pack.cu", "attn_tc.cu", "attn_tc2.cu", "attn_tc3.cu", "gemm_tc.cu", "fused_post.cu", "fused_pre.cu", "block_exec.cu", "plan_exec.cu", "net_exec.cu", "losses.cu", "fragments.cu", "knn.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-cudart", "shared"]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ for sm_100a into one shared library."""
    if not force and not needs_build():
        return LIB_PATH
    objs = []
    bdir = os.path.join(_HERE, "build")
    os.makedirs(bdir, exist_ok=True)
    procs = []
    for s in SOURCES:
        o = os.path.join(bdir, s.replace(".cu", ".o"))
        cmd = [_nvcc()] + NVCC_FLAGS + ["-c", os.path.join(CSRC, s), "-o", o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {s}:\n{out}")
        if verbose and out.strip():
            print(out)
    cmd = [_nvcc(), "-shared", "-cudart", "shared", "-o", LIB_PATH] + objs
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    return LIB_PATH


_P = ctypes.c_void_p
_I = ctypes.c_int
_L = ctypes.c_int64
_F = ctypes.c_float
_Z = ctypes.c_size_t

class BlockArgs(ctypes.Structure):
    """mirror of CdsegBlockArgs (include/cdseg_b200.h)"""
    _fields_ = [("n", ctypes.c_int64), ("C", _I), ("H", _I), ("T_dim", _I), ("B", _I),
                ("x", _P), ("conv_in", _P), ("nbr", _P), ("tile_mask", _P), ("batch", _P), ("conv_plan", _P), ("t_scene", _P),
                ("slot_src", _P), ("slot_dst", _P), ("patch_len", _P), ("T", _I), ("Kp", _I), ("scale", _F),
                ("conv_Bp", _P), ("conv_b", _P), ("lin_Bp", _P), ("lin_b", _P), ("cpe_g", _P), ("cpe_b", _P),
                ("t_W", _P), ("t_b", _P), ("n1_g", _P), ("n1_b", _P), ("qkv_Bp", _P), ("qkv_b", _P), ("proj_Bp", _P),
                ("proj_b", _P), ("n2_g", _P), ("n2_b", _P), ("fc1_Bp", _P), ("fc1_b", _P), ("fc2_Bp", _P), ("fc2_b", _P),
                ("ln_eps", _F), ("attn_mode", _I), ("out", _P), ("scratch", _P), ("scratch_bytes", _Z), ("ev", _P * 6)]


MAX_SCENES = 64


class PatchMap(ctypes.Structure):
    """mirror of CdsegPatchMap"""
    _fields_ = [("slot_src", _P), ("slot_dst", _P), ("point_slot", _P), ("patch_len", _P), ("T", _I), ("Kp", _I), ("K", _I), ("pad_", _I),
                ("pairs", _L)]


class PlanLevel(ctypes.Structure):
    """mirror of CdsegPlanLevel"""
    _fields_ = [("parent", _I), ("stride", _I), ("rowmap", _I * 4), ("K", _I), ("pm_mask", ctypes.c_uint), ("want_conv_plan", _I),
                ("stem_ksize", _I), ("n", _L), ("cap", _L), ("B", _I), ("depth", _I), ("c0", _I), ("pooling_depth", _I), ("slot", _I),
                ("pad_", _I), ("offset_host", _L * MAX_SCENES),
                ("grid", _P), ("batch", _P), ("code", _P), ("order", _P), ("inverse", _P),
                ("cluster", _P), ("idx_ptr", _P), ("head", _P), ("m_dev", _P), ("offset_dev", _P),
                ("perm", _P), ("inv_perm", _P), ("o_code", _P), ("o_order", _P), ("o_inverse", _P),
                ("nbr3", _P), ("tile_mask3", _P), ("conv_plan3", _P), ("nbr_stem", _P), ("pm", PatchMap * 4),
                ("ready", _P), ("ready_stem", _P)]


# name -> (restype, argtypes); must list every symbol declared in include/cdseg_b200.h
SIGNATURES = {
    "cdseg_abi_version": (_I, []),
    "cdseg_set_pdl": (None, [_I]),
    "cdseg_get_pdl": (_I, []),
    "cdseg_launch_count": (ctypes.c_ulonglong, []),
    "cdseg_launch_count_reset": (None, []),
    "cdseg_grid_max": (_I, [_P, _L, _P, _P]),
    "cdseg_offset2batch": (_I, [_P, _I, _L, _P, _P]),
    "cdseg_encode_codes": (_I, [_P, _P, _L, _I, ctypes.POINTER(_I), _I, _P, _P]),
    "cdseg_debug_encode_host": (_I, [_P, _P, _L, _I, _I, _P]),
    "cdseg_argsort_workspace_bytes": (_Z, [_I, _L]),
    "cdseg_argsort_rows": (_I, [_P, _I, _L, _I, _P, _P, _P, _Z, _P]),
    "cdseg_gather_rows": (_I, [_P, _P, _L, _I, _P, _P]),
    "cdseg_renumber": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _L, _P, _P, _P, _P, _P, _P]),
    "cdseg_patch_count": (_I, [ctypes.POINTER(_L), _I, _I, ctypes.POINTER(_I)]),
    "cdseg_patch_maps": (_I, [_P, ctypes.POINTER(_L), _I, _I, _I, _P, _P, _P, _P, _P]),
    "cdseg_pool_plan_workspace_bytes": (_Z, [_I, _L]),
    "cdseg_pool_plan": (_I, [_P, _P, _I, _L, _P, _L, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _L, _P, _P, _P, _P, _P, _Z, _P]),
    "cdseg_pool_reduce": (_I, [_P, _P, _P, _P, _L, _I, _P, _P, _I, _P, _P, _P]),
    "cdseg_unpool_add": (_I, [_P, _P, _P, _L, _I, _F, _P, _P]),
    "cdseg_nbr_workspace_bytes": (_Z, [_L]),
    "cdseg_hash_capacity": (_L, [_L]),
    "cdseg_nbr_build": (_I, [_P, _P, _L, _I, _P, _P, _Z, _P]),
    "cdseg_subm_conv": (_I, [_P, _P, _P, _P, _P, _P, _I, _L, _I, _I, _I, _P, _P]),
    "cdseg_attn_pack_f16": (_I, [_P, _L, _I, _I, _I, _P, _I, _I, _I, _P, _P, _P, _P]),
    "cdseg_attn_pack_f32": (_I, [_P, _L, _I, _I, _I, _P, _I, _I, _I, _P, _P, _P, _P]),
    "cdseg_attn_pack_f16v": (_I, [_P, _L, _I, _I, _I, _P, _I, _I, _I, _P, _P, _P, _I, _P]),
    "cdseg_attn_set_poly": (None, [_I]),
    "cdseg_attn_tc2": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _F, _P, _L, _P]),
    "cdseg_attn_tc_smem_bytes": (_Z, [_I]),
    "cdseg_attn_tc": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _F, _P, _L, _P]),
    "cdseg_attn_pack_split": (_I, [_P, _L, _I, _I, _I, _P, _I, _I, _I, _P, _P, _P, _I, _P]),
    "cdseg_attn_set_debug": (None, [_I]),
    "cdseg_attn_tc3": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _F, _I, _P, _L, _P]),
    "cdseg_attn_exact": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _F, _P, _L, _P]),
    "cdseg_add_layernorm": (_I, [_P, _P, _P, _P, _P, _P, _F, _L, _I, _P, _P, _P]),
    "cdseg_reduce_ln": (_I, [_P, _I, _P, _P, _P, _P, _P, _P, _P, _P, _F, _L, _I, _P, _P, _P]),
    "cdseg_scale_shift_act": (_I, [_P, _P, _P, _I, _L, _I, _P, _P]),
    "cdseg_small_linear": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _P]),
    "cdseg_rows_uniform": (_I, [_P, _P, _P, _L, _I, _P, _P]),
    "cdseg_set_gemm_precision": (None, [_I]),
    "cdseg_get_gemm_precision": (_I, []),
    "cdseg_gemm_packed_b_floats": (_Z, [_I, _I, _I]),
    "cdseg_gemm_pack_b": (_I, [_P, _I, _I, _I, _P, _P]),
    "cdseg_tile_tap_mask": (_I, [_P, _L, _I, _P, _P]),
    "cdseg_gemm_tc_workspace_bytes": (_Z, [_L, _I, _I]),
    "cdseg_gemm_tc_set_trace": (None, [_P, _I]),
    "cdseg_gemm_tc_set_narrow": (None, [_I]),
    "cdseg_post_attn": (_I, [_P, _P, _L, _I, _P, _P, _P, _P, _F, _P, _P, _P, _P, _P, _P]),
    "cdseg_pre_attn": (_I, [_P, _P, _L, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _F, _P, _P, _P, _P, _P]),
    "cdseg_pre_attn_set_trace": (None, [_P, _I]),
    "cdseg_conv_plan_bytes": (_Z, [_L]),
    "cdseg_conv_tile_plan": (_I, [_P, _L, _P, _P]),
    "cdseg_plan_arena_bytes": (_Z, [_L, _I, _I, _I, _I, _I, _I, _I]),
    "cdseg_plan_finish": (_I, []),
    "cdseg_debug_pick_split": (_I, [_L, _I]),
    "cdseg_debug_pick_split_block": (_I, [_L, _I]),
    "cdseg_plan_build": (_I, [_P, _P, _L, _I, ctypes.POINTER(_I), _I, ctypes.POINTER(PlanLevel), _I, _P, _I, ctypes.POINTER(ctypes.c_int32), _P, _Z, _P, _P]),
    "cdseg_net_arena_bytes": (_I, [_P, ctypes.POINTER(_Z), ctypes.POINTER(_Z)]),
    "cdseg_net_forward": (_I, [_P]),
    "cdseg_net_set_debug": (None, [_I]),
    "cdseg_struct_sizes": (_I, [ctypes.POINTER(_Z), _I]),
    "cdseg_gather_rows_pad": (_I, [_P, _P, _L, _I, _I, _P, _P]),
    "cdseg_set_fused_mask": (None, [_I]),
    "cdseg_set_fused_ctas_per_sm": (None, [_I]),
    "cdseg_block_scratch_bytes": (_Z, [_L, _I, _I, _I, _I, _I]),
    "cdseg_block_forward": (_I, [ctypes.POINTER(BlockArgs), _P]),
    "cdseg_event_create": (_P, []),
    "cdseg_event_destroy": (None, [_P]),
    "cdseg_event_elapsed_ms": (_I, [_P, _P, ctypes.POINTER(_F)]),
    "cdseg_conv_im2col_tc": (_I, [_P, _P, _I, _P, _L, _I, _P, _I, _P, _L, _P]),
    "cdseg_criteria_workspace_bytes": (_Z, [_L, _I]),
    "cdseg_criteria": (_I, [_P, _P, _L, _I, _L, _P, _P, _I, _I, _F, _F, _F, _I, _I, _I, _P, _P, _Z, _P]),
    "cdseg_criteria_grad": (_I, [_P, _P, _L, _I, _L, _P, _P, _I, _I, _F, _F, _F, _I, _I, _I, _I, _P, _P, _P, _P, _Z, _P]),
    "cdseg_adamw_step": (_I, [_P, _I, _F, _F, _F, _F, _F, _I, _P]),
    "cdseg_q_sample": (_I, [_P, _P, _P, _P, _P, _L, _I, _P, _P]),
    "cdseg_ddim_step": (_I, [_P, _P, _L, _F, _F, _F, _F, _I, _I, _P, _P]),
    "cdseg_axpy_scale": (_I, [_P, _P, _F, _F, _L, _P]),
    "cdseg_grid_sample_workspace_bytes": (_Z, [_L]),
    "cdseg_grid_sample_plan": (_I, [_P, _I, _L, ctypes.c_double, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _Z, _P]),
    "cdseg_fragment_index": (_I, [_P, _P, _I, _I, _P, _P]),
    "cdseg_vote_softmax_add": (_I, [_P, _P, _L, _I, _P, _P]),
    "cdseg_argmax_rows": (_I, [_P, _L, _I, _P, _P]),
    "cdseg_knn_workspace_bytes": (_Z, [_L]),
    "cdseg_knn_query": (_I, [_I, _I, _P, _P, _P, _P, _I, _L, _P, _P, _P, _Z, _P]),
    "cdseg_gemm_tc": (_I, [_P, _L, _P, _I, _P, _P, _L, _I, _I, _P, _P, _L, _I, _P, _L, _I, _P, _Z, _P]),
}

_lib = None


def load():
    """dlopen the in-tree library and bind every entry point.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback for the cdsegnet_b200 hot path)")
    import torch  # noqa: F401  (loads libcudart.so.12 first so both sides share one CUDA runtime)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class CdsegError(RuntimeError):
    pass


def check(status, what):
    if status != 0:
        kind = {-1: "invalid argument", -2: "workspace too small"}.get(status, f"CUDA error {status}")
        raise CdsegError(f"{what}: {kind}")
