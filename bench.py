#!/usr/bin/env python
"""bench.py -- points/sec of the CDSegNet single-step forward (BASELINE.json metric) on B200.

  python bench.py --gpus N --steps K --warmup W            # the CUDA path (this repo)
  python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host cores (CPU oracle port)

A "step" = one full DefaultSegmentorV2.inference(eval=False) of one synthetic ScanNet-shaped scene of
120 000 unique voxels (BASELINE.json configs[1]: "full CDSegNet (CN+NN) fp32 forward") through the full CDSegNet
(CN + NN + TransferModule, 101.4 M parameters): serialization (key encode + 4 radix argsorts), pooling hierarchy,
37 blocks, heads.  The headline runs the fp32-faithful path (dense layers on the 3-term fp16 split, attention in the
"tc32" tcgen05 mode) and its logits are checked IN THIS RUN against the CPU oracle on the same scene, weights, noise and
shuffles (`parity`); the fp16 flash-branch attention mode is timed beside it (`attention_f16`).
Scenes shard one per GPU (no collective on the inference forward) => weak scaling.
Other workloads (--workload nuscenes | scannet200 | batch8) time the same forward on the other BASELINE.json shapes.
Prints ONE JSON line on rank 0 (contract in the task statement; fields documented in DESIGN.md §Measurement).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_POINTS = 120000
METRIC = "points/sec (120k-point ScanNet-shaped scene, full CDSegNet single-step forward)"


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons DURING the timed region (B200_PROFILING.md recipe).  Reads NVML in-process (nvidia_ml_py:
    microseconds per sample, no child process competing with the enqueue thread for a core -- spawning nvidia-smi every
    200 ms from every rank cost the 2-GPU run measurable step time); falls back to the recipe's nvidia-smi query line."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.nvml, self.source = index, [], False, None, "nvidia-smi"
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml, self.source = pynvml, "nvml"
        except Exception:
            self.nvml = None

    def sample_nvml(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
        r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        bits = [n.nvmlClocksThrottleReasonHwSlowdown, n.nvmlClocksThrottleReasonHwThermalSlowdown,
                n.nvmlClocksThrottleReasonSwThermalSlowdown, n.nvmlClocksThrottleReasonSwPowerCap]
        return [str(sm), str(mx), "0"] + ["Active" if r & b else "Not Active" for b in bits]

    def run(self):
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    self.rows.append(self.sample_nvml())
                else:
                    out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                         capture_output=True, text=True, timeout=5).stdout.strip()
                    if out:
                        self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.02 if self.nvml is not None else 0.2)

    def summary(self):
        self.stop_flag = True
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        reasons = set()
        for r in self.rows:
            for nme, v in zip(self.NAMES, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        mx = max((float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()), default=None)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(self.rows), "source": self.source}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops"], d.get("bf16_tflops_sustained", d["bf16_tflops"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1590.0, 1400.0, "fallback (B200_PROFILING.md)"


WORKLOADS = {
    # name: (description, in_channels, num_classes)
    "scannet": ("ScanNet-shaped scene 120k unique voxels @0.02 m (BASELINE.json configs[1])", 6, 20),
    "scannet200": ("ScanNet200: the same 120k-voxel scene with the 200-class head (BASELINE.json configs[4] shape)", 6, 200),
    "nuscenes": ("nuScenes-shaped batch of 8 outdoor sweeps, ~240k voxels @0.05 m, depth-11 keys, 4 input channels, 16 classes "
                 "(BASELINE.json configs[3] shape, per-GPU batch)", 4, 16),
    "batch8": ("batch of 8 ScanNet-shaped scenes of 80k-102.4k voxels (BASELINE.json configs[2] shape, inference forward)", 6, 20),
}


def make_scene(seed=0, n=N_POINTS, workload="scannet"):
    from cdsegnet_b200 import synth
    if workload == "nuscenes":
        return synth.collate([synth.nuscenes_sweep(30000, seed + i) for i in range(8)])
    if workload == "batch8":
        sizes = [80000, 102400, 96000, 88000, 102400, 91000, 84000, 99000]
        return synth.collate([synth.scannet_scene(s, seed + i) for i, s in enumerate(sizes)])
    if n >= 60000:
        sc = synth.scannet_scene(n, seed)
    else:   # bounded CPU sample: same generator, smaller room so that surface density stays ScanNet-like
        f = (n / N_POINTS) ** 0.5
        sc = synth.scannet_scene(n, seed, room_m=(max(2.0, 8.0 * f), max(1.6, 6.0 * f), 3.0), n_boxes=max(2, int(12 * f)))
    return synth.collate([sc])


def random_weights(model, seed=0):
    """reference default init under a fixed seed + randomised BN running stats."""
    import torch
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            m.running_mean.copy_(0.1 * torch.randn(m.running_mean.shape, generator=g))
            m.running_var.copy_(0.5 + torch.rand(m.running_var.shape, generator=g))


def fixed_draws(n, in_ch, seed=5):
    """the random draws of one forward, fixed so that the GPU run and the oracle see the same ones: the Noise-Network input
    (default.py:393) and the eight curve-order shuffles (structure.py:95, ptv3.py:502)"""
    import numpy as np
    rng = np.random.default_rng(seed)
    return rng.standard_normal((n, in_ch)).astype(np.float32), [rng.permutation(4) for _ in range(8)]


def replay(perms):
    perms = [p.copy() for p in perms]
    return lambda k: perms.pop(0)


# --------------------------------------------------------------------------------------------
def cpu_forward_timed(n_points, seed=0, repeats=1, workload="scannet", sd=None, draws=None):
    """the oracle (CPU port of the reference forward, dense fp32 attention) on the host cores -> (points/s, cores, seconds, logits)."""
    import numpy as np
    import torch
    from oracle import ptv3_oracle as O
    import cdsegnet_b200 as cb
    from cdsegnet_b200 import configs
    # the port is a chain of small torch ops: beyond ~16 threads intra-op parallelism only adds contention
    # (measured: 128 threads on the GPU box's host were 10x slower than 16), so use min(cores, 16)
    cores = min(os.cpu_count(), 16)
    torch.set_num_threads(cores)
    _, in_ch, classes = WORKLOADS[workload]
    cfg = configs.backbone_cfg(in_channels=in_ch, num_classes=classes)
    if sd is None:
        torch.manual_seed(0)
        model = cb.PointTransformerV3(**cfg)
        random_weights(model)
        sd = {k: v.detach() for k, v in model.state_dict().items()}
    sc = make_scene(seed, n_points, workload)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    base = dict(coord=t(sc["coord"]), grid_coord=t(sc["grid_coord"]).long(), offset=t(sc["offset"]))
    n = len(sc["coord"])
    best = logits = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        if draws is None:
            noise, pf = torch.normal(0, 1, size=(n, in_ch)), None
        else:
            noise, pf = t(draws[0]), replay(draws[1])
        ts = 999 * torch.ones((n, 1), dtype=torch.int64)
        _, n_out = O.forward(sd, cfg, dict(base, feat=noise, t_emb=O.calc_t_emb(ts, 128)), dict(base, feat=t(sc["feat"])),
                             attn_mode="dense", perm_fn=pf)
        dt = time.perf_counter() - t0
        logits = n_out["feat"].numpy()
        best = dt if best is None else min(best, dt)
    return n / best, cores, best, logits


def run_reference(args):
    """Reference arm: the reference's algorithm (CPU oracle port, dense fp32 attention) on the host cores, on the SAME workload as the
    CUDA arm (the full 120k-point scene).  One forward takes 10-25 s, so at most REF_MAX_STEPS forwards are timed after one warm-up
    whatever --steps / --warmup ask for (the contract's "bounded sample"); the line says so."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    REF_MAX_STEPS = 3
    steps, warm = min(args.steps, REF_MAX_STEPS), min(args.warmup, 1)
    n = args.cpu_points or N_POINTS               # (ignored by the fixed-size nuscenes / batch8 workloads)
    for _ in range(warm):
        cpu_forward_timed(n, workload=args.workload)
    vals, t0 = [], time.perf_counter()
    for _ in range(steps):
        v, cores, dt, lg = cpu_forward_timed(n, workload=args.workload)
        vals.append(dt)
    npts = lg.shape[0]
    total = time.perf_counter() - t0
    value = npts * steps / sum(vals)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "points/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "steps_run": steps, "warmup_run": warm, "ms_per_step": 1e3 * sum(vals) / steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args.workload, npts, 1),
            "cpu_baseline": {"value": value, "unit": "points/s", "cores": cores, "kind": "port",
                             "sample": f"{steps} forward(s) of the full {npts}-point workload after {warm} warm-up "
                                       f"(a forward takes ~{sum(vals) / steps:.0f} s: the requested {args.steps} steps are capped at {REF_MAX_STEPS})"},
            "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": total}
    print(json.dumps(line))


def config_dict(workload, n, world):
    return {"workload": WORKLOADS[workload][0] + ", full CDSegNet (CN+NN+TransferModule, 101.4M params), single-step inference forward "
                                                "(DefaultSegmentorV2.inference), patch 1024, fp32-faithful path",
            "workload_key": workload, "points_per_step_per_gpu": int(n), "l2": "flushed (256 MiB write) between timed iterations",
            "parallelism": f"scene-per-GPU x{world} (same synthetic input on every rank), no collective"}


# --------------------------------------------------------------------------------------------
def run_cuda(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import cdsegnet_b200 as cb
    from cdsegnet_b200 import configs, ops

    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    desc, in_ch, classes = WORKLOADS[args.workload]
    torch.manual_seed(0)
    seg = cb.build_model(configs.segmentor_cfg(in_channels=in_ch, num_classes=classes))
    random_weights(seg)
    seg = seg.to(dev).eval()
    # BASELINE.json configs[1] is the fp32 forward: attention in the fp32-faithful tensor-core mode.  The shipped config's
    # enable_flash=True (fp16 flash-branch numerics) is timed separately below.
    seg.backbone.attention_mode = args.attention
    ops.set_gemm_precision(args.gemm)

    sc = make_scene(seed=0, workload=args.workload)   # the SAME input on every rank: weak scaling keeps the per-GPU work identical
    n = len(sc["coord"])
    host = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in sc.items()}
    host["grid_coord"] = host["grid_coord"].int().pin_memory()
    resident = {k: v.to(dev) for k, v in host.items()}
    noise_dev = torch.randn(n, in_ch, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2

    def step_resident():
        return seg.inference(resident, eval=False, noise=noise_dev)["seg_logits"]

    logits_host = torch.empty((n, classes), dtype=torch.float32).pin_memory()

    def step_e2e():
        inp = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
        out = seg.inference(inp, eval=False)["seg_logits"]              # draws the NN noise like default.py:393
        logits_host.copy_(out, non_blocking=True)
        return out

    # ---- parity forward (untimed): fixed noise + shuffles, compared with the oracle's logits further down -----------------
    draws = fixed_draws(n, in_ch)
    seg.backbone.perm_fn = replay(draws[1])
    parity_logits = seg.inference(resident, eval=False, noise=torch.from_numpy(draws[0]).to(dev))["seg_logits"].cpu().numpy()
    seg.backbone.perm_fn = None

    rank_ms = {}
    step_stats = {}

    def timed(fn, steps, warmup, profile=False):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        evs = []
        ops.launch_count_reset()
        if profile:
            ops.PROFILE = []
        t0 = time.perf_counter()
        for _ in range(steps):
            flush.zero_()                                                # L2 flush between timed iterations (untimed)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        if world > 1:
            dist.barrier()
        launches = ops.launch_count()
        per_step = [a.elapsed_time(b) for a, b in evs]
        ms = sum(per_step)
        step_stats[fn.__name__ + ("" if fn.__name__ not in step_stats else "'")] = per_step
        prof = ops.PROFILE
        ops.PROFILE = None
        tsum = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            allr = [torch.zeros_like(tsum) for _ in range(world)]
            dist.all_gather(allr, tsum)
            rank_ms.setdefault(fn.__name__, [float(a.item()) / steps for a in allr])     # first call = the headline timed run
            dist.all_reduce(tsum, op=dist.ReduceOp.MAX)
        return float(tsum.item()), launches, wall, prof

    # headline: no per-kernel events inside the timed region (the profiled pass further down is separate and shorter)
    sampler = ClockSampler(local)
    sampler.start()
    ms_total, launches, wall, _ = timed(step_resident, args.steps, args.warmup)
    clocks = sampler.summary()
    ms_e2e, _, _, _ = timed(step_e2e, args.steps, max(3, args.warmup // 2))

    value = world * n * args.steps / (ms_total / 1e3)
    e2e = world * n * args.steps / (ms_e2e / 1e3)
    hbm, tf_burst, tf_sust, which = peaks()

    # Per-kernel numbers: a separate short pass with CUDA events recorded by cdseg_block_forward around its own launches, on the
    # launching stream, live in this process (NOT part of the headline's timed region).
    #  * fz::pre_kernel at stage 0 (fused cpe conv + Linear + LayerNorm + residual + norm1 + qkv): HBM roofline; algorithmic bytes =
    #    read the block input once (n*C*4; conv operand and residual are the same tensor) + the neighbour table (n*27*4) +
    #    write x1 (n*C*4) and qkv (n*3C*4) + the packed weights once.  Tensor-pipe view next to it (FLOPs = 2 * (conv pairs + 4n) * C^2;
    #    the x3 MMAs of the fp16 hi/lo split are NOT counted).
    #  * the one dense contraction north_star names = the attention kernel at stage 0 (tensor / MUFU bound): 4*pairs*C FLOP.
    #  * `traffic` = dram__bytes_read.sum + dram__bytes_write.sum per launch from this round's ncu --set full captures, read from
    #    profiles/r02_ncu_traffic.json (written by profiles/ncu_traffic.py from the .ncu-rep files; names the capture).
    traffic = {}
    tp = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp))

    def kernel_lines(prof, steps, conv_pairs):
        lib = ops._lib.load()
        import ctypes

        def ms(e0, e1):
            v = ctypes.c_float(0)
            lib.cdseg_event_elapsed_ms(e0, e1, ctypes.byref(v))
            return v.value
        nmax = max(p["n"] for p in prof)
        sel = [p for p in prof if p["n"] == nmax and p["C"] == 32]
        t_attn = sum(ms(p["ev"][0], p["ev"][1]) for p in sel) / len(sel)
        t_post = sum(ms(p["ev"][2], p["ev"][3]) for p in sel) / len(sel)
        t_pre = sum(ms(p["ev"][4], p["ev"][5]) for p in sel) / len(sel)
        p0 = sel[0]
        n0, C0 = p0["n"], p0["C"]
        fl = 4.0 * p0["pairs"] * C0
        ex = p0["pairs"] * p0["H"]
        mode = seg.backbone.attention_mode
        attn = {"bound": "tensor", "kernel": "tc3::attn_tc3_kernel<%s> (stage 0, %d launches/step)" % ({"tc32": "1,0: hi/lo-split operands, 11 MMAs/chunk",
                                                                                                       "f16": "0,0: fp16 operands, 5 MMAs/chunk"}.get(mode, mode),
                                                                                                      len(sel) // steps),
                "achieved": fl / (t_attn * 1e-3) / 1e12, "peak": tf_sust, "unit": "TFLOP/s", "frac": fl / (t_attn * 1e-3) / 1e12 / tf_sust,
                "flops_per_launch": fl, "ms_per_launch": t_attn, "exp_per_launch": ex, "gexp_per_s": ex / (t_attn * 1e-3) / 1e9,
                "mufu_peak_gexp_per_s": 148 * 16 * 1.965, "traffic": traffic.get("attn_" + mode), "traffic_source": traffic.get("source")}
        wbytes = (27 + 1 + 3) * C0 * C0 * 4
        by = n0 * C0 * 4 + n0 * 27 * 4 + n0 * C0 * 4 + n0 * 3 * C0 * 4 + wbytes
        pre_fl = 2.0 * (conv_pairs + 4 * n0) * C0 * C0
        roof = {"bound": "hbm", "kernel": "fz::pre_kernel (stage 0: cpe conv + Linear + LN + residual + norm1 + qkv fused, n=%d C=%d, %d launches/step)"
                % (n0, C0, len(sel) // steps),
                "achieved": by / (t_pre * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s", "frac": by / (t_pre * 1e-3) / 1e9 / hbm,
                "peak_source": which, "bytes_per_launch": by, "ms_per_launch": t_pre, "traffic": traffic.get("pre"),
                "traffic_source": traffic.get("source"),
                "tensor_view": {"flops_per_launch": pre_fl, "achieved_tflops": pre_fl / (t_pre * 1e-3) / 1e12,
                                "frac_of_bf16_peak": pre_fl / (t_pre * 1e-3) / 1e12 / tf_sust}}
        pby = 3 * n0 * C0 * 4 + 9 * C0 * C0 * 4
        post = {"bound": "hbm", "kernel": "fz::post_kernel (stage 0: proj + residual + norm2 + fc1 + GELU + fc2 + residual fused)",
                "achieved": pby / (t_post * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s", "frac": pby / (t_post * 1e-3) / 1e9 / hbm,
                "bytes_per_launch": pby, "ms_per_launch": t_post, "traffic": traffic.get("post")}
        for p in prof:
            for e in p["ev"]:
                lib.cdseg_event_destroy(e)
        return roof, attn, post

    roof = attn = post = None
    kp = max(2, min(5, args.steps))
    _, _, _, prof = timed(step_resident, kp, 1, profile=True)
    if prof:
        nb3 = seg.backbone.last_plan.n_levels[0].nbr(3)
        conv_pairs = int((nb3 >= 0).sum().item())                      # (active output, active input) pairs of the k=3 conv at level 0
        roof, attn, post = kernel_lines(prof, kp, conv_pairs)
        # The Noise Network runs on a second stream beside the Conditional Network, so the events around one launch also cover
        # whatever the other stream had resident.  A few extra steps with the two-stream schedule switched off give the same
        # launches alone on the device (reported next to the in-step time).
        seg.backbone.overlap_streams = False
        _, _, _, prof1 = timed(step_resident, kp, 1, profile=True)
        seg.backbone.overlap_streams = True
        r1, a1, p1 = kernel_lines(prof1, kp, conv_pairs)
        for full, alone in ((roof, r1), (attn, a1), (post, p1)):
            full["single_stream"] = {"ms_per_launch": alone["ms_per_launch"], "achieved": alone["achieved"], "frac": alone["frac"]}

    # the other attention numerics on the same workload (f16 = the shipped config's enable_flash=True: flash-branch fp16 numerics)
    other = "f16" if args.attention != "f16" else "tc32"
    seg.backbone.attention_mode = other
    ms_other, _, _, _ = timed(step_resident, max(3, args.steps // 2), 3)
    seg.backbone.attention_mode = args.attention
    alt = {"attention_mode": other, "ms_per_step": ms_other / max(3, args.steps // 2),
           "value": world * n * max(3, args.steps // 2) / (ms_other / 1e3), "unit": "points/s"}

    line = {"metric": METRIC, "value": value, "unit": "points/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": ("f16 operands / f32 accumulate in every dense layer (one MMA per term, --gemm fp16); " if args.gemm == "fp16" else "") + {"tc32": "f32 (every dense layer, conv and the attention contraction run on tcgen05 with fp16 hi/lo-split operands and "
                              "fp32 accumulation: fp32-class results; norms / pooling / epilogues in fp32)",
                      "exact": "f32 (dense layers / conv: tcgen05 fp16 hi/lo split; attention: SIMT fp32)",
                      "f16": "f32 (dense layers, conv, norms) + f16 tensor-core attention with f32 accumulate (reference flash branch)"}[args.attention],
            "data": "synthetic", "config": dict(config_dict(args.workload, n, world), attention_mode=args.attention, gemm_precision=args.gemm),
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": e2e, "unit": "points/s",
                    "h2d_bytes_per_step": int(sum(v.numel() * v.element_size() for v in host.values()) + n * in_ch * 4),
                    "d2h_bytes_per_step": int(n * classes * 4), "ms_per_step": ms_e2e / args.steps},
            "ms_per_step_median": float(np.median(step_stats["step_resident"])), "ms_per_step_min": float(np.min(step_stats["step_resident"])),
            "ms_per_step_max": float(np.max(step_stats["step_resident"])),
            "roofline": roof, "roofline_attention": attn, "roofline_post": post, "attention_" + other: alt, "parity": None}
    if world > 1:
        line["ms_per_step_by_rank"] = rank_ms.get("step_resident")
    if rank == 0:
        if world == 1 and not args.no_cpu:
            sd = {k[len("backbone."):]: v.detach().cpu() for k, v in seg.state_dict().items() if k.startswith("backbone.")}
            v, cores, dt, ref = cpu_forward_timed(args.cpu_points or N_POINTS, workload=args.workload, sd=sd,
                                                  draws=None if args.cpu_points else draws)
            line["cpu_baseline"] = {"value": v, "unit": "points/s", "cores": cores, "kind": "port",
                                    "sample": f"1 forward of the {ref.shape[0]}-point workload, same model, oracle port with dense fp32 attention ({dt:.1f} s)"}
            if not args.cpu_points:
                err = np.abs(parity_logits - ref)
                line["parity"] = {"max_abs": float(err.max()), "mean_abs": float(err.mean()), "tolerance": 1e-3 if (args.attention != "f16" and args.gemm == "fp32") else None,
                                  "logit_abs_max": float(np.abs(ref).max()), "argmax_agreement": float((parity_logits.argmax(1) == ref.argmax(1)).mean()),
                                  "mode": f"attention {args.attention}, native block executor, same scene / weights / Noise-Network input / curve shuffles as the "
                                          "CPU oracle (dense fp32 attention, ptv3.py:264-280)",
                                  "ok": bool(err.max() < (1e-3 if (args.attention != "f16" and args.gemm == "fp32") else 5e-2))}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default="scannet", choices=sorted(WORKLOADS))
    ap.add_argument("--attention", default="tc32", choices=["tc32", "f16", "exact"],
                    help="attention numerics of the headline run (tc32 = fp32-faithful tensor-core mode; f16 = flash-branch numerics)")
    ap.add_argument("--gemm", default="fp32", choices=["fp32", "fp16"],
                    help="dense-layer precision: fp32 = 3-term fp16 hi/lo split (fp32-class results); fp16 = one MMA per term (autocast numerics)")
    ap.add_argument("--cpu-points", type=int, default=0, help="points per CPU forward (0 = the full workload; a smaller value disables `parity`)")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
