#!/usr/bin/env python
"""bench.py -- points/sec of the CDSegNet single-step forward (BASELINE.json metric) on B200.

  python bench.py --gpus N --steps K --warmup W            # the CUDA path (this repo)
  python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host cores (CPU oracle port)

A "step" = one full DefaultSegmentorV2.inference(eval=False) of one synthetic ScanNet-shaped scene of
120 000 unique voxels (BASELINE.json configs[1]) through the full CDSegNet (CN + NN + TransferModule,
101.4 M parameters): serialization (key encode + 4 radix argsorts), pooling hierarchy, 37 blocks, heads.
Scenes shard one per GPU (no collective on the inference forward) => weak scaling.
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


def make_scene(seed, n=N_POINTS):
    from cdsegnet_b200 import synth
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


# --------------------------------------------------------------------------------------------
def cpu_forward_timed(n_points, seed=0, repeats=1):
    """the oracle (CPU port of the reference forward, dense fp32 attention) on the host cores."""
    import numpy as np
    import torch
    from oracle import ptv3_oracle as O
    import cdsegnet_b200 as cb
    from cdsegnet_b200 import configs
    # the port is a chain of small torch ops: beyond ~16 threads intra-op parallelism only adds contention
    # (measured: 128 threads on the GPU box's host were 10x slower than 16), so use min(cores, 16)
    cores = min(os.cpu_count(), 16)
    torch.set_num_threads(cores)
    cfg = configs.backbone_cfg()
    torch.manual_seed(0)
    model = cb.PointTransformerV3(**cfg)
    random_weights(model)
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    sc = make_scene(seed, n_points)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    base = dict(coord=t(sc["coord"]), grid_coord=t(sc["grid_coord"]).long(), offset=t(sc["offset"]))
    n = len(sc["coord"])
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        noise = torch.normal(0, 1, size=(n, 6))
        ts = 999 * torch.ones((n, 1), dtype=torch.int64)
        O.forward(sd, cfg, dict(base, feat=noise, t_emb=O.calc_t_emb(ts, 128)), dict(base, feat=t(sc["feat"])), attn_mode="dense")
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return n / best, cores, best


def reference_points(args):
    """points per CPU step.  cpu_baseline leg (one forward): the whole 120k workload (~13-25 s of CPU work).  Reference arm:
    the largest scene of the same generator that keeps (steps + warmup) forwards within a few minutes at the oracle's measured
    ~0.2 ms per point (8-16 host cores), never below 20k points."""
    if args.cpu_points:
        return args.cpu_points
    if args.impl != "reference":
        return N_POINTS
    n = int(200.0 / ((args.steps + args.warmup) * 2.0e-4))
    return max(20000, min(N_POINTS, n // 1000 * 1000))


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    n = reference_points(args)
    vals = []
    for _ in range(args.warmup):
        cpu_forward_timed(n)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        v, cores, dt = cpu_forward_timed(n)
        vals.append(dt)
    total = time.perf_counter() - t0
    value = n * args.steps / sum(vals)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "points/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(vals) / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "ScanNet-shaped scene 120k unique voxels @0.02 m, full CDSegNet (CN+NN+TransferModule, 101.4M params), "
                                   "single-step inference forward, patch 1024, 1 scene per step; CPU oracle port (dense fp32 attention) on a "
                                   f"{n}-point scene of the same generator",
                       "points_per_step": n},
            "cpu_baseline": {"value": value, "unit": "points/s", "cores": cores, "kind": "port",
                             "sample": f"{args.steps} forward(s) of a {n}-point scene (bounded sample of the 120k workload)"},
            "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": total}
    print(json.dumps(line))


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

    torch.manual_seed(0)
    seg = cb.build_model(configs.segmentor_cfg())           # enable_flash=True: the shipped config (fp16 tensor-core attention)
    random_weights(seg)
    seg = seg.to(dev).eval()

    sc = make_scene(seed=0)           # the SAME scene on every rank: weak scaling keeps the per-GPU work identical (different seeds give
                                      # 120k-point scenes whose step times differ by up to 8 %, profiles/r01f_bench_4gpu.json)
    n = len(sc["coord"])
    host = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in sc.items()}
    host["grid_coord"] = host["grid_coord"].int().pin_memory()
    resident = {k: v.to(dev) for k, v in host.items()}
    noise_dev = torch.randn(n, 6, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2

    def step_resident():
        return seg.inference(resident, eval=False, noise=noise_dev)["seg_logits"]

    logits_host = torch.empty((n, 20), dtype=torch.float32).pin_memory()

    def step_e2e():
        inp = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
        out = seg.inference(inp, eval=False)["seg_logits"]              # draws the NN noise like default.py:393
        logits_host.copy_(out, non_blocking=True)
        return out

    rank_ms = {}

    def timed(fn, steps, warmup, profile_attn=False):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        evs = []
        ops.launch_count_reset()
        if profile_attn:
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
        ms = sum(a.elapsed_time(b) for a, b in evs)
        prof = ops.PROFILE
        ops.PROFILE = None
        tsum = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            allr = [torch.zeros_like(tsum) for _ in range(world)]
            dist.all_gather(allr, tsum)
            rank_ms.setdefault(fn.__name__, [float(a.item()) / steps for a in allr])     # first call = the headline timed run
            dist.all_reduce(tsum, op=dist.ReduceOp.MAX)
        return float(tsum.item()), launches, wall, prof

    sampler = ClockSampler(local)
    sampler.start()
    ms_total, launches, wall, prof = timed(step_resident, args.steps, args.warmup, profile_attn=True)
    clocks = sampler.summary()
    ms_e2e, _, _, _ = timed(step_e2e, args.steps, max(1, args.warmup // 2))

    value = world * n * args.steps / (ms_total / 1e3)
    e2e = world * n * args.steps / (ms_e2e / 1e3)
    hbm, tf_burst, tf_sust, which = peaks()

    # Per-kernel numbers, measured live with CUDA events on the launching stream inside the timed steps above (the events are
    # recorded by cdseg_block_forward around its own launches).
    #  * dominant launch configuration by device time (profiles/r01d_launches_step_v4.md) = fz::pre_kernel at stage 0: the fused
    #    cpe conv + Linear + LayerNorm + residual + norm1 + qkv chain, 12 launches/step.  Its roofline is HBM: algorithmic bytes =
    #    read the block input once (n*C*4; conv operand and residual are the same tensor) + the neighbour table (n*27*4) +
    #    write x1 (n*C*4) and qkv (n*3C*4) + the packed weights once.  The tensor-pipe view of the same launch is reported next
    #    to it (FLOPs = 2 * (conv pairs + 4n) * C^2, x3 MMAs for the fp16 hi/lo split are NOT counted).
    #  * the one dense contraction north_star names = tc2::attn_tc2_kernel at stage 0 (tensor/MUFU bound).
    #  * `traffic` = dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full captures
    #    (profiles/r01d_ncu_*.md, same workload, same launch).
    NCU_TRAFFIC = {"pre": 23.377152e6 + 5.710592e6, "attn": 31.441664e6 + 0.014848e6}

    def kernel_lines(prof, steps):
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
        attn = {"bound": "tensor", "kernel": "tc2::attn_tc2_kernel (stage 0, %d launches/step)" % (len(sel) // steps),
                "achieved": fl / (t_attn * 1e-3) / 1e12, "peak": tf_sust, "unit": "TFLOP/s", "frac": fl / (t_attn * 1e-3) / 1e12 / tf_sust,
                "flops_per_launch": fl, "ms_per_launch": t_attn, "exp_per_launch": ex, "gexp_per_s": ex / (t_attn * 1e-3) / 1e9,
                "mufu_peak_gexp_per_s": 148 * 16 * 1.965, "traffic": NCU_TRAFFIC["attn"]}
        wbytes = (27 + 1 + 3) * C0 * C0 * 4
        by = n0 * C0 * 4 + n0 * 27 * 4 + n0 * C0 * 4 + n0 * 3 * C0 * 4 + wbytes
        pre_fl = 2.0 * (conv_pairs + 4 * n0) * C0 * C0
        roof = {"bound": "hbm", "kernel": "fz::pre_kernel (stage 0: cpe conv + Linear + LN + residual + norm1 + qkv fused, n=%d C=%d, %d launches/step)"
                % (n0, C0, len(sel) // steps),
                "achieved": by / (t_pre * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s", "frac": by / (t_pre * 1e-3) / 1e9 / hbm,
                "peak_source": which, "bytes_per_launch": by, "ms_per_launch": t_pre, "traffic": NCU_TRAFFIC["pre"],
                "tensor_view": {"flops_per_launch": pre_fl, "achieved_tflops": pre_fl / (t_pre * 1e-3) / 1e12,
                                "frac_of_bf16_peak": pre_fl / (t_pre * 1e-3) / 1e12 / tf_sust}}
        pby = 3 * n0 * C0 * 4 + 9 * C0 * C0 * 4
        post = {"bound": "hbm", "kernel": "fz::post_kernel (stage 0: proj + residual + norm2 + fc1 + GELU + fc2 + residual fused)",
                "achieved": pby / (t_post * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s", "frac": pby / (t_post * 1e-3) / 1e9 / hbm,
                "bytes_per_launch": pby, "ms_per_launch": t_post, "traffic": None}
        for p in prof:
            for e in p["ev"]:
                lib.cdseg_event_destroy(e)
        return roof, attn, post

    roof = attn = post = None
    if prof:
        nb3 = seg.backbone.last_plan.n_levels[0].nbr(3)
        conv_pairs = int((nb3 >= 0).sum().item())                      # (active output, active input) pairs of the k=3 conv at level 0
        roof, attn, post = kernel_lines(prof, args.steps)
        # In the timed region above the Noise Network runs on a second stream beside the Conditional Network, so the
        # events around one launch also cover whatever the other stream had resident.  A few extra steps with the
        # two-stream schedule switched off give the same launches alone on the device (reported next to the in-step time).
        seg.backbone.overlap_streams = False
        k1 = max(2, min(5, args.steps))
        _, _, _, prof1 = timed(step_resident, k1, 1, profile_attn=True)
        seg.backbone.overlap_streams = True
        r1, a1, p1 = kernel_lines(prof1, k1)
        for full, alone in ((roof, r1), (attn, a1), (post, p1)):
            full["single_stream"] = {"ms_per_launch": alone["ms_per_launch"], "achieved": alone["achieved"], "frac": alone["frac"]}

    line = {"metric": METRIC, "value": value, "unit": "points/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (dense layers, conv, norms) + f16 tensor-core attention with f32 accumulate (reference flash branch)",
            "data": "synthetic",
            "config": {"workload": "ScanNet-shaped scene 120k unique voxels @0.02 m, full CDSegNet (CN+NN+TransferModule, 101.4M params), "
                                   "single-step inference forward, patch 1024, 1 scene per GPU",
                       "points_per_step_per_gpu": n, "l2": "flushed (256 MiB write) between timed iterations",
                       "parallelism": f"scene-per-GPU x{world} (same synthetic scene on every rank), no collective"},
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": e2e, "unit": "points/s",
                    "h2d_bytes_per_step": int(sum(v.numel() * v.element_size() for v in host.values()) + n * 6 * 4),
                    "d2h_bytes_per_step": int(n * 20 * 4), "ms_per_step": ms_e2e / args.steps},
            "roofline": roof, "roofline_attention": attn, "roofline_post": post}
    if world > 1:
        line["ms_per_step_by_rank"] = rank_ms.get("step_resident")
    if rank == 0:
        if world == 1 and not args.no_cpu:
            npts = reference_points(args)
            v, cores, dt = cpu_forward_timed(npts)
            line["cpu_baseline"] = {"value": v, "unit": "points/s", "cores": cores, "kind": "port",
                                    "sample": f"1 forward of a {npts}-point scene, same model, oracle port with dense fp32 attention ({dt:.1f} s)"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--cpu-points", type=int, default=0, help="points per CPU forward (0 = auto, see reference_points)")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
