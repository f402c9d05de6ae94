#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native CUDA-Learn-Notes hot paths (contract: see task / DESIGN.md §6).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--quick]
  torchrun --nproc-per-node N ... bench.py --gpus N ...          (one rank per GPU, NCCL)

One JSON line on stdout (rank 0).  Headline `value` = HGEMM TFLOPS on BASELINE config #2 (fp16 NN, M=N=K=8192,
the metric BASELINE.json is quoted on); a "step" is one GEMM with A, B, C resident in HBM.  HGEMM does not shard
(SURVEY §8e): with --gpus N every rank runs an independent replica (weak scaling, value = sum of flops / max time).
The same line carries
  sweep       HGEMM at 2048 / 4096 / 8192 / 16384 (+ torch.matmul = cuBLAS on the same box)
  attention   FA-2 forward: config #3 (4,48,8192,64) and the config #5 shard; at N > 1 config #5 (32,64,8192,128) batch-
              sharded over the ranks with ONE NCCL broadcast of the packed inputs (timed separately), no reduction
  ffpa        FFPA forward config #4 (1,32,4096,512)
  ref_gpu     the reference's own mma.sync kernels (oracle/_ref, built from /root/reference) timed in the same run
  e2e, roofline, cpu_baseline, clocks, gpu_launches     as the contract defines them.
`--impl reference` times the reference's CPU-runnable path (torch.matmul on the host cores, the comparator its own
scripts use — kernels/sgemm/sgemm.py:L135, kernels/hgemm/hgemm.py:L349) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "cuda-learn-notes_b200"))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

HGEMM_MNK = 8192
SWEEP = (2048, 4096, 8192, 16384)
FA2_CFG3 = (4, 48, 8192, 64)
FA2_CFG5 = (32, 64, 8192, 128)
FFPA_CFG4 = (1, 32, 4096, 512)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"tflops_burst": d.get("bf16_tflops"), "tflops_sustained": d.get("bf16_tflops_sustained"),
                "hbm_gbs": d.get("hbm_gbs"), "source": "MEASURED_PEAKS.json (of measured)"}
    return {"tflops_burst": 1590.0, "tflops_sustained": 1400.0, "hbm_gbs": 6650.0,
            "source": "B200_PROFILING.md fallback (of fallback)"}


class ClockSampler:
    """Samples SM clock / power / throttle reasons through NVML every ~2 ms while the timed region runs (the timed
    region of a GEMM step is tens of milliseconds: `nvidia-smi -lms` is too coarse for it)."""

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.stop_flag = False
        self.thread = None
        self.err = None

    def _loop(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            while not self.stop_flag:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                self.samples.append((time.perf_counter(), sm, reasons, pw))
                time.sleep(0.002)
        except Exception as e:  # noqa
            self.err = repr(e)[:200]

    def start(self):
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()
        time.sleep(0.02)

    def mark(self):
        return time.perf_counter()

    def stop(self, t0=None, t1=None):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=2)
        rows = [r for r in self.samples if (t0 is None or r[0] >= t0) and (t1 is None or r[0] <= t1)] or self.samples
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: %s" % self.err]}
        sm = sorted(r[1] for r in rows)
        bits = 0
        for r in rows:
            bits |= int(r[2])
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
                 0x80: "hw_power_brake_slowdown"}
        reasons = sorted(n for b, n in names.items() if bits & b)
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": getattr(self, "max_mhz", None),
                "power_w_max": max(r[3] for r in rows), "samples": len(rows), "reasons": reasons,
                "how": "NVML polled every ~2 ms during the timed steps"}


def cuda_time(fn, steps, warmup, barrier=None):
    """W warm-up calls, then exactly K calls timed with CUDA events on the current (launching) stream."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if barrier:
        barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    if barrier:
        barrier()
    return e0.elapsed_time(e1) / steps  # ms per step


def graph_time(fn, launches_per_graph, replays=5, warm_replays=1):
    """ms per call of `fn` when `launches_per_graph` calls are captured into ONE CUDA graph and the graph is replayed:
    kernel time without the per-call host cost.  Warm-up runs on the capture stream (per-stream workspaces)."""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    s.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        for _ in range(launches_per_graph):
            fn()
    return cuda_time(g.replay, replays, warm_replays) / launches_per_graph


def max_over_ranks(ms, dist_on):
    if not dist_on:
        return ms
    import torch.distributed as dist
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def cpu_info():
    model = ""
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                model = line.split(":", 1)[1].strip()
                break
    except Exception:
        pass
    return {"cores": os.cpu_count(), "torch_threads": torch.get_num_threads(), "model": model}


# ---------------------------------------------------------------------------------------------------- CPU legs
def cpu_hgemm_sample(n, budget_s=12.0, samples=3):
    """torch.matmul on the host cores on a row-slab sample of the n^3 problem (same K, same B).  All cores, two untimed
    warm-up products (thread pool spin-up, first touch of the operands), then `samples` timed samples of budget_s /
    samples seconds each; the MEDIAN sample is reported with the spread (the round-1 single 12 s sample moved by 10x
    between boxes)."""
    torch.set_num_threads(os.cpu_count() or 1)
    rows = 256
    cache = cpu_hgemm_sample.__dict__.setdefault("cache", {})
    if n not in cache:
        torch.manual_seed(1)
        cache[n] = (torch.randn(rows, n, dtype=torch.float32), torch.randn(n, n, dtype=torch.float32))
    a, b = cache[n]
    torch.matmul(a, b)
    torch.matmul(a, b)  # warm-up x2
    vals = []
    total_reps, total_dt = 0, 0.0
    for _ in range(samples):
        t0 = time.perf_counter()
        reps = 0
        while True:
            torch.matmul(a, b)
            reps += 1
            dt = time.perf_counter() - t0
            if dt > budget_s / samples or reps >= 50:
                break
        vals.append(2.0 * rows * n * n * reps / dt * 1e-12)
        total_reps += reps
        total_dt += dt
    vals.sort()
    tflops = vals[len(vals) // 2]
    return tflops, ("torch.matmul fp32 on host, %d threads: %d x (%d x %d) @ (%d x %d) row-slab of the %d^3 problem in %d samples, "
                    "%.1f s; median %.3f, min %.3f, max %.3f TFLOP/s" % (torch.get_num_threads(), total_reps, rows, n, n, n, n,
                                                                         samples, total_dt, tflops, vals[0], vals[-1]))


def cpu_sgemm_config1():
    """BASELINE config #1: SGEMM fp32 1024^3 through the reference's own CPU-runnable path, torch.matmul on host tensors
    (kernels/sgemm/sgemm.py:L135) - plumbing, no GPU.  Median of 5 after 2 warm-ups."""
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(1)
    a, b = torch.randn(1024, 1024), torch.randn(1024, 1024)
    torch.matmul(a, b)
    torch.matmul(a, b)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        c = torch.matmul(a, b)
        ts.append(time.perf_counter() - t0)
    ts.sort()
    ref = (a.double() @ b.double())
    return {"workload": "sgemm_f32_m1024_n1024_k1024 on host CPU (BASELINE config #1)", "ms": ts[2] * 1e3,
            "gflops": 2.0 * 1024 ** 3 / ts[2] * 1e-9, "ms_min_max": [ts[0] * 1e3, ts[-1] * 1e3], "threads": torch.get_num_threads(),
            "max_abs_err_vs_fp64": float((c.double() - ref).abs().max())}


def load_traffic(key):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the named kernel from the committed ncu
    summary of this round (profiles/r02_roofline_traffic.json, written by tools/ncu_summary.py from an
    `ncu --set full` capture of this command's kernels); None when no capture is committed."""
    p = os.path.join(ROOT, "profiles", "r02_roofline_traffic.json")
    try:
        d = json.load(open(p))
        e = d.get(key)
        return (e["dram_bytes"], "profiles/r02_roofline_traffic.json: " + e.get("source", "")) if e else (None, None)
    except Exception:
        return None, None


def roofline_obj(kernel, flops, ms, peaks, traffic_key, algo_bytes, timed_region_ms):
    """`roofline` for one tensor-bound kernel.  Peak: the burst cuBLAS figure when the whole timed region is shorter than
    ~100 ms (the GPU is still on its boost clocks), the sustained one for longer regions (both from MEASURED_PEAKS.json)."""
    ach = flops / (ms * 1e-3) * 1e-12
    burst = timed_region_ms < 100.0
    peak = peaks["tflops_burst"] if burst else peaks["tflops_sustained"]
    traffic, src = load_traffic(traffic_key)
    return {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
            "peak_regime": "burst" if burst else "sustained", "frac_of_burst": ach / peaks["tflops_burst"],
            "frac_of_sustained": ach / peaks["tflops_sustained"], "peak_source": peaks["source"],
            "timed_region_ms": timed_region_ms, "kernel": kernel, "algorithmic_flops_per_launch": flops,
            "algorithmic_bytes_per_launch": algo_bytes, "traffic": traffic, "traffic_source": src}


def headline_config():
    """`config` of the JSON line - identical for both arms (the driver compares them)."""
    return {"workload": "hgemm_nn_f16_m8192_n8192_k8192", "baseline_config": "#2 HGEMM fp16 NN square",
            "multi_gpu": "replicas only (a single GEMM does not shard without a collective)",
            "l2": "operands 3 x 128 MiB > 126 MB L2 (inputs larger than L2, no flush needed)",
            "preconditioning": "40 untimed launches + 1 s pause before the W warm-up steps (GPU out of its idle P-state)",
            "randn_seed": 1}


def run_reference_impl(args, world, rank):
    """--impl reference: the reference's CPU-runnable path (torch.matmul on host cores), rank 0 only."""
    if rank != 0:
        return
    info = cpu_info()
    n = HGEMM_MNK
    vals = []
    sample = ""
    total = args.warmup + args.steps
    per = max(1.0, min(10.0, 150.0 / max(total, 1)))
    for i in range(total):
        v, sample = cpu_hgemm_sample(n, budget_s=per)
        if i >= args.warmup:
            vals.append(v)
    v = sum(vals) / len(vals)
    ms = 2.0 * n ** 3 / (v * 1e12) * 1e3
    line = {"impl": "reference", "metric": "hgemm_tflops", "value": v, "unit": "TFLOP/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 (cpu)", "data": "synthetic",
            "config": headline_config(),
            "reference_note": "reference arm = torch.matmul on host cores (the reference's own CPU-runnable comparator, "
                              "kernels/hgemm/hgemm.py:L349); each step is a bounded row-slab sample, ms_per_step extrapolated",
            "cpu_baseline": {"value": v, "unit": "TFLOP/s", "cores": info["cores"], "kind": "port", "sample": sample,
                             "cpu": info["model"]},
            "e2e": {"value": v, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------- reference GPU kernels
def load_ref_hgemm():
    import ctypes
    p = os.path.join(ROOT, "oracle", "_ref", "libref_hgemm.so")
    if not os.path.exists(p):
        return None
    try:
        lib = ctypes.CDLL(p)
        lib.ref_hgemm_mma_stages_dsmem_nn.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] * 5
        lib.ref_hgemm_mma_stages_dsmem_nn.restype = ctypes.c_int
        return lib
    except OSError:
        return None


def load_ref_module(name):
    import importlib.util
    p = os.path.join(ROOT, "oracle", "_ref", name + ".so")
    if not os.path.exists(p):
        return None
    try:
        spec = importlib.util.spec_from_file_location(name, p)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod
    except Exception as e:  # noqa
        return None


def ref_swizzle_stride(N):  # hgemm.py:L71-81
    f = 0.5 if N <= 4096 else 0.25
    s = int(N * f)
    return s if s >= 256 else 1


# ---------------------------------------------------------------------------------------------------- config #5 sharded
NVLINK_MEASURED_GBPS = 770.0   # B200_PROFILING.md: measured peer copy per direction per GPU (nominal 900)


def bench_sharded_cfg5(world, rank, dev, barrier, dist_on, peaks, reps=5, warm=2):
    """BASELINE config #5, FA-2 forward (32,64,8192,128), batch-sharded over `world` GPUs.
    Rank 0 holds Q, K, V (12.9 GB packed).  Timed, max over ranks, CUDA events, `warm` untimed + `reps` timed passes each:
      compute only (inputs already distributed), the three input distributions alone, and distribution + compute end to end.
    Rank 0 also runs the WHOLE problem alone (the N = 1 point on this box) and every shard is compared bit for bit with
    that result."""
    from b200k import ops, sharded
    B_, H_, N_, D_ = FA2_CFG5
    fl = 4.0 * B_ * H_ * N_ * N_ * D_
    shape = (B_, H_, N_, D_)
    rec = {"workload": "fa2_fwd_b32_h64_n8192_d128 (BASELINE config #5)", "n_gpus": world, "scaling": "strong",
           "algorithmic_flops": fl, "timing": "CUDA events, %d warm-up + %d timed passes, max over ranks" % (warm, reps)}
    n_launch = 0

    def timed(fn, r=reps, w=warm):
        for _ in range(w):
            fn()
        torch.cuda.synchronize()
        if barrier:
            barrier()
        ts = []
        for _ in range(r):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if barrier:
                barrier()
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(max_over_ranks(e0.elapsed_time(e1), dist_on))
        ts.sort()
        return ts[len(ts) // 2], ts[0], ts[-1]

    qkv = None
    o_full = None
    if rank == 0:
        g = torch.Generator(device=dev).manual_seed(5)
        qkv = torch.empty(3, B_, H_, N_, D_, dtype=torch.half, device=dev)
        for t in range(3):
            qkv[t].normal_(generator=g)
        o_full = torch.empty(B_, H_, N_, D_, dtype=torch.half, device=dev)
    # ---- the N = 1 point: rank 0 alone runs the whole problem as ONE call (2^31 elements per tensor)
    t_n1 = [None]
    if rank == 0:
        for _ in range(warm):
            ops.fa2_fwd(qkv[0], qkv[1], qkv[2], o_full)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            ops.fa2_fwd(qkv[0], qkv[1], qkv[2], o_full)
        e1.record()
        torch.cuda.synchronize()
        t_n1[0] = e0.elapsed_time(e1) / 3
        n_launch += warm + 3
    if dist_on:
        import torch.distributed as dist
        dist.broadcast_object_list(t_n1, src=0)
    t1 = t_n1[0]
    rec["n1_whole_problem_ms"] = t1
    rec["n1_tflops"] = fl / t1 * 1e-9
    if not dist_on:
        rec.update({"compute_ms": t1, "aggregate_tflops": fl / t1 * 1e-9, "aggregate_tflops_incl_distribution": fl / t1 * 1e-9,
                    "efficiency_vs_n1": 1.0, "note": "N = 1: no distribution; the whole (32,64,8192,128) problem as one "
                    "kernel launch on one GPU (16 GiB of tensors)"})
        rec["_launches"] = n_launch
        return rec

    lo, hi = sharded.shard_bounds(B_, world, rank)
    o = torch.empty(hi - lo, H_, N_, D_, dtype=torch.half, device=dev)
    qkv_bytes = 3 * B_ * H_ * N_ * D_ * 2
    modes = {}
    keep = {}
    for mode in sharded.MODES:
        # distribution alone (the handles are waited for on the stream; the events bracket the transfer itself)
        def dist_only(mode=mode):
            sh = sharded.distribute_qkv(qkv, shape, dev, mode=mode, chunk_batches=1)
            sh.wait_all()
            keep["sh"] = sh
        td, td_min, td_max = timed(dist_only)
        sh = keep["sh"]
        # compute alone on the distributed shard
        tc, _, _ = timed(lambda: sharded.attention_on_shard(sh, out=o))
        n_launch += (warm + reps) * max(1, len(sh.chunks))
        ok = sharded.shards_equal_to(o, o_full, B_)
        del sh
        keep.clear()
        # distribution + compute, what a caller of config #5 pays
        def both(mode=mode):
            sharded.sharded_attention(qkv, shape, dev, mode=mode, out=o, chunk_batches=1)
        tb, tb_min, tb_max = timed(both)
        n_launch += (warm + reps) * (1 if mode != "pipelined" else (hi - lo))
        ok2 = sharded.shards_equal_to(o, o_full, B_)
        sent = qkv_bytes if mode == "broadcast" else qkv_bytes * (world - 1) / world
        recv = qkv_bytes if mode == "broadcast" else qkv_bytes / world
        modes[mode] = {
            "distribution_ms": td, "distribution_ms_min_max": [td_min, td_max],
            "source_egress_GBps": sent / td * 1e-6, "receiver_ingress_GBps": recv / td * 1e-6,
            "frac_of_nvlink_measured_770GBps": sent / td * 1e-6 / NVLINK_MEASURED_GBPS,
            "frac_of_nvlink_nominal_900GBps": sent / td * 1e-6 / 900.0,
            "bytes_leaving_source": sent, "bytes_per_receiver": recv,
            "compute_ms": tc, "aggregate_tflops_compute_only": fl / tc * 1e-9,
            "total_ms": tb, "total_ms_min_max": [tb_min, tb_max],
            "aggregate_tflops_incl_distribution": fl / tb * 1e-9,
            "speedup_vs_n1_incl_distribution": t1 / tb, "efficiency_vs_n1_incl_distribution": t1 / tb / world,
            "bit_equal_to_single_gpu_run": bool(ok and ok2)}
        torch.cuda.empty_cache()
    d = modes["broadcast"]
    best = min(modes, key=lambda m: modes[m]["total_ms"])
    floor_ms = qkv_bytes * (world - 1) / world / (NVLINK_MEASURED_GBPS * 1e6)
    rec.update({
        "default_mode": "broadcast (one NCCL broadcast of the packed QKV, the north star's wording)",
        "compute_ms": d["compute_ms"], "bcast_ms": d["distribution_ms"], "bcast_GBps": d["source_egress_GBps"],
        "aggregate_tflops": d["aggregate_tflops_compute_only"],
        "aggregate_tflops_incl_distribution": d["aggregate_tflops_incl_distribution"],
        "efficiency_vs_n1": t1 / d["compute_ms"] / world,
        "efficiency_vs_n1_incl_distribution": d["efficiency_vs_n1_incl_distribution"],
        "best_mode": best, "best_total_ms": modes[best]["total_ms"],
        "best_aggregate_tflops_incl_distribution": modes[best]["aggregate_tflops_incl_distribution"],
        "best_speedup_vs_n1_incl_distribution": modes[best]["speedup_vs_n1_incl_distribution"],
        "distribution_floor_ms": floor_ms,
        "distribution_floor": "inputs start on ONE GPU: (G-1)/G of 12.9 GB must leave rank 0 through its own NVLink ports "
                              "(measured 770 GB/s per direction); no scheme can finish before that",
        "best_total_vs_floor": floor_ms / modes[best]["total_ms"],
        "parity": "every shard bit-equal to rank 0's single-GPU one-call result" if all(
            m["bit_equal_to_single_gpu_run"] for m in modes.values()) else "MISMATCH",
        "modes": modes})
    rec["_launches"] = n_launch
    del o, qkv, o_full
    torch.cuda.empty_cache()
    return rec


# ---------------------------------------------------------------------------------------------------- our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--quick", action="store_true", help="headline only (skip sweep / attention / ffpa / ref_gpu)")
    ap.add_argument("--sections", default="all", help="comma list of the extra sections to run: sweep,attention,ffpa,widened,"
                    "sharded (default all; the headline, e2e, roofline and cpu_baseline always run)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist_on = world > 1

    if args.impl == "reference":
        run_reference_impl(args, world, rank)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device. The product path has no CPU fallback; use --impl reference for the CPU arm.")
    torch.cuda.set_device(local_rank)
    if dist_on:
        # stdout carries exactly one JSON line, and NCCL_DEBUG output goes to stdout by default: send it to a per-rank
        # file instead and replay it on stderr at the end (so "nranks N" / "Init COMPLETE" stay checkable).
        nccl_log = None
        if os.environ.get("NCCL_DEBUG") and not os.environ.get("NCCL_DEBUG_FILE"):
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            nccl_log = os.path.join(ROOT, "gpurun_out", "nccl_debug_n%d_rank%d.log" % (world, rank))
            os.environ["NCCL_DEBUG_FILE"] = nccl_log
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        barrier = lambda: (dist.barrier(), torch.cuda.synchronize())  # noqa: E731
    else:
        barrier = None
    from b200k import ops  # the C-ABI library; raises if missing

    peaks = load_peaks()
    dev = torch.device("cuda", local_rank)
    torch.manual_seed(1 + rank)
    launches = 0
    out = {}

    # ------------------------------------------------------------------ headline: HGEMM 8192^3 (or selected workload)
    n = HGEMM_MNK
    a = torch.randn(n, n, dtype=torch.half, device=dev)
    b = torch.randn(n, n, dtype=torch.half, device=dev)
    c = torch.zeros(n, n, dtype=torch.half, device=dev)
    flops = 2.0 * n ** 3

    def step():
        ops.hgemm(a, b, c)

    sampler = ClockSampler(torch.cuda.current_device() if "CUDA_VISIBLE_DEVICES" not in os.environ else local_rank)
    if rank == 0 and not os.environ.get("B200K_BENCH_NO_SAMPLER"):   # (diagnostic switch: does NVML polling perturb the GEMM?)
        sampler.start()
    # Preconditioning (untimed, before the W warm-up steps): a fresh process finds the GPU in its idle P-state (the lease
    # record shows 120 MHz), and the first milliseconds of work run while the clocks are still ramping.  ~30 ms of the same
    # GEMM wake it up, a short pause lets the power-cap controller settle, then the contract's W warm-ups and K timed
    # steps follow.  Measured on one box, same process: 1555 TFLOP/s without this, 1651 with (the regime every A/B number
    # in profiles/r02_hgemm_ab_bench_protocol.jsonl was taken in).
    for _ in range(40):
        step()
    torch.cuda.synchronize()
    time.sleep(1.0)
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    t_mark0 = sampler.mark()
    ms = cuda_time(step, args.steps, 0, barrier)
    t_mark1 = sampler.mark()
    clocks = sampler.stop(t_mark0, t_mark1) if rank == 0 else None
    launches += args.steps
    ms_max = max_over_ranks(ms, dist_on)
    value = flops * world / (ms_max * 1e-3) * 1e-12

    # ---- e2e: the same GEMM through the public API with HOST (pinned) buffers, H2D + D2H inside the timed region.
    # Two figures: (1) serial - copy in, multiply, copy out, one step after the other on one stream; (2) pipelined, the
    # headline - what a host-fed consumer of the library would run: three streams and two sets of device buffers, so the
    # H2D copy of step k+1 and the D2H copy of step k-1 overlap the GEMM of step k (PCIe is full duplex).  Every step
    # still moves its own 256 MiB of operands in and 128 MiB of result out inside the timed region.
    ah = torch.randn(n, n, dtype=torch.half).pin_memory()
    bh = torch.randn(n, n, dtype=torch.half).pin_memory()
    ch = [torch.empty(n, n, dtype=torch.half).pin_memory() for _ in range(2)]

    def step_e2e():
        a.copy_(ah, non_blocking=True)
        b.copy_(bh, non_blocking=True)
        ops.hgemm(a, b, c)
        ch[0].copy_(c, non_blocking=True)

    e2e_steps = max(4, min(args.steps, 10))
    ms_e2e_serial = max_over_ranks(cuda_time(step_e2e, e2e_steps, 3, barrier), dist_on)

    dbuf = [(a, b, c), (torch.empty_like(a), torch.empty_like(b), torch.empty_like(c))]
    s_main = torch.cuda.current_stream()
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_mm = [torch.cuda.Event() for _ in range(2)]
    ev_out = [torch.cuda.Event() for _ in range(2)]

    def pipelined(k_steps):
        for k in range(k_steps):
            slot = k & 1
            A_, B_, C_ = dbuf[slot]
            with torch.cuda.stream(s_in):
                s_in.wait_event(ev_mm[slot])      # the GEMM that last read this slot's operands is done
                A_.copy_(ah, non_blocking=True)
                B_.copy_(bh, non_blocking=True)
                ev_in[slot].record(s_in)
            s_main.wait_event(ev_in[slot])
            s_main.wait_event(ev_out[slot])       # the D2H copy that last read this slot's C is done
            ops.hgemm(A_, B_, C_)                 # launches on the current (main) stream
            ev_mm[slot].record(s_main)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_mm[slot])
                ch[slot].copy_(C_, non_blocking=True)
                ev_out[slot].record(s_out)
        s_main.wait_stream(s_in)
        s_main.wait_stream(s_out)

    pipelined(3)
    torch.cuda.synchronize()
    if barrier:
        barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    pipelined(e2e_steps)
    t1.record()
    torch.cuda.synchronize()
    if barrier:
        barrier()
    ms_e2e = max_over_ranks(t0.elapsed_time(t1) / e2e_steps, dist_on)
    launches += 2 * e2e_steps + 6
    e2e_ok = bool(torch.equal(ch[0], ch[1]))  # both slots computed the same product from the same host operands
    e2e = {"value": flops * world / (ms_e2e * 1e-3) * 1e-12, "unit": "TFLOP/s", "ms_per_step": ms_e2e,
           "h2d_bytes_per_step": 2 * n * n * 2, "d2h_bytes_per_step": n * n * 2,
           "api": "b200k.ops.hgemm(a, b, c) (the toy_hgemm-style call) fed from pinned host buffers; 2-deep pipeline on "
                  "three CUDA streams: H2D of step k+1 and D2H of step k-1 overlap the GEMM of step k",
           "serial_ms_per_step": ms_e2e_serial, "serial_value": flops * world / (ms_e2e_serial * 1e-3) * 1e-12,
           "results_identical_across_slots": e2e_ok}
    del ah, bh, ch, dbuf

    roofline = roofline_obj("hgemm_tcgen05_kernel (b200k_hgemm_f16, variant AUTO: 512x256 pair tile at this size)", flops, ms, peaks,
                            "hgemm_8192", 3.0 * n * n * 2, ms * args.steps)
    rooflines = {"hgemm_8192": roofline}

    want = (lambda name: args.sections == "all" or name in args.sections.split(","))
    if not args.quick:
        # -------------------------------------------------------------- HGEMM sweep + cuBLAS (torch.matmul) + reference mma.sync
        if want("sweep"):
            ref_h = load_ref_hgemm()
            sweep = []
            for m in SWEEP:
                A = a[:m, :m].contiguous() if m <= n else torch.randn(m, m, dtype=torch.half, device=dev)
                B = b[:m, :m].contiguous() if m <= n else torch.randn(m, m, dtype=torch.half, device=dev)
                C = torch.empty(m, m, dtype=torch.half, device=dev)
                it = 20 if m <= 8192 else 5
                t_ours = cuda_time(lambda: ops.hgemm(A, B, C), it, 3)
                launches += it + 3
                t_cublas = cuda_time(lambda: torch.matmul(A, B, out=C), it, 3)
                t_ours_g = t_cublas_g = None
                graph_err = None
                if m <= 4096 and not dist_on:
                    # a 17 us kernel is launch-bound from Python: the same `it` launches captured once and replayed.
                    # Single-process runs only (no collective library thread next to a stream capture); a failed capture
                    # is recorded, it does not take the bench line down.
                    try:
                        t_ours_g = graph_time(lambda: ops.hgemm(A, B, C), it)
                        t_cublas_g = graph_time(lambda: torch.matmul(A, B, out=C), it)
                        launches += 2 + it * 6
                    except Exception as e:  # noqa: BLE001
                        t_ours_g = t_cublas_g = None
                        graph_err = repr(e)[:160]
                row = {"mnk": m, "tflops": 2.0 * m ** 3 / t_ours * 1e-9, "cublas_tflops": 2.0 * m ** 3 / t_cublas * 1e-9,
                       "frac_of_peak_burst": 2.0 * m ** 3 / t_ours * 1e-9 / peaks["tflops_burst"],
                       "frac_of_peak_sustained": 2.0 * m ** 3 / t_ours * 1e-9 / peaks["tflops_sustained"],
                       "timed_region_ms": t_ours * it, "peak_regime": "burst" if t_ours * it < 100.0 else "sustained"}
                if t_ours_g and t_cublas_g:
                    row["graph_tflops"] = 2.0 * m ** 3 / t_ours_g * 1e-9
                    row["cublas_graph_tflops"] = 2.0 * m ** 3 / t_cublas_g * 1e-9
                if graph_err:
                    row["graph_err"] = graph_err
                if ref_h is not None:
                    best = None
                    for st in (2, 3, 4):
                        stride = ref_swizzle_stride(m)
                        stride = stride if stride in (512, 1024, 2048, 4096) and not (st == 4 and stride == 512) else 2048
                        rc = ref_h.ref_hgemm_mma_stages_dsmem_nn(A.data_ptr(), B.data_ptr(), C.data_ptr(), m, m, m, st, stride)
                        torch.cuda.synchronize()
                        if rc != 0:
                            continue
                        t = cuda_time(lambda: ref_h.ref_hgemm_mma_stages_dsmem_nn(A.data_ptr(), B.data_ptr(), C.data_ptr(), m, m, m, st, stride),
                                      5 if m > 8192 else 10, 2)
                        tf = 2.0 * m ** 3 / t * 1e-9
                        if best is None or tf > best[0]:
                            best = (tf, st, stride)
                    if best:
                        row["ref_mma_tflops"] = best[0]
                        row["ref_mma_cfg"] = "hgemm_mma_m16n8k16_mma2x4_warp4x4x2_stages_dsmem stages=%d swizzle_stride=%d" % best[1:]
                sweep.append(row)
                del A, B, C
            out["sweep"] = sweep
        del a, b, c
        torch.cuda.empty_cache()

        # -------------------------------------------------------------- attention (single GPU configs)
        ref_fa = load_ref_module("ref_flash_attn_lib")
        att = {}

        def bench_attn(shape, fn, tag):
            B_, H_, N_, D_ = shape
            q, k, v = [torch.randn(B_, H_, N_, D_, dtype=torch.half, device=dev) for _ in range(3)]
            o = torch.empty_like(q)
            fl = 4.0 * B_ * H_ * N_ * N_ * D_
            t = cuda_time(lambda: fn(q, k, v, o), 10, 3)
            r = {"shape": list(shape), "ms": t, "tflops": fl / t * 1e-9, "frac_of_peak_burst": fl / t * 1e-9 / peaks["tflops_burst"]}
            rooflines[tag] = roofline_obj(("fa2_fwd_tcgen05_kernel" if D_ <= 128 else "ffpa3_fwd_tcgen05_kernel<2>") + " " + tag, fl, t, peaks,
                                          tag, 4.0 * B_ * H_ * N_ * D_ * 2, t * 10)
            if ref_fa is not None and D_ <= 128:
                try:
                    o2 = torch.zeros_like(q)
                    t2 = cuda_time(lambda: ref_fa.flash_attn_mma_stages_split_q_shared_qkv(q, k, v, o2, 2), 3, 1)
                    r["ref_mma_share_qkv_stage2_tflops"] = fl / t2 * 1e-9
                    r["max_abs_diff_vs_ref"] = float((o.float() - o2.float()).abs().max().item())
                except Exception as e:  # noqa
                    r["ref_err"] = repr(e)[:160]
            try:
                import torch.nn.functional as F
                t3 = cuda_time(lambda: F.scaled_dot_product_attention(q, k, v), 3, 1)
                r["sdpa_tflops"] = fl / t3 * 1e-9
            except Exception as e:  # noqa
                r["sdpa_err"] = repr(e)[:160]
            return r

        if want("attention"):
            att["cfg3_fa2_b4_h48_n8192_d64"] = bench_attn(FA2_CFG3, ops.fa2_fwd, "fa2_cfg3_d64")
            launches += 13
            att["cfg5_shard_b4_h64_n8192_d128"] = bench_attn((4, 64, 8192, 128), ops.fa2_fwd, "fa2_cfg5_shard_d128")
            launches += 13
            out["attention"] = att
        if want("ffpa"):
            ffpa = bench_attn(FFPA_CFG4, ops.ffpa_fwd, "ffpa_cfg4_d512")
            launches += 13
            ref_ffpa = load_ref_module("pyffpa_cuda")
            if ref_ffpa is not None:
                try:
                    B_, H_, N_, D_ = FFPA_CFG4
                    q, k, v = [torch.randn(B_, H_, N_, D_, dtype=torch.half, device=dev) for _ in range(3)]
                    o2 = torch.zeros_like(q)
                    fl = 4.0 * B_ * H_ * N_ * N_ * D_
                    for name in ("ffpa_mma_acc_f32_L1", "ffpa_mma_acc_f16_L1"):
                        best = 0.0
                        for st in (1, 2, 3, 4):
                            t2 = cuda_time(lambda: getattr(ref_ffpa, name)(q, k, v, o2, st), 3, 1)
                            best = max(best, fl / t2 * 1e-9)
                        ffpa["ref_" + name + "_tflops"] = best
                except Exception as e:  # noqa
                    ffpa["ref_err"] = repr(e)[:160]
            out["ffpa"] = {"cfg4_b1_h32_n4096_d512": ffpa}
            torch.cuda.empty_cache()

        # -------------------------------------------------------------- the widened rows, N = 1 only
        if want("widened"):
            if world == 1:
                try:
                    gd = {}
                    for dt, nm in ((torch.bfloat16, "bf16"), (torch.float32, "tf32")):
                        A = torch.randn(n, n, device=dev).to(dt)
                        Bm = torch.randn(n, n, device=dev).to(dt)
                        C = torch.empty(n, n, device=dev).to(dt)
                        t_o = cuda_time(lambda: ops.gemm(A, Bm, C), 10, 3)
                        launches += 13
                        prev = torch.backends.cuda.matmul.allow_tf32
                        torch.backends.cuda.matmul.allow_tf32 = True
                        t_c = cuda_time(lambda: torch.matmul(A, Bm, out=C), 10, 3)
                        torch.backends.cuda.matmul.allow_tf32 = prev
                        gd[nm + "_8192"] = {"tflops": flops / t_o * 1e-9, "cublas_tflops": flops / t_c * 1e-9}
                        del A, Bm, C
                    out["gemm_dtypes"] = gd
                except Exception as e:  # noqa
                    out["gemm_dtypes"] = {"error": repr(e)[:200]}
                try:
                    hbm = peaks.get("hbm_gbs")
                    ne = 64 * 1024 * 1024
                    xa, xb = torch.randn(ne, device=dev), torch.randn(ne, device=dev)
                    xc = torch.empty_like(xa)
                    xr = torch.randn(16384, 8192, dtype=torch.half, device=dev)
                    yr = torch.empty_like(xr)
                    sup = {}
                    for nm, fn, nbytes in (
                            ("elementwise_add_f32", lambda: ops.elementwise_add(xa, xb, xc), 3 * ne * 4),
                            ("block_all_reduce_sum_f32", lambda: ops.block_all_reduce_sum(xa), ne * 4),
                            ("safe_softmax_f16_h8192", lambda: ops.softmax(xr, yr, ops.SOFTMAX_SAFE), 2 * xr.numel() * 2),
                            ("rms_norm_f16_k8192", lambda: ops.rms_norm(xr, yr, 1.0), 2 * xr.numel() * 2),
                            ("layer_norm_f16_k8192", lambda: ops.layer_norm(xr, yr, 1.0, 0.0), 2 * xr.numel() * 2),
                            ("gelu_f32", lambda: ops.activation(xa, xc, "gelu"), 2 * ne * 4),
                            ("dot_prod_f32", lambda: ops.dot_prod(xa, xb), 2 * ne * 4)):
                        t_s = cuda_time(fn, 10, 3)
                        launches += 13
                        sup[nm] = {"gbps": nbytes / t_s * 1e-6, "frac_of_measured_hbm_peak": (nbytes / t_s * 1e-6 / hbm) if hbm else None}
                    out["support_hbm"] = sup
                    del xa, xb, xc, xr, yr
                except Exception as e:  # noqa
                    out["support_hbm"] = {"error": repr(e)[:200]}
                torch.cuda.empty_cache()

        # -------------------------------------------------------------- config #5: batch-sharded attention over the ranks
        if want("sharded"):
            sharded_rec = bench_sharded_cfg5(world, rank, dev, barrier, dist_on, peaks)
            launches += sharded_rec.pop("_launches", 0)
            e2e["sharded_attention"] = sharded_rec

    # ------------------------------------------------------------------ CPU baseline (rank 0, N = 1 only)
    cpu_baseline = None
    config1 = None
    if rank == 0 and world == 1:
        v, sample = cpu_hgemm_sample(n, budget_s=12.0)
        info = cpu_info()
        cpu_baseline = {"value": v, "unit": "TFLOP/s", "cores": info["cores"], "kind": "port", "sample": sample,
                        "cpu": info["model"]}
        config1 = cpu_sgemm_config1()

    if rank == 0:
        line = {"metric": "hgemm_tflops", "value": value, "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_max, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f16 (fp32 accumulate)", "data": "synthetic",
                "config": headline_config(),
                "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "rooflines": rooflines, "cpu_baseline": cpu_baseline,
                "config1_sgemm_cpu": config1,
                "clocks": clocks, "peaks": peaks,
                "comm": ({"backend": "nccl", "nranks": world, "nccl_version": ".".join(map(str, torch.cuda.nccl.version()))}
                         if dist_on else None)}
        line.update(out)
        print(json.dumps(line), flush=True)
    if dist_on:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
        if nccl_log and os.path.exists(nccl_log):
            try:
                keep = [ln for ln in open(nccl_log, errors="replace") if ("nranks" in ln or "Init COMPLETE" in ln
                                                                          or "NCCL version" in ln or "NVLS" in ln)]
                sys.stderr.write("".join(keep[:40]))
                sys.stderr.flush()
            except Exception:
                pass


if __name__ == "__main__":
    main()
