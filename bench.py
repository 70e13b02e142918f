#!/usr/bin/env python
"""bench.py -- 8-layer RVQ frames/s (S0-S3 + A0-A3) on synthetic 768-d features, BASELINE.json's metric.

Workload (BASELINE.json configs[1]): one synthetic 1 h clip = 270 000 frames x 768-d fp32 features (75 Hz),
4 + 4 RVQ layers, codebook 1024, per GPU. A *step* is one pass of the hot path over that batch: the semantic
stack and the acoustic stack both encode the same resident frames into 8 int16 index streams; with N > 1 GPUs every
rank encodes its own 270 000 frames (weak scaling, replicated codebooks) and the index streams are all-gathered.

  value     whole-job frames/s, inputs resident in HBM, CUDA events on torch's current stream, max over ranks
  e2e       same metric through the public API with HOST buffers (`ResidualVectorQuantizer.encode_host`):
            pinned H2D of the features and D2H of the int16 index streams inside the timed region
  roofline  dominant kernel = tcgen05 distance GEMM + top-4 epilogue; algorithmic work 2*K*D flop per frame-layer;
            duration measured live with CUDA events around each launch (nat_rvq_encode_profile_f32)
  cpu_baseline  the oracle port of the reference's CPU path (torch.cdist + argmin chain) on a bounded sample

`--impl reference` times that CPU path alone (the reference is pure Python and cannot travel to the GPU box;
oracle/rvq_oracle.py is its restatement, bit-identical here, see tests/test_oracle_vs_reference.py).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

FRAMES = 270_000
DIM = 768
CODEBOOK = 1024
LAYERS_PER_STACK = 4
CPU_BASELINE_FRAMES = 270_000      # cpu_baseline leg: the whole workload once, about 10 s on 16 host threads
CPU_SAMPLE_FRAMES = 131_072        # --impl reference: one step = this many frames of the same clip (about 4 s)


def ncu_traffic_bytes():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
    `ncu --set full` capture of this workload (profiles/stack_traffic.json; bench.py itself never runs under ncu)."""
    path = os.path.join(ROOT, "profiles", "stack_traffic.json")
    try:
        with open(path) as f:
            t = json.load(f)
        return float(t["dram_bytes_read"]) + float(t["dram_bytes_write"])
    except (OSError, KeyError, ValueError):
        return None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"bf16_tflops": p.get("bf16_tflops", 1590.0), "bf16_tflops_sustained": p.get("bf16_tflops_sustained", 1400.0),
                "hbm_gbs": p.get("hbm_gbs", 6650.0), "source": "measured"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def make_stacks(device):
    """Two 4-layer stacks with the reference's seeded construction order."""
    from neural_audio_tokenizer_b200 import ResidualVectorQuantizer
    torch.manual_seed(42)
    stacks = [ResidualVectorQuantizer(DIM, CODEBOOK, LAYERS_PER_STACK, use_stochastic=False).eval() for _ in range(2)]
    return [s.to(device) for s in stacks]


def cpu_reference_rate(frames: int, steps: int = 1, warmup: int = 0):
    """frames/s of the CPU restatement of the reference path (8 layers, argmin mode), all host threads."""
    from oracle import rvq_oracle
    torch.manual_seed(42)
    cbs = [[torch.randn(CODEBOOK, DIM) for _ in range(LAYERS_PER_STACK)] for _ in range(2)]
    x = torch.randn(1, DIM, frames, generator=torch.Generator().manual_seed(1234))
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        for stack in cbs:
            rvq_oracle.rvq_forward(x, stack)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    total = sum(times)
    return frames * len(times) / total, total / len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = torch.get_num_threads()
    rate, sec = cpu_reference_rate(CPU_SAMPLE_FRAMES, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": "rvq_frames_per_sec_8_layers", "value": rate, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{CPU_SAMPLE_FRAMES} frames x {DIM}-d sample of the 270000-frame clip, 4+4 RVQ layers, "
                               f"codebook {CODEBOOK}, argmin mode, CPU"},
        "cpu_baseline": {"value": rate, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": f"{CPU_SAMPLE_FRAMES} frames per step (oracle/rvq_oracle.py: torch.cdist + argmin chain "
                                   f"of nat.py:1358-1420, {cores} threads of {os.cpu_count()} logical CPUs)"},
        "e2e": {"value": rate, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=FRAMES)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--kernel-only", action="store_true",
                    help="timed device loop only (no e2e / profile / CPU legs): the command ncu wraps")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)
    # stdout carries exactly one JSON line: whatever libraries print there meanwhile (NCCL's version banner) goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(obj), flush=True)
        os.dup2(2, 1)

    import torch.distributed as dist
    from neural_audio_tokenizer_b200 import _lib
    from neural_audio_tokenizer_b200.sharding import all_gather_codes, shard_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU path")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"        # keep stdout to the one JSON line (NCCL prints its version there)
        dist.init_process_group("nccl", device_id=device)
    lib = _lib.load()

    n_local = args.frames                     # weak scaling: every GPU encodes its own 270k-frame clip
    n_total = n_local * world
    stacks = make_stacks(device)
    gen = torch.Generator(device=device).manual_seed(1234 + rank)
    x = torch.randn(1, DIM, n_local, device=device, generator=gen)
    L_total = 2 * LAYERS_PER_STACK
    codes = torch.empty((L_total, n_local), dtype=torch.int16, device=device)
    handles = [s._pack.get(s._codebooks()) for s in stacks]
    ws_bytes = lib.nat_rvq_workspace_bytes(handles[0], n_local)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=device)
    stream = torch.cuda.current_stream(device)

    def step():
        for i, h in enumerate(handles):
            out = codes[i * LAYERS_PER_STACK:(i + 1) * LAYERS_PER_STACK]
            _lib.check(lib.nat_rvq_encode_f32(h, x.data_ptr(), _lib.LAYOUT_BCT, 1, n_local, out.data_ptr(),
                                              _lib.CODES_I16, None, None, 0.25, None, ws.data_ptr(), ws_bytes, 0,
                                              stream.cuda_stream))
        if world > 1:
            return all_gather_codes(codes, n_total)
        return codes

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    launches0 = lib.nat_launch_count()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        barrier()
        start.record()
        for _ in range(args.steps):
            step()
        stop.record()
        barrier()
    ms = start.elapsed_time(stop)
    launches = lib.nat_launch_count() - launches0
    if world > 1:
        t = torch.tensor([ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    ms_per_step = ms / args.steps
    value = n_total / (ms_per_step * 1e-3)

    if args.kernel_only:
        if rank == 0:
            emit({"metric": "rvq_frames_per_sec_8_layers", "value": value, "unit": "frames/s",
                  "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
                  "gpu_launches": int(launches), "kernel_only": True})
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- end to end through the public API with host buffers -------------------------------------------------
    xh = x.cpu().pin_memory()
    host_out = [torch.empty((LAYERS_PER_STACK, 1, n_local), dtype=torch.int16, pin_memory=True) for _ in stacks]

    def e2e_step():
        for s, o in zip(stacks, host_out):
            s.encode_host(xh, code_dtype=torch.int16, out=o)      # returns after the D2H of the codes has landed

    e2e_steps = max(2, min(args.steps, 5))
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    if world > 1:
        t = torch.tensor([e2e_s], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = t.item()
    e2e_value = n_total / e2e_s
    same = all(torch.equal(o[:, 0].cpu(), codes[i * LAYERS_PER_STACK:(i + 1) * LAYERS_PER_STACK].cpu())
               for i, o in enumerate(host_out))

    # ---- roofline of the dominant kernel, measured live -------------------------------------------------------
    prof = (ctypes.c_float * _lib.PROF_FIELDS)()
    prof_sum = [0.0] * _lib.PROF_FIELDS
    prof_reps = 3
    for _ in range(prof_reps):
        for i, h in enumerate(handles):
            out = codes[i * LAYERS_PER_STACK:(i + 1) * LAYERS_PER_STACK]
            # single-stream so that every launch is timed alone and covers all n_local frames
            _lib.check(lib.nat_rvq_encode_profile_f32(h, x.data_ptr(), _lib.LAYOUT_BCT, 1, n_local, out.data_ptr(),
                                                      _lib.CODES_I16, None, None, 0.25, None, ws.data_ptr(), ws_bytes,
                                                      _lib.RVQ_SINGLE_STREAM, stream.cuda_stream, prof))
            for k in range(_lib.PROF_FIELDS):
                prof_sum[k] += prof[k]
    gemm_launches = prof_sum[6]
    gemm_ms_per_launch = prof_sum[1] / max(gemm_launches, 1)
    fused = os.environ.get("NAT_RVQ_FUSED", "1") != "0"
    # fused: one launch = all 4 layers of a stack; per-layer kernels: one launch = one layer
    flops_per_launch = 2.0 * CODEBOOK * DIM * n_local * (LAYERS_PER_STACK if fused else 1)
    achieved = flops_per_launch / (gemm_ms_per_launch * 1e-3) / 1e12
    peaks = measured_peaks()
    kernel_ms = {name: prof_sum[k] / prof_reps for k, name in enumerate(_lib.PROF_NAMES) if k < 6}
    stats = torch.zeros((LAYERS_PER_STACK, _lib.STAT_FIELDS), dtype=torch.int64, device=device)
    _lib.check(lib.nat_rvq_encode_f32(handles[0], x.data_ptr(), _lib.LAYOUT_BCT, 1, n_local, codes.data_ptr(),
                                      _lib.CODES_I16, None, None, 0.25, stats.data_ptr(), ws.data_ptr(), ws_bytes, 0,
                                      stream.cuda_stream))
    torch.cuda.synchronize()
    st = stats.cpu().tolist()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    line = {
        "metric": "rvq_frames_per_sec_8_layers", "value": value, "unit": "frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f16 tensor-core coarse pass (fp32 accumulate) + f64 exact re-rank; "
                                                          "indices bit-exact vs fp32 argmin",
        "data": "synthetic",
        "config": {"workload": f"synthetic 1 h clip per GPU: {n_local} frames x {DIM}-d fp32 features (75 Hz), 4+4 RVQ "
                               f"layers, codebook {CODEBOOK}, argmin mode, int16 index streams"
                               + (", NCCL all-gather of the index streams" if world > 1 else ""),
                   "frames_per_gpu": n_local, "dim": DIM, "codebook_size": CODEBOOK, "layers": L_total,
                   "l2": "inputs larger than L2 (0.83 GB features per GPU, 126 MB L2)",
                   "parallelism": f"frame-sharded x{world}, replicated codebooks"},
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": 2 * n_local * DIM * 4,
                "d2h_bytes_per_step": L_total * n_local * 2, "api": "ResidualVectorQuantizer.encode_host (pinned host "
                "tensors in, int16 host tensors out)", "steps": e2e_steps, "matches_device_path": bool(same)},
        "gpu_launches": int(launches),
        "clocks": clocks.summary(),
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                     "frac": achieved / peaks["bf16_tflops_sustained"],
                     "traffic": ncu_traffic_bytes() if (fused and n_local == FRAMES) else None,
                     "kernel": "rvq_stack_kernel (4 layers per launch: tcgen05 GEMM + candidates + exact decision + "
                               "residual update)" if fused else "rvq_gemm_topk_kernel", "peak_kind": f"{peaks['source']} sustained bf16 (kernel timed "
                     "inside the step)", "frac_of_burst": achieved / peaks["bf16_tflops"],
                     "ms_per_launch": gemm_ms_per_launch, "flop_per_launch": flops_per_launch,
                     "kernel_ms_per_step": kernel_ms},
        "decision_stats_semantic_stack": {"certified": [r[0] for r in st], "reranked": [r[1] for r in st],
                                          "full_scan": [r[2] for r in st]},
    }
    if not args.no_cpu_baseline and world == 1:
        cores = torch.get_num_threads()
        rate, sec = cpu_reference_rate(min(CPU_BASELINE_FRAMES, n_local), steps=1, warmup=0)
        line["cpu_baseline"] = {"value": rate, "unit": "frames/s", "cores": cores, "kind": "port",
                                "sample": f"{min(CPU_BASELINE_FRAMES, n_local)} frames of the same workload (all of it), one "
                                          f"pass, {sec:.1f} s (oracle/rvq_oracle.py, torch CPU, {cores} threads of "
                                          f"{os.cpu_count()} logical CPUs)"}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
