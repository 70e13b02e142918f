#!/usr/bin/env python
"""bench.py -- 8-layer RVQ frames/s (S0-S3 + A0-A3) on synthetic 768-d features, BASELINE.json's metric.

Workload (BASELINE.json configs[1]): one synthetic 1 h clip = 270 000 frames x 768-d fp32 features (75 Hz),
4 + 4 RVQ layers, codebook 1024, per GPU. A *step* is one pass of the hot path over that batch: the semantic
stack and the acoustic stack both encode the same resident frames into 8 int16 index streams (one native call,
`encode_stacks`: one layer-0 preparation, one persistent launch); with N > 1 GPUs every rank encodes its own
270 000 frames (weak scaling, replicated codebooks) and the index streams are all-gathered over NCCL.

  value         whole-job frames/s, inputs resident in HBM, CUDA events on torch's current stream, max over ranks
  e2e           same metric through the public API with HOST buffers (`encode_stacks_host`): pinned H2D of the
                features (once per chunk, both stacks run on it) and D2H of the int16 index streams inside the timed region
  parity        the device's index streams against the oracle's on the SAME inputs (all 270 000 frames at N = 1, a
                stated sample otherwise): exact frames, near-tie flips (relative distance gap < 1e-6), real mismatches
                (the run fails when there is one), cascade tokens
  forward_form  frames/s of `rvq(x)` with the quantised sum and the losses, the form install() puts into the
                reference's forward (nat.py:3239-3240)
  roofline      dominant kernel = the fused stack kernel (tcgen05 distance GEMM + candidates + exact decision + residual
                update); algorithmic work 2*K*D flop per frame-layer; duration measured live with CUDA events around
                each launch on the launch stream (nat_rvq_encode_stacks_profile_f32); peak = measured burst bf16
  frontend      mel / spectral kernels on one hour of audio (BASELINE config 4 geometry) and torchaudio on host cores
  cpu_baseline  the reference's CPU path (argmin mode) over the whole workload once on the box's host cores: the
                unmodified reference from baseline/_ref when it travelled (kind "reference"), else the oracle port

`--impl reference` times that CPU path alone, all host threads, on bounded samples of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes
import hashlib
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
CPU_SAMPLE_FRAMES = 131_072        # --impl reference: one step = this many frames of the same clip (about 4 s)
PARITY_SAMPLE_FRAMES = 32_768      # parity leg when N > 1 (rank 0's first frames)
STOCHASTIC_SAMPLE_FRAMES = 16_384  # default (sampling) mode of the reference on the host cores
FRONTEND_CPU_SECONDS = 300         # torchaudio baseline: 5 min of the 1 h waveform


def build_id():
    """sha256 over the kernel sources: ties a committed ncu capture to the build it was taken on."""
    h = hashlib.sha256()
    csrc = os.path.join(ROOT, "neural_audio_tokenizer_b200", "csrc")
    for name in sorted(os.listdir(csrc)):
        with open(os.path.join(csrc, name), "rb") as f:
            h.update(name.encode() + b"\0" + f.read())
    return h.hexdigest()[:16]


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel from the committed
    `ncu --set full` capture (profiles/stack_traffic.json; bench.py itself never runs under ncu). Reported only when
    the capture was taken on THIS build of the kernels; otherwise null plus the reason."""
    path = os.path.join(ROOT, "profiles", "stack_traffic.json")
    try:
        with open(path) as f:
            t = json.load(f)
    except (OSError, ValueError):
        return None, "no capture committed"
    if t.get("build_id") != build_id():
        return None, f"capture is of build {t.get('build_id')}, this is {build_id()}"
    try:
        return float(t["dram_bytes_read"]) + float(t["dram_bytes_write"]), t.get("source", path)
    except (KeyError, ValueError):
        return None, "capture file incomplete"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"bf16_tflops": p.get("bf16_tflops", 1590.0), "bf16_tflops_sustained": p.get("bf16_tflops_sustained", 1400.0),
                "hbm_gbs": p.get("hbm_gbs", 6650.0), "source": "measured"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 25 ms while the timed region runs (it lasts ~50 ms)."""

    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "25"],
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


# ------------------------------------------------------------------------------------------------ workload
def make_codebooks():
    """The reference's seeded construction order: torch.manual_seed(42), then the semantic
    and the acoustic stack, one randn(K, D) per layer (nat.py:2115)."""
    torch.manual_seed(42)
    return [[torch.randn(CODEBOOK, DIM) for _ in range(LAYERS_PER_STACK)] for _ in range(2)]


def make_features(frames: int, rank: int = 0):
    """[1, D, frames] fp32 on the HOST from the CPU generator: the same tensor feeds the device arm, the end-to-end
    arm and the CPU arm."""
    return torch.randn(1, DIM, frames, generator=torch.Generator().manual_seed(1234 + rank))


def make_stacks(device):
    from neural_audio_tokenizer_b200 import ResidualVectorQuantizer
    torch.manual_seed(42)
    stacks = [ResidualVectorQuantizer(DIM, CODEBOOK, LAYERS_PER_STACK, use_stochastic=False).eval() for _ in range(2)]
    return [s.to(device) for s in stacks]


def host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm uses the box's cores."""
    n = os.cpu_count() or 1
    torch.set_num_threads(n)
    return torch.get_num_threads()


def reference_stacks():
    """The UNMODIFIED reference's quantiser stacks (baseline/_ref via oracle/ref_shim) in argmin mode, or None."""
    try:
        from oracle import ref_shim
        if not ref_shim.reference_available():
            return None
        nat = ref_shim.load_reference()
    except Exception as e:      # the reference needs transformers / scipy to import; the port stands in without them
        sys.stderr.write(f"bench.py: reference not importable ({type(e).__name__}: {e}); using the oracle port\n")
        return None
    torch.manual_seed(42)
    stacks = [nat.ResidualVectorQuantizer(DIM, CODEBOOK, LAYERS_PER_STACK).eval() for _ in range(2)]
    for s in stacks:
        for q in s.quantizers:
            q.use_stochastic = False            # the argmin contract (nat.py:2155-2157)
    return stacks


def cpu_arm(x, steps: int = 1, warmup: int = 0, keep_codes: bool = False):
    """frames/s of the reference's CPU path (both stacks, argmin mode) on x [1, D, n]; optionally its codes [8, n]."""
    from oracle import rvq_oracle
    ref = reference_stacks()
    cbs = make_codebooks()
    if ref is not None:
        for s, c in zip(ref, cbs):
            assert all(torch.equal(q.codebook, t) for q, t in zip(s.quantizers, c)), "seeded codebooks drifted"
    times, codes = [], None
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        outs = []
        with torch.no_grad():
            for k in range(2):
                outs.append(ref[k](x)[1] if ref is not None else rvq_oracle.rvq_forward(x, cbs[k])[1])
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        if keep_codes:
            codes = torch.stack([c.reshape(-1) for o in outs for c in o]).numpy()
    total = sum(times)
    kind = "reference" if ref is not None else "port"
    return x.shape[-1] * len(times) / total, total / len(times), kind, codes


def reference_on_gpu_block(x_dev):
    """The UNMODIFIED reference's own RVQ forward (torch ops: cdist, argmin, embedding) on the SAME GPU and inputs:
    what moving the reference to cuda buys without this library. Reported beside the CPU arm, never as the baseline."""
    ref = reference_stacks()
    if ref is None:
        return {"unavailable": "reference not carried (baseline/_ref)"}
    ref = [s.to(x_dev.device) for s in ref]
    def step():
        with torch.no_grad():
            for s in ref:
                s(x_dev)
    try:
        step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            step()
        e1.record()
        torch.cuda.synchronize()
    except torch.cuda.OutOfMemoryError as e:
        return {"unavailable": f"out of memory: {str(e)[:80]}"}
    ms = e0.elapsed_time(e1) / 3
    return {"value": x_dev.shape[-1] / (ms * 1e-3), "unit": "frames/s", "ms_per_step": ms,
            "what": "the unmodified reference's ResidualVectorQuantizer.forward (argmin mode, both stacks, quantised sum "
                    "and losses: compare forward_form) as torch eager ops on this GPU, same inputs"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_threads()
    x = make_features(FRAMES)[:, :, :CPU_SAMPLE_FRAMES].contiguous()
    rate, sec, kind, _ = cpu_arm(x, steps=args.steps, warmup=args.warmup)
    what = ("the unmodified reference (baseline/_ref/neural_audio_tokenizer.py, ResidualVectorQuantizer.forward, "
            "use_stochastic=False)" if kind == "reference" else "oracle/rvq_oracle.py (torch.cdist + argmin chain of nat.py:1358-1420)")
    line = {
        "impl": "reference", "metric": "rvq_frames_per_sec_8_layers", "value": rate, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"synthetic 1 h clip: {FRAMES} frames x {DIM}-d fp32 features (75 Hz), 4+4 RVQ layers, "
                               f"codebook {CODEBOOK}, argmin mode; each step a {CPU_SAMPLE_FRAMES}-frame sample of it "
                               "(frames/s is intensive: the rate does not depend on the sample length)",
                   "frames_per_step": CPU_SAMPLE_FRAMES, "dim": DIM, "codebook_size": CODEBOOK, "layers": 2 * LAYERS_PER_STACK},
        "cpu_baseline": {"value": rate, "unit": "frames/s", "cores": cores, "kind": kind,
                         "sample": f"{CPU_SAMPLE_FRAMES} frames per step through {what}, {cores} threads of "
                                   f"{os.cpu_count()} logical CPUs"},
        "e2e": {"value": rate, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ legs
def parity_block(x_host, codes_dev, n):
    """Device codes [8, >= n] against the oracle on the first n frames of the same host tensor."""
    from oracle import rvq_oracle
    x = x_host[:, :, :n].contiguous()
    t0 = time.perf_counter()
    rate, sec, kind, ref = cpu_arm(x, keep_codes=True)
    got = codes_dev[:, :n].cpu().numpy()
    cbs = make_codebooks()
    x_nd = x[0].numpy().T                                   # [n, D] view, rows fetched only for mismatching frames
    out = {"frames": n, "layers": 2 * LAYERS_PER_STACK, "exact_frames": 0, "near_tie_flips": 0, "real_mismatches": 0,
           "cascade_tokens": 0, "flips": []}
    exact = None
    for k in range(2):
        sl = slice(k * LAYERS_PER_STACK, (k + 1) * LAYERS_PER_STACK)
        rep = rvq_oracle.classify_mismatches(x_nd, [c.numpy() for c in cbs[k]], ref[sl], got[sl])
        for key in ("near_tie_flips", "real_mismatches", "cascade_tokens"):
            out[key] += rep[key]
        out["flips"] += [dict(f, stack="SA"[k]) for f in rep["flips"][:8]]
        ok = (ref[sl] == got[sl]).all(axis=0)
        exact = ok if exact is None else (exact & ok)
    out["exact_frames"] = int(exact.sum())
    out["exact_match_fraction"] = out["exact_frames"] / n
    out["checker"] = ("unmodified reference (baseline/_ref)" if kind == "reference" else "oracle port") + \
                     ", same host tensor as the device arm, near-tie = relative distance gap < 1e-6 (fp64)"
    out["seconds"] = time.perf_counter() - t0
    return out, rate, sec, kind


def frontend_block(device, peaks):
    """Mel and spectral kernels on one hour of audio at BASELINE config 4's geometry (24 kHz, hop 320, n_fft 2048),
    and torchaudio's MelSpectrogram on the host cores over a bounded sample of the same waveform."""
    from neural_audio_tokenizer_b200 import MelSpectrogram, spectral_stats
    sr, hop, secs = 24000, 320, 3600
    S = sr * secs
    wave = torch.randn(1, S, device=device, generator=torch.Generator(device=device).manual_seed(7)) * 0.1
    mt = MelSpectrogram(sample_rate=sr, n_fft=2048, hop_length=hop, n_mels=128).to(device)

    def timeit(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    mel_ms = timeit(lambda: mt(wave))
    T = 1 + S // hop
    alg_bytes = T * (hop * 4 + 128 * 4)
    flop = T * 0.5 * 5 * 2048 * 11                       # one complex 2048-point FFT serves two real frames
    spec_ms = timeit(lambda: spectral_stats(wave[0], sr, 2048, hop))
    Ts = 1 + (S - 2048) // hop
    out = {"geometry": f"{secs} s at {sr} Hz, n_fft 2048, hop {hop}, 128 mels", "frames": T,
           "mel_ms": mel_ms, "frames_per_s": T / (mel_ms * 1e-3), "algorithmic_gb_per_s": alg_bytes / mel_ms / 1e6,
           "hbm_frac": alg_bytes / (mel_ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "fp32_tflops": flop / mel_ms / 1e9,
           "bound": "fp32 / shared memory (FFT), not HBM: SURVEY.md section 7",
           "spectral_ms": spec_ms, "spectral_frames_per_s": Ts / (spec_ms * 1e-3)}
    try:
        import torchaudio
        cpu_wave = wave[:, :sr * FRONTEND_CPU_SECONDS].cpu()
        ref = torchaudio.transforms.MelSpectrogram(sample_rate=sr, n_fft=2048, hop_length=hop, n_mels=128, normalized=True)
        ref(cpu_wave[:, :sr])
        t0 = time.perf_counter()
        m = ref(cpu_wave)
        dt = time.perf_counter() - t0
        out["cpu_torchaudio_frames_per_s"] = m.shape[-1] / dt
        out["cpu_sample"] = f"{FRONTEND_CPU_SECONDS} s of the same waveform, torchaudio {torchaudio.__version__}, {torch.get_num_threads()} threads"
        dev = mt(wave[:, :sr * FRONTEND_CPU_SECONDS])[0].cpu()
        out["max_abs_err_over_max"] = float((dev - m[0]).abs().max() / m.abs().max())
    except Exception as e:
        out["cpu_torchaudio_frames_per_s"] = None
        out["cpu_sample"] = f"torchaudio unavailable: {type(e).__name__}"
    del wave
    torch.cuda.empty_cache()
    return out


def stochastic_cpu_block(x_host):
    """The reference's DEFAULT mode (sampling, nat.py:2150-2154) on the host cores, bounded sample, both stacks."""
    from oracle import rvq_oracle
    x = x_host[:, :, :STOCHASTIC_SAMPLE_FRAMES].contiguous()
    cbs = make_codebooks()
    torch.manual_seed(0)
    t0 = time.perf_counter()
    for k in range(2):
        rvq_oracle.rvq_forward_sampling(x, cbs[k], [0.5] * LAYERS_PER_STACK)
    dt = time.perf_counter() - t0
    return {"value": STOCHASTIC_SAMPLE_FRAMES / dt, "unit": "frames/s", "kind": "port",
            "sample": f"{STOCHASTIC_SAMPLE_FRAMES} frames, 8 layers, temperature 0.5 (multinomial = exponential_ + argmax)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=FRAMES)
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU legs (parity included)")
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
    from neural_audio_tokenizer_b200 import _lib, encode_stacks, encode_stacks_host, HostContext
    from neural_audio_tokenizer_b200.sharding import CodeGatherer

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
    x_host = make_features(n_local, rank).pin_memory()
    x = x_host.to(device, non_blocking=True)
    L_total = 2 * LAYERS_PER_STACK
    codes = torch.empty((L_total, 1, n_local), dtype=torch.int16, device=device)
    handles = [s._pack.get(s._codebooks()) for s in stacks]
    harr = (ctypes.c_void_p * 2)(*[h.value for h in handles])
    ws_bytes = lib.nat_rvq_stacks_workspace_bytes(harr, 2, n_local)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=device)
    stream = torch.cuda.current_stream(device)
    # index all-gather: NCCL's all_gather_into_tensor on a side stream; NAT_BENCH_GATHER=peer uses the copy engines
    # over NVLink peer memory instead (no kernel beside the persistent stack kernel: 466 M frames/s on 8 GPUs against
    # 447 M, profiles/r2_v4_bench_8gpu_peer.json). NCCL stays the default because the peer path's two-phase teardown
    # was written after the round's GPU budget was spent (its first version hung torchrun jobs at exit).
    gather_kind = os.environ.get("NAT_BENCH_GATHER", "nccl") if world > 1 else None
    gatherer, gather_note = None, None
    if gather_kind == "peer":
        from neural_audio_tokenizer_b200.sharding import PeerCodeGatherer
        try:
            gatherer = PeerCodeGatherer(L_total, n_local, world, device)
        except RuntimeError as e:                # raised on every rank alike (the set-up agrees before it returns)
            gather_kind, gather_note = "nccl", f"peer-memory exchange refused: {e}"
            sys.stderr.write(f"bench.py: {gather_note}; using NCCL\n")
    if world > 1 and gatherer is None:
        gatherer = CodeGatherer(L_total, n_local, world, device)

    def step():
        encode_stacks(stacks, x, torch.int16, out=codes, workspace=ws)
        if gatherer is not None:
            return gatherer.all_gather(codes[:, 0])
        return codes

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        gathered = step()
    barrier()
    launches0 = lib.nat_launch_count()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        barrier()
        start.record()
        for _ in range(args.steps):
            gathered = step()
        if gatherer is not None:
            gatherer.wait()                     # the last step's exchange is part of the timed region
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
                  "gpu_launches": int(launches), "kernel_only": True, "build_id": build_id()})
        if world > 1:
            if hasattr(gatherer, "close"):
                gatherer.close()
            dist.destroy_process_group()
        return

    gather_ok = None
    if world > 1:
        # every rank's shard sits where shard_range puts it, and the whole equals the sum of the parts
        mine = gathered[:, rank * n_local:(rank + 1) * n_local]
        local_sum = codes[:, 0].long().sum()
        total = local_sum.clone()
        dist.all_reduce(total)
        gather_ok = bool(torch.equal(mine, codes[:, 0])) and int(gathered.long().sum()) == int(total)

    # ---- end to end through the public API with host buffers -------------------------------------------------
    host_out = torch.empty((L_total, 1, n_local), dtype=torch.int16, pin_memory=True)
    ctx = HostContext(device)

    def e2e_step():
        encode_stacks_host(stacks, x_host, torch.int16, out=host_out, ctx=ctx)   # returns once the codes have landed

    e2e_steps = max(2, min(args.steps, 5))
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    h2d_gbs_rank = n_local * DIM * 4 / e2e_s / 1e9
    if world > 1:
        t = torch.tensor([e2e_s], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = t.item()
    e2e_value = n_total / e2e_s
    same = bool(torch.equal(host_out, codes.cpu()))
    ctx.close()

    # ---- roofline of the dominant kernel, measured live -------------------------------------------------------
    prof = (ctypes.c_float * _lib.PROF_FIELDS)()
    prof_sum = [0.0] * _lib.PROF_FIELDS
    prof_reps = 5
    xarr = (ctypes.c_void_p * 2)(x.data_ptr(), x.data_ptr())
    for _ in range(prof_reps):
        _lib.check(lib.nat_rvq_encode_stacks_profile_f32(harr, 2, xarr, _lib.LAYOUT_BCT, 1, n_local, codes.data_ptr(),
                                                         _lib.CODES_I16, ws.data_ptr(), ws_bytes, 0, stream.cuda_stream, prof))
        for k in range(_lib.PROF_FIELDS):
            prof_sum[k] += prof[k]
    gemm_launches = prof_sum[6]
    gemm_ms_per_launch = prof_sum[1] / max(gemm_launches, 1)
    layers_per_launch = L_total * prof_reps / max(gemm_launches, 1)       # 8: both stacks in one launch; 4 / 1: fallbacks
    flops_per_launch = 2.0 * CODEBOOK * DIM * n_local * layers_per_launch
    achieved = flops_per_launch / (gemm_ms_per_launch * 1e-3) / 1e12
    peaks = measured_peaks()
    kernel_ms = {name: prof_sum[k] / prof_reps for k, name in enumerate(_lib.PROF_NAMES) if k < 6}
    traffic, traffic_note = ncu_traffic() if n_local == FRAMES else (None, "capture is of the 270 000-frame workload")

    # ---- the form install() puts into the reference's forward: quantised sum + losses --------------------------
    def forward_step():
        for s in stacks:
            s(x)
    with torch.no_grad():
        forward_step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            forward_step()
        e1.record()
        torch.cuda.synchronize()
    fwd_ms = e0.elapsed_time(e1) / 3

    # ---- the reference's DEFAULT selection mode (sampling, nat.py:2150-2154) with device-side Philox noise -------
    def sampling_step():
        for s in stacks:
            for q in s.quantizers:
                q.use_stochastic = True
            s.sampling_mode = "philox"
            s.encode(x)
    with torch.no_grad():
        try:
            for _ in range(2):
                sampling_step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                sampling_step()
            e1.record()
            torch.cuda.synchronize()
            smp_ms = e0.elapsed_time(e1) / 3
        finally:
            for s in stacks:
                for q in s.quantizers:
                    q.use_stochastic = False

    stats = torch.zeros((LAYERS_PER_STACK, _lib.STAT_FIELDS), dtype=torch.int64, device=device)
    ws1 = torch.empty(lib.nat_rvq_workspace_bytes(handles[0], n_local), dtype=torch.uint8, device=device)
    scratch = torch.empty((LAYERS_PER_STACK, n_local), dtype=torch.int16, device=device)
    _lib.check(lib.nat_rvq_encode_f32(handles[0], x.data_ptr(), _lib.LAYOUT_BCT, 1, n_local, scratch.data_ptr(),
                                      _lib.CODES_I16, None, None, 0.25, stats.data_ptr(), ws1.data_ptr(), ws1.numel(), 0,
                                      stream.cuda_stream))
    torch.cuda.synchronize()
    st = stats.cpu().tolist()
    single_matches = bool(torch.equal(scratch, codes[:LAYERS_PER_STACK, 0]))
    del ws1, scratch

    if rank != 0:
        if world > 1:
            dist.barrier()                    # rank 0's CPU legs run while the others wait here
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
                               + (", all-gather of the index streams" if world > 1 else ""),
                   "frames_per_gpu": n_local, "dim": DIM, "codebook_size": CODEBOOK, "layers": L_total,
                   "l2": "inputs larger than L2 (0.83 GB features per GPU, 126 MB L2)",
                   "parallelism": f"frame-sharded x{world}, replicated codebooks", "build_id": build_id()},
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": n_local * DIM * 4,
                "d2h_bytes_per_step": L_total * n_local * 2, "api": "encode_stacks_host / nat_tokenize_host_f32 (pinned "
                "host tensor in, int16 host tensor out; each chunk uploaded once, both stacks run on it)",
                "steps": e2e_steps, "matches_device_path": same, "h2d_gb_per_s_this_rank": h2d_gbs_rank},
        "gpu_launches": int(launches),
        "clocks": clocks.summary(),
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                     "frac": achieved / peaks["bf16_tflops"], "traffic": traffic, "traffic_note": traffic_note,
                     "kernel": f"rvq_stack_kernel ({layers_per_launch:g} layers per launch: tcgen05 GEMM + candidates + "
                               "exact decision + residual update)",
                     "peak_kind": f"{peaks['source']} burst bf16 (kernel timed alone by CUDA events around its launch)",
                     "frac_of_sustained": achieved / peaks["bf16_tflops_sustained"],
                     "ms_per_launch": gemm_ms_per_launch, "flop_per_launch": flops_per_launch,
                     "kernel_ms_per_step": kernel_ms},
        "forward_form": {"value": n_local / (fwd_ms * 1e-3), "unit": "frames/s", "ms_per_step": fwd_ms,
                         "what": "rvq(x) for both stacks with the quantised sum and vq_loss (nat.py:3239-3240 as grafted "
                                 "by install()), this rank's frames"},
        "sampling_form": {"value": n_local / (smp_ms * 1e-3), "unit": "frames/s", "ms_per_step": smp_ms,
                          "what": "encode(x) of both stacks in the reference's default sampling mode (temperature 0.5), "
                                  "sampling_mode 'philox': device noise, equal to the reference in distribution only "
                                  "(tests/test_rvq_sampling_gpu.py); compare cpu_baseline_stochastic"},
        "decision_stats_semantic_stack": {"certified": [r[0] for r in st], "reranked": [r[1] for r in st],
                                          "full_scan": [r[2] for r in st],
                                          "reranked_in_fp64": [r[3] for r in st],
                                          "single_stack_call_matches_two_stack_launch": single_matches},
    }
    if gather_ok is not None:
        line["all_gather"] = {"ok": gather_ok, "bytes_per_rank": L_total * n_local * 2,
                              "transport": ("copy engines over NVLink peer memory (nat_peer_all_gather: 2-D peer copies into "
                                            "every rank's final layout, stream memory operations as the barrier; no kernel)"
                                            if gather_kind == "peer" else "NCCL all_gather_into_tensor + one strided copy"),
                              "overlap": "issued on a side stream per step; the next step's kernels do not wait for it"}
        if gather_note:
            line["all_gather"]["note"] = gather_note
    if not args.no_cpu_baseline:
        cores = host_threads()
        n_par = min(n_local, FRAMES) if world == 1 else min(n_local, PARITY_SAMPLE_FRAMES)
        par, rate, sec, kind = parity_block(x_host, codes[:, 0], n_par)
        line["parity"] = par
        if world == 1:
            line["cpu_baseline"] = {"value": rate, "unit": "frames/s", "cores": cores, "kind": kind,
                                    "sample": f"{n_par} frames of the same workload (all of it) on the same inputs, one "
                                              f"pass, {sec:.1f} s, {cores} threads of {os.cpu_count()} logical CPUs; its "
                                              "codes are the parity block's reference"}
            line["cpu_baseline_stochastic"] = stochastic_cpu_block(x_host)
            line["reference_on_this_gpu"] = reference_on_gpu_block(x)
            line["frontend"] = frontend_block(device, peaks)
        if par["real_mismatches"] > 0:
            emit(line)
            raise SystemExit(f"bench.py: {par['real_mismatches']} real index mismatches against the reference path")
    emit(line)
    if world > 1:
        dist.barrier()
        if hasattr(gatherer, "close"):
            gatherer.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
