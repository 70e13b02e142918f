"""Ad-hoc probe (not a test): BASELINE config 3, strong scaling. 10 000 synthetic 30 s clips (2 250 frames x 768-d,
MERT-95M-shaped) = 22.5 M frames, dealt to the ranks round-robin by clip id (`shard_clips`), generated on the device
from seed + clip id (never on the host), encoded by both stacks in batches of 120 clips (270 000 frames per call), the
int16 index streams all-gathered over NCCL per batch on a side stream. Run alone (1 GPU) or under torchrun."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from neural_audio_tokenizer_b200 import ResidualVectorQuantizer, encode_stacks
from neural_audio_tokenizer_b200.sharding import CodeGatherer, shard_clips

world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=dev)
CLIPS, FRAMES, D, K, BATCH = int(os.environ.get("PROBE_CLIPS", 10000)), 2250, 768, 1024, 120
torch.manual_seed(42)
stacks = [ResidualVectorQuantizer(D, K, 4, use_stochastic=False).eval().to(dev) for _ in range(2)]
mine = list(shard_clips(CLIPS, world, rank))
per_rank = -(-CLIPS // world)
n_batches = -(-per_rank // BATCH)
gen = torch.Generator(device=dev)
x = torch.empty((1, D, BATCH * FRAMES), device=dev)
codes = torch.empty((8, 1, BATCH * FRAMES), dtype=torch.int16, device=dev)
gather = CodeGatherer(8, BATCH * FRAMES, world, dev) if world > 1 else None


def fill(batch):
    """features of this rank's clips of the batch, each from its own seed; short batches keep the previous tail"""
    ids = mine[batch * BATCH:(batch + 1) * BATCH]
    for j, cid in enumerate(ids):
        gen.manual_seed(1000 + cid)
        x[0, :, j * FRAMES:(j + 1) * FRAMES].normal_(generator=gen)
    return len(ids)


def run(timed):
    done = 0
    for b in range(n_batches):
        n_clips = fill(b)                          # generation is part of the job but not of the metric's hot path:
        if timed is not None:                      # timed separately below
            timed[0].record()
        encode_stacks(stacks, x, torch.int16, out=codes)
        if gather is not None:
            gather.all_gather(codes[:, 0])
        if timed is not None:
            timed[1].record(); torch.cuda.synchronize(); timed[2] += timed[0].elapsed_time(timed[1])
        done += n_clips
    if gather is not None:
        gather.wait()
    return done

run(None)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t = [torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), 0.0]
t0 = time.perf_counter()
done = run(t)
torch.cuda.synchronize()
wall = time.perf_counter() - t0
enc_ms = torch.tensor([t[2]], device=dev)
if world > 1:
    dist.all_reduce(enc_ms, op=dist.ReduceOp.MAX)
if rank == 0:
    frames = CLIPS * FRAMES
    print(f"config 3: {CLIPS} clips x {FRAMES} frames x {D}-d on {world} GPU(s): encode + index all-gather {enc_ms.item():.1f} ms "
          f"(max over ranks, CUDA events) = {frames / enc_ms.item() / 1e3:.1f} M frames/s; with on-device feature generation "
          f"{wall * 1e3:.0f} ms wall on rank 0", flush=True)
if world > 1:
    dist.destroy_process_group()
