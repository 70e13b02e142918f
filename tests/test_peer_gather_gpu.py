"""All-gather of the index streams over NVLink peer memory (sharding.PeerCodeGatherer, csrc_host/peer_exchange.cpp) on two
GPUs, one process each: equal to torch.distributed's all-gather of the same blocks over many steps, ragged totals
included. Skipped on a box with one GPU (the driver's GPU test box; run with `gpurun --gpus 2`)."""
import os
import socket

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, steps, result):
    import torch.distributed as dist
    from neural_audio_tokenizer_b200.sharding import CodeGatherer, PeerCodeGatherer, shard_range
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    dev = torch.device("cuda", rank)
    L = 8
    start, stop = shard_range(n_total, world, rank)
    n_local = stop - start
    peer = PeerCodeGatherer(L, n_local, world, dev, n_total=n_total)
    nccl = CodeGatherer(L, n_local, world, dev, n_total=n_total)
    ok = True
    gen = torch.Generator(device=dev).manual_seed(100 + rank)
    outs = []
    for step in range(steps):
        local = torch.randint(0, 1024, (L, n_local), generator=gen, device=dev, dtype=torch.int16)
        a = peer.all_gather(local)
        b = nccl.all_gather(local)
        local.zero_()                                   # both staged their input: the caller may overwrite it at once
        outs.append((a, b, step))
        if len(outs) == 2:                              # consume one step late: the exchange overlaps the next step
            peer.wait(); nccl.wait()
            a0, b0, s0 = outs.pop(0)
            ok = ok and bool(torch.equal(a0, b0)) and tuple(a0.shape) == (L, n_total)
    peer.wait(); nccl.wait()
    for a0, b0, s0 in outs:
        ok = ok and bool(torch.equal(a0, b0))
    result[rank] = ok
    peer.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [4096, 4099])
def test_peer_all_gather_equals_nccl(n_total):
    import torch.multiprocessing as mp
    world = 2
    with mp.Manager() as m:
        result = m.dict()
        mp.spawn(_worker, args=(world, _free_port(), n_total, 7, result), nprocs=world, join=True)
        assert dict(result) == {0: True, 1: True}
