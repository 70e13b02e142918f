"""N>1 host logic on CPU: two gloo ranks shard the frames, encode their range, all-gather the int16 index streams.

The per-rank compute here is the oracle (tests may use it as a stand-in for the device call); what is under test is
the partition / pad / gather / unpad plumbing of neural_audio_tokenizer_b200.sharding, which bench.py uses with NCCL.
"""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from neural_audio_tokenizer_b200.sharding import CodeGatherer, all_gather_codes, shard_range
from oracle import rvq_oracle


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_frames, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        cbs = [torch.randn(64, 16) for _ in range(8)]                 # replicated codebooks: 4 semantic + 4 acoustic
        x = torch.randn(1, 16, n_frames, generator=torch.Generator().manual_seed(1))
        start, stop = shard_range(n_frames, world, rank)
        local = x[:, :, start:stop]
        sem = rvq_oracle.rvq_encode(local, cbs[:4])
        ac = rvq_oracle.rvq_encode(local, cbs[4:])
        local_codes = torch.stack([c[0] for c in sem + ac])          # [8, n_local]
        full = all_gather_codes(local_codes, n_frames)
        assert full.dtype == torch.int16 and full.shape == (8, n_frames)
        # the per-step gatherer bench.py uses (side stream on the GPU, in order here): same layout, twice in a row
        g = CodeGatherer(8, stop - start, world, "cpu", n_total=n_frames)
        for _ in range(3):
            again = g.all_gather(local_codes.to(torch.int16))
            g.wait()
            assert torch.equal(again, full)
        np.save(os.path.join(out_dir, f"rank{rank}.npy"), full.numpy())
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_allgather_matches_single_process(tmp_path):
    n_frames = 1001                                                   # odd: the last shard is short and gets padded
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n_frames, str(tmp_path)), nprocs=2, join=True)
    torch.manual_seed(0)
    cbs = [torch.randn(64, 16) for _ in range(8)]
    x = torch.randn(1, 16, n_frames, generator=torch.Generator().manual_seed(1))
    ref = torch.stack([c[0] for c in rvq_oracle.rvq_encode(x, cbs[:4]) + rvq_oracle.rvq_encode(x, cbs[4:])]).numpy()
    for rank in range(2):
        got = np.load(os.path.join(str(tmp_path), f"rank{rank}.npy"))
        np.testing.assert_array_equal(got, ref.astype(np.int16))


def test_single_process_gather_is_identity():
    codes = torch.arange(24, dtype=torch.int64).reshape(4, 6)
    out = all_gather_codes(codes, 6)
    assert out.dtype == torch.int16 and torch.equal(out.long(), codes)
