"""Frame / clip sharding across the GPUs of one box and the all-gather of the index streams.

Frames are independent inside the RVQ (the only sequential dependency is across layers within a frame,
nat.py:1398-1405), so the path shards with no data-path collective: codebooks are replicated, every rank encodes a
contiguous frame range, and only the int16 index streams ([L_total, frames] = 16 bytes per frame for 8 layers) are
exchanged. NCCL has no 16-bit integer type, so the payload travels as bytes. Works on any torch.distributed backend
(`nccl` on the B200 box, `gloo` in the CPU tests).
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous [start, stop) of `n_items` for `rank`: equal ceil-sized shards, the tail ranks may be short/empty."""
    per = -(-n_items // world_size) if world_size > 0 else n_items
    start = min(n_items, rank * per)
    return start, min(n_items, start + per)


def shard_clips(n_clips: int, world_size: int, rank: int):
    """Round-robin clip ids for a corpus (BASELINE.json config 3: clip_id mod G)."""
    return range(rank, n_clips, world_size)


def all_gather_codes(local_codes: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """local_codes [L, n_local] (any integer dtype, K <= 32768) for this rank's `shard_range` -> [L, n_total] int16
    on every rank. Shards are padded to the common ceil size because all-gather needs equal counts."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    L = local_codes.shape[0]
    per = -(-n_total // world)
    start, stop = shard_range(n_total, world, rank)
    if local_codes.shape[1] != stop - start:
        raise ValueError(f"rank {rank} holds {local_codes.shape[1]} frames, its shard is {stop - start}")
    send = torch.zeros((L, per), dtype=torch.int16, device=local_codes.device)
    send[:, :stop - start] = local_codes.to(torch.int16)
    if world == 1:
        return send[:, :n_total].contiguous()
    recv = torch.empty((world, L, per), dtype=torch.int16, device=local_codes.device)
    dist.all_gather_into_tensor(recv.view(torch.uint8).view(-1), send.view(torch.uint8).view(-1), group=group)
    # [world, L, per] -> [L, world * per] -> drop the padding of the last shards
    return recv.permute(1, 0, 2).reshape(L, world * per)[:, :n_total].contiguous()


class CodeGatherer:
    """Overlapped form of `all_gather_codes` for a loop of encode steps: the exchange of step i runs on a side stream
    under the kernels of step i + 1.

    Per step, on the caller's stream: a copy of the rank's [L, n_local] index streams into one of two staging buffers
    (a few MB; the encode kernels may then overwrite their output at once). On the side stream: one
    `all_gather_into_tensor` of the staging buffer as bytes (NCCL has no 16-bit integer type) and one strided copy
    that lays the [world, L, per] blocks out as [L, world * per], the layout the consumers index. `wait()` makes the
    caller's stream wait for everything outstanding; `all_gather` returns the output tensor of this step (valid after
    `wait()` or after the next-but-one call). Shards are padded to the common ceil size."""

    def __init__(self, n_layers: int, n_local: int, world: int, device, n_total: int = None, group=None):
        self.L, self.world, self.group = n_layers, world, group
        self.n_total = n_total if n_total is not None else n_local * world
        self.per = -(-self.n_total // world)
        self.n_local = n_local
        self.on_gpu = torch.device(device).type == "cuda"    # host tensors (gloo, the CPU tests): same plumbing, in order
        self.comm = torch.cuda.Stream(device=device) if self.on_gpu else None
        self.send = [torch.zeros((n_layers, self.per), dtype=torch.int16, device=device) for _ in range(2)]
        self.recv = [torch.empty((world, n_layers, self.per), dtype=torch.int16, device=device) for _ in range(2)]
        self.out = [torch.empty((n_layers, world * self.per), dtype=torch.int16, device=device) for _ in range(2)]
        self.staged = [torch.cuda.Event() for _ in range(2)] if self.on_gpu else None
        self.done = [torch.cuda.Event() for _ in range(2)] if self.on_gpu else None
        self.used = [False, False]
        self.i = 0

    def all_gather(self, local_codes: torch.Tensor) -> torch.Tensor:
        if tuple(local_codes.shape) != (self.L, self.n_local):
            raise ValueError(f"expected [{self.L}, {self.n_local}] index streams, got {tuple(local_codes.shape)}")
        k = self.i & 1
        self.i += 1
        result = self.out[k][:, :self.n_total] if self.world * self.per != self.n_total else self.out[k]
        if not self.on_gpu:
            self.send[k][:, :self.n_local].copy_(local_codes)
            dist.all_gather_into_tensor(self.recv[k].view(torch.uint8).view(-1), self.send[k].view(torch.uint8).view(-1),
                                        group=self.group)
            self.out[k].view(self.L, self.world, self.per).copy_(self.recv[k].permute(1, 0, 2))
            return result
        cur = torch.cuda.current_stream(local_codes.device)
        if self.used[k]:
            cur.wait_event(self.done[k])                    # the exchange that last used this slot (two steps ago)
        self.send[k][:, :self.n_local].copy_(local_codes)
        self.staged[k].record(cur)
        with torch.cuda.stream(self.comm):
            self.comm.wait_event(self.staged[k])
            dist.all_gather_into_tensor(self.recv[k].view(torch.uint8).view(-1), self.send[k].view(torch.uint8).view(-1),
                                        group=self.group)
            self.out[k].view(self.L, self.world, self.per).copy_(self.recv[k].permute(1, 0, 2))
            self.done[k].record(self.comm)
        self.used[k] = True
        return result

    def wait(self) -> None:
        if not self.on_gpu:
            return
        cur = torch.cuda.current_stream(self.send[0].device)
        for k in range(2):
            if self.used[k]:
                cur.wait_event(self.done[k])


class _DeviceArray:
    """A raw device allocation of the native library seen as a torch tensor (__cuda_array_interface__, no copy)."""

    def __init__(self, ptr: int, shape, typestr: str, owner):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3}
        self._owner = owner


class PeerCodeGatherer:
    """`CodeGatherer` without a collective kernel: every rank's [L, n_local] index streams are written by the copy
    engines over NVLink straight into their column range of every rank's [L, world * per] output
    (nat_peer_all_gather, csrc_host/peer_exchange.cpp). No SM is involved, so the exchange does not compete with the
    persistent stack kernel of the next step the way NCCL's all-gather kernel does. One process per GPU on one box
    (CUDA IPC); torch.distributed is used once, to pass the IPC handles around.

    Same interface as CodeGatherer: `all_gather(local)` stages the rank's streams on the caller's stream, runs the
    exchange on a side stream and returns the output tensor of this step, valid after `wait()` (or in the side
    stream's order) and until the call after next (three output buffers: a peer may start writing step s + 2 while
    this rank still reads step s)."""

    def __init__(self, n_layers: int, n_local: int, world: int, device, n_total: int = None, group=None):
        from . import _lib
        self._lib_mod = _lib
        lib = _lib.load()
        self.L, self.world, self.group = n_layers, world, group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n_total = n_total if n_total is not None else n_local * world
        self.per = -(-self.n_total // world)
        self.n_local = n_local
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("PeerCodeGatherer moves device memory over NVLink: it needs CUDA devices (CodeGatherer "
                               "runs on any torch.distributed backend)")
        import ctypes
        self._ctx = ctypes.c_void_p()
        # Set-up is collective: a rank that fails locally still takes part in every exchange below, and all ranks
        # agree on the outcome before anyone returns (or raises), so a caller can fall back to CodeGatherer everywhere.
        err = None
        with torch.cuda.device(self.device):
            mine = (ctypes.c_ubyte * _lib.PEER_HANDLE_BYTES)()
            try:
                _lib.check_peer(lib.nat_peer_create(world, self.rank, n_layers, self.per * 2, ctypes.byref(self._ctx)))
                _lib.check_peer(lib.nat_peer_export(self._ctx, mine))
            except Exception as e:                              # noqa: BLE001 - reported after the agreement
                err = e
            if world > 1:
                send = torch.frombuffer(bytearray(mine), dtype=torch.uint8).to(self.device)
                recv = torch.empty(world * _lib.PEER_HANDLE_BYTES, dtype=torch.uint8, device=self.device)
                dist.all_gather_into_tensor(recv, send, group=group)
                if err is None:
                    try:
                        blob = bytes(recv.cpu().numpy().tobytes())
                        _lib.check_peer(lib.nat_peer_connect(self._ctx, ctypes.c_char_p(blob)))
                    except Exception as e:                      # noqa: BLE001
                        err = e
                self._agree(err, "set-up")                      # also: every rank has opened every buffer
                # one empty exchange proves the stream memory operations work here before a caller relies on them
                try:
                    probe = torch.zeros((n_layers, self.per), dtype=torch.int16, device=self.device)
                    got = ctypes.c_void_p()
                    _lib.check_peer(lib.nat_peer_all_gather(self._ctx, probe.data_ptr(), self.per * 2,
                                                       torch.cuda.current_stream(self.device).cuda_stream, ctypes.byref(got)))
                    torch.cuda.synchronize(self.device)
                except Exception as e:                          # noqa: BLE001
                    err = e
                self._agree(err, "first exchange")
            elif err is not None:
                raise err
        self.out = [torch.as_tensor(_DeviceArray(lib.nat_peer_buffer(self._ctx, k), (n_layers, world * self.per), "<i2", self),
                                    device=self.device) for k in range(3)]
        self.comm = torch.cuda.Stream(device=self.device)
        self.send = [torch.zeros((n_layers, self.per), dtype=torch.int16, device=self.device) for _ in range(2)]
        self.staged = [torch.cuda.Event() for _ in range(2)]
        self.done = [torch.cuda.Event() for _ in range(2)]
        self.used = [False, False]
        self.i = 1 if world > 1 else 0                          # the probe exchange was the library's step 1

    def _agree(self, err, what: str) -> None:
        flag = torch.tensor([0 if err is None else 1], dtype=torch.int32, device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=self.group)
        if int(flag.item()) != 0:
            lib = self._lib_mod.load()
            if self._ctx:
                lib.nat_peer_disconnect(self._ctx)
            dist.barrier(group=self.group)                      # nobody maps anybody's buffers any more
            if self._ctx:
                lib.nat_peer_destroy(self._ctx)
                self._ctx = None
            raise RuntimeError(f"peer-memory exchange unavailable ({what}) on at least one rank"
                               + (f"; this rank: {err}" if err is not None else ""))

    def all_gather(self, local_codes: torch.Tensor) -> torch.Tensor:
        import ctypes
        if tuple(local_codes.shape) != (self.L, self.n_local):
            raise ValueError(f"expected [{self.L}, {self.n_local}] index streams, got {tuple(local_codes.shape)}")
        lib = self._lib_mod.load()
        self.i += 1
        k = self.i & 1                                          # staging buffers alternate
        kb = self.i % 3                                         # the library's output buffer of step i
        cur = torch.cuda.current_stream(self.device)
        if self.used[k]:
            cur.wait_event(self.done[k])                        # the staging buffer's previous exchange has been sent
        self.send[k][:, :self.n_local].copy_(local_codes)
        self.staged[k].record(cur)
        self.comm.wait_event(self.staged[k])
        got = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            self._lib_mod.check_peer(lib.nat_peer_all_gather(self._ctx, self.send[k].data_ptr(), self.per * 2,
                                                        self.comm.cuda_stream, ctypes.byref(got)))
        assert got.value == self.out[kb].data_ptr()
        self.done[k].record(self.comm)
        self.used[k] = True
        return self.out[kb][:, :self.n_total] if self.world * self.per != self.n_total else self.out[kb]

    def wait(self) -> None:
        cur = torch.cuda.current_stream(self.device)
        for k in range(2):
            if self.used[k]:
                cur.wait_event(self.done[k])

    def close(self) -> None:
        """Collective. Two-phase: everybody stops writing, everybody unmaps the peers' buffers, and only then does
        anyone free memory the others had mapped (freeing an exported allocation an importer still maps is undefined;
        the first version of this method freed right after one barrier and hung torchrun jobs at exit)."""
        if self._ctx:
            lib = self._lib_mod.load()
            multi = dist.is_initialized() and self.world > 1
            torch.cuda.synchronize(self.device)
            if multi:
                dist.barrier(group=self.group)                  # no rank still pushes into anybody's buffers
            lib.nat_peer_disconnect(self._ctx)
            if multi:
                dist.barrier(group=self.group)                  # no rank still maps anybody's buffers
            lib.nat_peer_destroy(self._ctx)
            self._ctx = None
