"""The two stacks of a tokenizer in one native call (nat_rvq_encode_stacks_f32 / nat_tokenize_host_f32): the same
index streams as stack-by-stack calls and as the oracle, for shared and separate inputs, ragged shapes, inputs of a
few tiles, batches, every code width, host buffers (one upload per chunk) and two host contexts in flight."""
import ctypes
import threading

import numpy as np
import pytest
import torch

from oracle import rvq_oracle

pytestmark = pytest.mark.gpu


def _stacks(D, K, L0, L1, seed=5):
    from neural_audio_tokenizer_b200 import ResidualVectorQuantizer
    torch.manual_seed(seed)
    return [ResidualVectorQuantizer(D, K, L, use_stochastic=False).eval().cuda() for L in (L0, L1)]


def _oracle_codes(stacks, xs):
    out = []
    for s, x in zip(stacks, xs):
        cbs = [q.codebook.cpu() for q in s.quantizers]
        out += [c for c in rvq_oracle.rvq_forward(x.cpu(), cbs)[1]]
    return torch.stack(out)                                           # [sum L, B, T]


def _assert_matches_oracle(stacks, xs, got):
    ref = _oracle_codes(stacks, xs).numpy()
    got = got.cpu().numpy().astype(np.int64)
    l0 = 0
    for s, x in zip(stacks, xs):
        L = len(s.quantizers)
        cbs = [q.codebook.cpu().numpy() for q in s.quantizers]
        rows = x.cpu().permute(0, 2, 1).reshape(-1, x.shape[1]).numpy()
        rep = rvq_oracle.classify_mismatches(rows, cbs, ref[l0:l0 + L].reshape(L, -1), got[l0:l0 + L].reshape(L, -1))
        assert rep["real_mismatches"] == 0, rep["flips"][:3]
        l0 += L


@pytest.mark.parametrize("D,K,L0,L1,B,T", [(768, 1024, 4, 4, 1, 20000),      # one fused launch, CTA pairs, both halves of the grid
                                           (768, 1024, 4, 4, 1, 20001 - 128),  # odd tile count: phantom tiles
                                           (256, 512, 4, 2, 2, 9000),          # stacks of different depth, batch of 2
                                           (80, 300, 3, 3, 1, 5000),           # ragged D and K
                                           (768, 1024, 4, 4, 1, 300),          # a few tiles: the latency path per stack
                                           (64, 128, 2, 2, 3, 1)])             # one frame per batch item
def test_encode_stacks_equals_per_stack_calls_and_oracle(D, K, L0, L1, B, T):
    from neural_audio_tokenizer_b200 import encode_stacks
    stacks = _stacks(D, K, L0, L1)
    x = torch.randn(B, D, T, generator=torch.Generator().manual_seed(3)).cuda()
    both = encode_stacks(stacks, x, torch.int16)
    assert both.shape == (L0 + L1, B, T) and both.dtype == torch.int16
    single = torch.cat([torch.stack(s.encode(x)) for s in stacks])
    assert torch.equal(both.long(), single)
    if B * T <= 20000:
        _assert_matches_oracle(stacks, [x, x], both)
    for dt in (torch.int32, torch.int64):
        assert torch.equal(encode_stacks(stacks, x, dt).long(), single)


def test_encode_stacks_with_one_input_per_stack():
    """The reference feeds the stacks different features (nat.py:3239-3240): separate preparations, one launch."""
    from neural_audio_tokenizer_b200 import encode_stacks
    stacks = _stacks(768, 1024, 4, 4, seed=9)
    xs = [torch.randn(1, 768, 12000, generator=torch.Generator().manual_seed(s)).cuda() for s in (1, 2)]
    both = encode_stacks(stacks, xs, torch.int16)
    single = torch.cat([torch.stack(s.encode(x)) for s, x in zip(stacks, xs)])
    assert torch.equal(both.long(), single)
    _assert_matches_oracle(stacks, xs, both)
    one = encode_stacks(stacks[:1], xs[0], torch.int16)               # a "tokenizer" of one stack
    assert torch.equal(one.long(), single[:4])


def test_encode_stacks_errors():
    from neural_audio_tokenizer_b200 import encode_stacks
    stacks = _stacks(32, 64, 2, 2)
    with pytest.raises(ValueError):
        encode_stacks(stacks, torch.randn(1, 31, 8, device="cuda"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        encode_stacks(stacks, torch.randn(1, 32, 8))
    with pytest.raises(ValueError):
        encode_stacks(stacks, [torch.randn(1, 32, 8, device="cuda")])
    stacks[1].quantizers[0].use_stochastic = True
    with pytest.raises(NotImplementedError):
        encode_stacks(stacks, torch.randn(1, 32, 8, device="cuda"))
    assert encode_stacks(_stacks(32, 64, 2, 2), torch.randn(1, 32, 0, device="cuda")).shape == (4, 1, 0)


@pytest.mark.parametrize("B,T", [(1, 150000), (3, 7001), (1, 40)])
def test_encode_stacks_host_uploads_once_and_matches_device_path(B, T, monkeypatch):
    from neural_audio_tokenizer_b200 import HostContext, encode_stacks, encode_stacks_host
    monkeypatch.setenv("NAT_HOST_CHUNK_ROWS", "32768")               # several chunks, double-buffered staging
    stacks = _stacks(256, 512, 4, 4)
    x = torch.randn(B, 256, T, generator=torch.Generator().manual_seed(8))
    dev = encode_stacks(stacks, x.cuda(), torch.int16).cpu()
    out = encode_stacks_host(stacks, x.pin_memory(), torch.int16)
    assert not out.is_cuda and torch.equal(out, dev)
    ctx = HostContext("cuda")
    again = encode_stacks_host(stacks, x, torch.int16, ctx=ctx)      # pageable input, caller-owned context, reused
    again2 = encode_stacks_host(stacks, x, torch.int16, ctx=ctx)
    assert torch.equal(again, dev) and torch.equal(again2, dev)
    ctx.close()
    assert torch.equal(stacks[0].encode_host(x, code_dtype=torch.int16), dev[:4])    # single-stack entry point (ABI 1)


def test_host_contexts_on_two_threads_share_the_codebook_handles():
    """Nothing of a host-buffer call lives on the codebook handle: two threads, each with its own context and stream,
    tokenize different clips through the same stacks at the same time."""
    from neural_audio_tokenizer_b200 import HostContext, encode_stacks, encode_stacks_host
    stacks = _stacks(128, 256, 4, 4)
    xs = [torch.randn(1, 128, 60000, generator=torch.Generator().manual_seed(s)).pin_memory() for s in (1, 2)]
    want = [encode_stacks(stacks, x.cuda(), torch.int16).cpu() for x in xs]
    got, errs = [None, None], []

    def work(i):
        try:
            torch.cuda.set_device(0)
            ctx = HostContext("cuda")
            with torch.cuda.stream(torch.cuda.Stream()):
                for _ in range(4):
                    got[i] = encode_stacks_host(stacks, xs[i], torch.int16, ctx=ctx)
            ctx.close()
        except Exception as e:           # surfaced in the main thread
            errs.append(e)

    threads = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errs, errs
    assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1])


@pytest.mark.parametrize("T_sem,T_ac", [(9000, 7013), (6000, 6000), (300, 260)])
def test_streams_of_different_length_are_aligned_like_the_reference(T_sem, T_ac):
    """nat.py:3225-3236: the longer feature stream is brought to the shorter time base with
    F.interpolate(mode='linear', align_corners=False) before both quantiser calls. Folded into the layer-0
    preparation's loads where the stacks share a launch form, a kernel of its own elsewhere; in both cases the index
    streams equal quantising torch's CPU interpolation (the reference's own arithmetic)."""
    import torch.nn.functional as F
    from neural_audio_tokenizer_b200 import encode_stacks
    stacks = _stacks(256, 512, 4, 4, seed=21)
    sem = torch.randn(1, 256, T_sem, generator=torch.Generator().manual_seed(1))
    ac = torch.randn(1, 256, T_ac, generator=torch.Generator().manual_seed(2))
    T = min(T_sem, T_ac)
    ref_in = [t if t.shape[-1] == T else F.interpolate(t, size=T, mode="linear", align_corners=False) for t in (sem, ac)]
    got = encode_stacks(stacks, [sem.cuda(), ac.cuda()], torch.int16)
    assert got.shape == (8, 1, T)
    want = torch.cat([torch.stack(s.encode(x.cuda())) for s, x in zip(stacks, ref_in)])     # quantise torch's CPU interpolation
    assert torch.equal(got.long(), want)
    _assert_matches_oracle(stacks, ref_in, got)
