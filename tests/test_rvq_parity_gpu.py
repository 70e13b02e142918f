"""Parity of the CUDA RVQ path (through the C ABI) against the oracle and the golden vectors. Needs a B200."""
import ctypes

import numpy as np
import pytest
import torch

from conftest import golden_rvq_inputs, load_golden
from oracle import rvq_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["small_input_path", "fused_kernel_only"], autouse=True)
def _dispatch(request, monkeypatch):
    """Inputs of a few tiles take the split-GEMM + row-argmin path by default; every test here also runs with that path
    switched off, so that the persistent fused kernel is exercised on the small golden cases as well."""
    monkeypatch.setenv("NAT_RVQ_SMALL", "1" if request.param == "small_input_path" else "0")
    yield

RVQ_CASES = ["rvq_small", "rvq_ragged", "rvq_ties", "rvq_single_frame", "rvq_768x1024", "rvq_512x4096",
             "rvq_1024x1024"]


def _dropin(cbs: torch.Tensor, **kw):
    from neural_audio_tokenizer_b200 import ResidualVectorQuantizer
    L, K, D = cbs.shape
    rvq = ResidualVectorQuantizer(D, K, L, use_stochastic=False).cuda().eval()
    with torch.no_grad():
        for q, cb in zip(rvq.quantizers, cbs):
            q.codebook.copy_(cb)
    for k, v in kw.items():
        setattr(rvq, k, v)
    return rvq


def _check_codes(x, cbs, ref_codes, got_codes, allow_near_ties=True):
    """ref/got: [L, B, T]. Bit-exact, or every primary flip is a near-tie (relative distance gap < 1e-6)."""
    L = ref_codes.shape[0]
    ref2, got2 = ref_codes.reshape(L, -1), got_codes.reshape(L, -1)
    if np.array_equal(ref2, got2):
        return {"near_tie_flips": 0, "real_mismatches": 0, "flips": []}
    x_nd = x.permute(0, 2, 1).reshape(-1, x.shape[1]).numpy()
    rep = rvq_oracle.classify_mismatches(x_nd, [c.numpy() for c in cbs], ref2, got2)
    assert rep["real_mismatches"] == 0, f"real index mismatches: {rep['flips'][:5]}"
    assert allow_near_ties, f"near-tie flips where none are expected: {rep['flips'][:5]}"
    return rep


@pytest.mark.parametrize("exact_scan", [False, True], ids=["tensor", "exact_scan"])
@pytest.mark.parametrize("name", RVQ_CASES)
def test_golden_codes_quantized_loss(name, exact_scan):
    g = load_golden(name)
    x, cbs = golden_rvq_inputs(g)
    rvq = _dropin(cbs, exact_scan=exact_scan)
    with torch.no_grad():
        quantized, codes, losses = rvq(x.cuda())
    torch.cuda.synchronize()
    assert len(codes) == int(g["L"]) and all(c.dtype == torch.int64 and c.shape == (x.shape[0], x.shape[2])
                                             for c in codes)
    got = np.stack([c.cpu().numpy() for c in codes])
    rep = _check_codes(x, cbs, g["codes"], got)
    assert losses["num_layers"] == int(g["L"])
    if rep["near_tie_flips"] == 0:
        # floating-point outputs. quantized: bit-exact (same fp32 op order as nat.py:2167/1405/1408).
        # vq_loss: the reference reduces in fp32 (MKL/vectorised order), we reduce in fp64: relative 2e-6.
        np.testing.assert_array_equal(quantized[:, :8, :8].cpu().numpy(), g["quantized_head"])
        assert abs(quantized.double().sum().item() - float(g["quantized_checksum"])) <= 1e-9 * max(
            1.0, abs(float(g["quantized_checksum"]))) + 1e-6
        if "quantized" in g.files:
            np.testing.assert_array_equal(quantized.cpu().numpy(), g["quantized"])
        assert abs(losses["vq_loss"].item() - float(g["vq_loss"])) <= 2e-6 * abs(float(g["vq_loss"]))


def test_ties_go_to_lowest_index():
    g = load_golden("rvq_ties")
    x, cbs = golden_rvq_inputs(g)
    for exact in (False, True):
        codes = _dropin(cbs, exact_scan=exact).encode(x.cuda())
        assert int(codes[0][0, 0]) == 3          # code 3 is duplicated at row 7; frame 0 equals it exactly


def test_tensor_core_accumulators_match_fp16_matmul():
    """The tcgen05 pass in isolation: raw accumulators vs an fp64 matmul of the same fp16 operands, and the
    accumulation-error constant of the decision window (DESIGN.md 'Exactness') holds with a wide margin."""
    from neural_audio_tokenizer_b200 import _lib
    lib = _lib.load()
    torch.manual_seed(3)
    L, K, D, N = 2, 1024, 768, 300
    cbs = [torch.randn(K, D, device="cuda") * (0.5 + l) for l in range(L)]
    x = torch.randn(N, D, device="cuda") * 2.5
    ptrs = (ctypes.c_void_p * L)(*[c.data_ptr() for c in cbs])
    handle = ctypes.c_void_p()
    _lib.check(lib.nat_rvq_codebooks_create(ptrs, L, K, D, None, ctypes.byref(handle)))
    try:
        ws_bytes = lib.nat_rvq_workspace_bytes(handle, N)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
        for layer in range(L):
            scores = torch.full((N, 1024), float("nan"), device="cuda")
            sx = torch.empty(N, device="cuda")
            sc = torch.empty(1, device="cuda")
            _lib.check(lib.nat_debug_rvq_scores(handle, layer, x.data_ptr(), N, scores.data_ptr(), sx.data_ptr(),
                                                sc.data_ptr(), ws.data_ptr(), ws_bytes, None))
            torch.cuda.synchronize()
            xt = (x * sx[:, None]).half().double()
            ct = (cbs[layer] * sc).half().double()
            ref = xt @ ct.T
            err = (scores.double() - ref).abs()
            bound = xt.norm(dim=1)[:, None] * ct.norm(dim=1)[None, :]
            assert torch.isfinite(scores).all()
            rel = (err / bound).max().item()
            assert rel < 2.0 ** -22 * D / 16, f"tensor-core accumulation error {rel:.3e} too close to the bound"
            assert 1.0 <= (x * sx[:, None]).abs().max(dim=1).values.min().item() and \
                (x * sx[:, None]).abs().max().item() < 2.0
    finally:
        lib.nat_rvq_codebooks_destroy(handle)


@pytest.mark.parametrize("D,K,L,N", [(768, 1024, 4, 6000), (1024, 4096, 2, 2000), (96, 300, 3, 1000),
                                      (512, 4096, 4, 1500)])
def test_tensor_path_equals_exact_scan(D, K, L, N):
    """Both device paths return the true fp64 argmin, so they must agree bit for bit (no near-tie allowance)."""
    torch.manual_seed(D + K)
    cbs = torch.randn(L, K, D)
    x = torch.randn(1, D, N).cuda()
    a = _dropin(cbs, collect_stats=True)
    ca = a.encode(x)
    cb_ = _dropin(cbs, exact_scan=True).encode(x)
    for u, v in zip(ca, cb_):
        assert torch.equal(u, v)
    st = a.last_stats.cpu().numpy()
    assert (st[:, :3].sum(axis=1) == N).all(), st
    assert st[:, 0].sum() > 0.5 * N * L, f"tensor-core pass certified too few frames: {st}"


def test_against_oracle_50k_frames():
    torch.manual_seed(42)
    L, K, D, N = 4, 1024, 768, 50_000
    cbs = torch.randn(L, K, D)
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(1, D, N, generator=g)
    rvq = _dropin(cbs, collect_stats=True)
    with torch.no_grad():
        quantized, codes, losses = rvq(x.cuda())
    ref_q, ref_codes, ref_losses = rvq_oracle.rvq_forward(x, list(cbs))
    got = np.stack([c.cpu().numpy() for c in codes])
    rep = _check_codes(x, cbs, np.stack([c.numpy() for c in ref_codes]), got)
    assert rep["near_tie_flips"] <= 5
    assert abs(losses["vq_loss"].item() - float(ref_losses["vq_loss"])) <= 1e-5 * float(ref_losses["vq_loss"])
    if rep["near_tie_flips"] == 0:
        assert torch.equal(quantized.cpu(), ref_q)


def test_layouts_dtypes_and_chunking_agree():
    from neural_audio_tokenizer_b200 import _lib
    lib = _lib.load()
    torch.manual_seed(5)
    L, K, D, B, T = 3, 512, 320, 2, 700
    N = B * T
    cbs = [torch.randn(K, D, device="cuda") for _ in range(L)]
    x = torch.randn(B, D, T, device="cuda")
    rows = x.permute(0, 2, 1).reshape(N, D).contiguous()
    ptrs = (ctypes.c_void_p * L)(*[c.data_ptr() for c in cbs])
    handle = ctypes.c_void_p()
    _lib.check(lib.nat_rvq_codebooks_create(ptrs, L, K, D, None, ctypes.byref(handle)))
    try:
        def run(src, layout, b, t, dtype, code, ws_rows, want_q):
            ws_bytes = lib.nat_rvq_workspace_bytes(handle, ws_rows)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
            codes = torch.empty((L, N), dtype=dtype, device="cuda")
            q = torch.empty_like(src) if want_q else None
            loss = torch.empty(L, device="cuda")
            _lib.check(lib.nat_rvq_encode_f32(handle, src.data_ptr(), layout, b, t, codes.data_ptr(), code,
                                              q.data_ptr() if want_q else None, loss.data_ptr(), 0.25, None,
                                              ws.data_ptr(), ws_bytes, 0, None))
            torch.cuda.synchronize()
            return codes, q, loss
        c64, q_bct, loss_a = run(x, _lib.LAYOUT_BCT, B, T, torch.int64, _lib.CODES_I64, N, True)
        c32, q_rows, loss_b = run(rows, _lib.LAYOUT_ROWS, 1, N, torch.int32, _lib.CODES_I32, N, True)
        c16, _, loss_c = run(x, _lib.LAYOUT_BCT, B, T, torch.int16, _lib.CODES_I16, 256, False)   # 6 chunks of 256
        assert torch.equal(c64, c32.long()) and torch.equal(c64, c16.long())
        assert torch.equal(q_bct.permute(0, 2, 1).reshape(N, D), q_rows)
        assert torch.equal(loss_a, loss_b)
        assert torch.allclose(loss_a, loss_c, rtol=1e-6)          # chunked partial sums add in a different order
        ref = rvq_oracle.rvq_encode(x.cpu(), [c.cpu() for c in cbs])
        _check_codes(x.cpu(), torch.stack([c.cpu() for c in cbs]), np.stack([r.numpy() for r in ref]),
                     c64.view(L, B, T).cpu().numpy())
        # host-buffer entry point
        host_codes = torch.empty((L, N), dtype=torch.int16).pin_memory()
        xh = x.cpu().pin_memory()
        _lib.check(lib.nat_rvq_encode_host_f32(handle, xh.data_ptr(), _lib.LAYOUT_BCT, B, T, host_codes.data_ptr(),
                                               _lib.CODES_I16, None))
        assert torch.equal(host_codes.long(), c64.cpu())
    finally:
        lib.nat_rvq_codebooks_destroy(handle)


def test_decode_matches_oracle_and_roundtrip():
    g = load_golden("rvq_small")
    x, cbs = golden_rvq_inputs(g)
    rvq = _dropin(cbs)
    codes = [torch.from_numpy(c).cuda() for c in g["codes"]]
    dec = rvq.decode(codes)
    np.testing.assert_array_equal(dec.cpu().numpy(), g["decoded"])
    # fewer code lists than layers are tolerated (nat.py:1442); an empty list gives the reference's zeros
    part = rvq.decode(codes[:2])
    assert torch.equal(part.cpu(), rvq_oracle.rvq_decode([c.cpu() for c in codes[:2]], list(cbs)))
    assert rvq.decode([]).shape == (1, 64, 1)
    # encode(decode(layer-0 codes)) returns the layer-0 codes: distance clamps to exactly 0 at the chosen code
    one = _dropin(cbs[:1])
    again = one.encode(one.decode(codes[:1]))
    assert torch.equal(again[0], codes[0])


@pytest.mark.parametrize("B,T,D,K,L,used", [(3, 45, 64, 128, 4, 4), (2, 33, 50, 96, 3, 3), (1, 1000, 768, 1024, 4, 4),
                                             (2, 70, 1024, 64, 6, 5), (1, 31, 36, 16, 2, 1)])
def test_decode_shapes_vs_oracle(B, T, D, K, L, used):
    """decode into [B, D, T] goes through a 32-frame tile: clips that end inside a tile, D that is no multiple of
    4 or 32, fewer code lists than layers, every code dtype (nat.py:1438-1444)."""
    from neural_audio_tokenizer_b200 import _lib
    gen = torch.Generator().manual_seed(B * 1000 + T)
    cbs = torch.randn(L, K, D, generator=gen)
    codes = [torch.randint(0, K, (B, T), generator=gen) for _ in range(used)]
    rvq = _dropin(cbs)
    ref = rvq_oracle.rvq_decode(codes, list(cbs))
    assert torch.equal(rvq.decode([c.cuda() for c in codes]).cpu(), ref)
    # the C ABI with narrower index streams, both layouts
    lib = _lib.load()
    handle = rvq._pack.get(rvq._codebooks())
    for dt, tdt in ((_lib.CODES_I32, torch.int32), (_lib.CODES_I16, torch.int16)):
        cd = torch.stack(codes).reshape(used, B * T).to(tdt).cuda().contiguous()
        out = torch.empty((B, D, T), device="cuda")
        _lib.check(lib.nat_rvq_decode_f32(handle, cd.data_ptr(), dt, used, B, T, _lib.LAYOUT_BCT, out.data_ptr(), None))
        rows = torch.empty((B * T, D), device="cuda")
        _lib.check(lib.nat_rvq_decode_f32(handle, cd.data_ptr(), dt, used, B, T, _lib.LAYOUT_ROWS, rows.data_ptr(), None))
        torch.cuda.synchronize()
        assert torch.equal(out.cpu(), ref)
        assert torch.equal(rows.reshape(B, T, D).transpose(1, 2).cpu(), ref)


def test_vector_quantizer_single_layer_and_2d_input():
    from neural_audio_tokenizer_b200 import VectorQuantizer
    g = load_golden("rvq_small")
    x, cbs = golden_rvq_inputs(g)
    vq = VectorQuantizer(64, 128, use_stochastic=False).cuda().eval()
    with torch.no_grad():
        vq.codebook.copy_(cbs[0])
    q, codes, loss = vq(x[0].cuda())                         # 2-D [C, T] input squeezes the batch dim again
    assert q.shape == (64, 50) and codes.shape == (50,)
    np.testing.assert_array_equal(codes.cpu().numpy(), g["codes"][0, 0])
    rq, rc, rl = rvq_oracle.vq_layer(x, cbs[0])
    assert torch.equal(q.cpu(), rq[0]) and abs(loss.item() - rl.item()) <= 2e-6 * rl.item()
    # codebook mutated in place (the cache loader does copy_, nat.py:593): derived device state must refresh
    with torch.no_grad():
        vq.codebook.copy_(cbs[1])
    _, codes2, _ = vq(x.cuda())
    np.testing.assert_array_equal(codes2.cpu().numpy(), rvq_oracle.vq_layer(x, cbs[1])[1].numpy())


def test_training_forward_samples_and_updates_ema_like_the_reference():
    """training_mode=True: every layer samples (nat.py:2150) and then runs `_update_ema` (nat.py:2179-2181,
    2205-2221) before the next layer sees the residual. Host-drawn noise from a seeded generator reproduces the
    restated reference chain: codes equal, quantised sum equal, EMA buffers and codebooks equal up to the fp32
    summation order of the one-hot matmul (device GEMM vs MKL)."""
    from neural_audio_tokenizer_b200 import ResidualVectorQuantizer
    torch.manual_seed(21)
    rvq = ResidualVectorQuantizer(16, 32, 3)
    cbs = [q.codebook.clone() for q in rvq.quantizers]
    cnt = [q.ema_count.clone() for q in rvq.quantizers]
    wgt = [q.ema_weight.clone() for q in rvq.quantizers]
    x = torch.randn(2, 16, 40, generator=torch.Generator().manual_seed(22))
    rvq = rvq.cuda().eval()
    torch.manual_seed(77)
    q_dev, codes_dev, losses_dev = rvq(x.cuda(), training_mode=True)
    assert not rvq.training
    torch.manual_seed(77)
    q_ref, codes_ref, losses_ref = rvq_oracle.rvq_forward_training(x, cbs, cnt, wgt)
    for a, b in zip(codes_dev, codes_ref):
        np.testing.assert_array_equal(a.cpu().numpy(), b.numpy())
    assert torch.allclose(q_dev.cpu(), q_ref, rtol=1e-5, atol=1e-5)
    assert abs(losses_dev["vq_loss"].item() - float(losses_ref["vq_loss"])) <= 1e-4 * float(losses_ref["vq_loss"])
    for l, ql in enumerate(rvq.quantizers):
        assert torch.allclose(ql.codebook.cpu(), cbs[l], rtol=1e-5, atol=1e-5)
        assert torch.allclose(ql.ema_count.cpu(), cnt[l], rtol=1e-6, atol=1e-6)
        assert torch.allclose(ql.ema_weight.cpu(), wgt[l], rtol=1e-5, atol=1e-5)


def test_collapsed_codebook_entries_do_not_flood_the_rerank():
    """Dead EMA entries collapse onto one vector (nat.py:2205-2221 divides a sum that decayed to zero by a count that
    decayed to zero): hundreds of exact duplicates. Ties go to the lowest index, so later duplicates are masked out of
    the coarse pass: indices still equal the reference's, and no frame falls back to the exact scan."""
    from neural_audio_tokenizer_b200 import _lib
    torch.manual_seed(13)
    D, K, L, N = 128, 512, 3, 4000
    cbs = torch.randn(L, K, D)
    dead = torch.randperm(K)[:300]
    cbs[:, dead] = 0.0                                                 # 300 collapsed entries per layer
    cbs[1, dead[:5]] = -0.0                                            # -0 == +0: still duplicates
    x = torch.randn(1, D, N) * 0.05                                    # frames near the origin: the zero code wins often
    rvq = _dropin(cbs, collect_stats=True)
    got = torch.stack(rvq.encode(x.cuda())).cpu().numpy()
    ref = np.stack([c.numpy() for c in rvq_oracle.rvq_forward(x, list(cbs))[1]])
    _check_codes(x, cbs, ref, got)
    assert (got == int(dead.min())).any()                              # the collapsed vector is chosen, by its lowest index
    st = rvq.last_stats.cpu().numpy()
    assert st[:, _lib.STAT_FIELDS - 2].sum() == 0, st                  # NAT_STAT_FULL_SCAN: nobody scanned all K codes
    rvq.exact_scan = True                                              # the exact path (no masking there) agrees
    np.testing.assert_array_equal(torch.stack(rvq.encode(x.cuda())).cpu().numpy(), got)


def test_errors_and_modes():
    from neural_audio_tokenizer_b200 import ResidualVectorQuantizer
    rvq = ResidualVectorQuantizer(32, 64, 2).cuda().eval()           # reference default: use_stochastic=True
    x = torch.randn(1, 32, 10, device="cuda")
    sampled = rvq(x)                                                  # eval-mode sampling runs natively now
    assert len(sampled[1]) == 2 and sampled[1][0].shape == (1, 10)    # (tests/test_rvq_sampling_gpu.py checks the codes)
    rvq.sampling_mode = "delegate"
    with pytest.raises(NotImplementedError):
        rvq(x)
    rvq.sampling_mode = "host_noise"
    for q in rvq.quantizers:
        q.use_stochastic = False
    out = rvq(x, training_mode=True)                                  # training runs natively too (next test)
    assert len(out) == 3 and len(out[1]) == 2 and not rvq.training
    with pytest.raises(ValueError):
        rvq(torch.randn(32, device="cuda"))
    with pytest.raises(ValueError):
        rvq(torch.randn(1, 33, 10, device="cuda"))
    q, codes, losses = rvq(x.transpose(1, 2).contiguous().transpose(1, 2))     # non-contiguous view
    assert q.shape == x.shape and len(codes) == 2
    empty = rvq.encode(torch.randn(1, 32, 0, device="cuda"))
    assert empty[0].shape == (1, 0)


def test_pipeline_fixture_tokens():
    """BASELINE config 1 at the path's boundary: the features the reference pipeline fed its quantisers on the test
    tone give the S/A arrays of its NDJSON frames."""
    g = load_golden("pipeline_tone_argmin")
    for feats, cbs, key in ((g["sem_in"], g["sem_codebooks"], "S"), (g["ac_in"], g["ac_codebooks"], "A")):
        rvq = _dropin(torch.from_numpy(cbs))
        rvq.codes_on_cpu = True
        codes = rvq.encode(torch.from_numpy(feats).cuda())
        assert all(not c.is_cuda for c in codes)
        got = np.stack([c[0].numpy() for c in codes], axis=1)
        np.testing.assert_array_equal(got, g[key])


def test_two_lane_overlap_matches_single_stream():
    """Long inputs are split over two internal streams; indices and the quantised sum must not depend on that."""
    from neural_audio_tokenizer_b200 import _lib
    lib = _lib.load()
    torch.manual_seed(8)
    L, K, D, N = 4, 512, 256, 90_000
    cbs = [torch.randn(K, D, device="cuda") for _ in range(L)]
    x = torch.randn(1, D, N, device="cuda")
    ptrs = (ctypes.c_void_p * L)(*[c.data_ptr() for c in cbs])
    handle = ctypes.c_void_p()
    _lib.check(lib.nat_rvq_codebooks_create(ptrs, L, K, D, None, ctypes.byref(handle)))
    try:
        ws_bytes = lib.nat_rvq_workspace_bytes(handle, N)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
        outs = []
        for flags in (0, _lib.RVQ_SINGLE_STREAM):
            codes = torch.empty((L, N), dtype=torch.int32, device="cuda")
            q = torch.empty_like(x)
            loss = torch.empty(L, device="cuda")
            stats = torch.zeros((L, 4), dtype=torch.int64, device="cuda")
            _lib.check(lib.nat_rvq_encode_f32(handle, x.data_ptr(), _lib.LAYOUT_BCT, 1, N, codes.data_ptr(),
                                              _lib.CODES_I32, q.data_ptr(), loss.data_ptr(), 0.25, stats.data_ptr(),
                                              ws.data_ptr(), ws_bytes, flags,
                                              torch.cuda.current_stream().cuda_stream))
            torch.cuda.synchronize()
            outs.append((codes, q, loss, stats))
        assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
        assert torch.allclose(outs[0][2], outs[1][2], rtol=1e-6)
        assert torch.equal(outs[0][3], outs[1][3]) and (outs[0][3][:, :3].sum(dim=1) == N).all()
        ref = rvq_oracle.rvq_encode(x[:, :, :4000].cpu(), [c.cpu() for c in cbs])
        _check_codes(x[:, :, :4000].cpu(), torch.stack([c.cpu() for c in cbs]), np.stack([r.numpy() for r in ref]),
                     outs[0][0][:, :4000].reshape(L, 1, 4000).cpu().numpy())
    finally:
        lib.nat_rvq_codebooks_destroy(handle)


@pytest.mark.parametrize("D,K,L,N", [(256, 512, 4, 60_000), (768, 1024, 4, 40_000), (1024, 1024, 2, 20_000),
                                      (96, 300, 3, 1000)])
def test_fused_stack_kernel_equals_layer_kernels(D, K, L, N, monkeypatch):
    """The one-launch stack kernel (any tile grouping) and the per-layer kernels take the same exact decisions:
    indices, quantised sum, per-layer losses and decision counters are identical."""
    from neural_audio_tokenizer_b200 import _lib
    lib = _lib.load()
    torch.manual_seed(D * 7 + K)
    cbs = [torch.randn(K, D, device="cuda") for _ in range(L)]
    x = torch.randn(1, D, N, device="cuda")
    ptrs = (ctypes.c_void_p * L)(*[c.data_ptr() for c in cbs])
    handle = ctypes.c_void_p()
    _lib.check(lib.nat_rvq_codebooks_create(ptrs, L, K, D, None, ctypes.byref(handle)))
    try:
        ws_bytes = lib.nat_rvq_workspace_bytes(handle, N)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")

        def run(fused, group, want_q_loss):
            monkeypatch.setenv("NAT_RVQ_FUSED", "1" if fused else "0")
            monkeypatch.setenv("NAT_RVQ_GROUP", str(group))
            codes = torch.full((L, N), -1, dtype=torch.int32, device="cuda")
            q = torch.empty_like(x) if want_q_loss else None
            loss = torch.empty(L, device="cuda") if want_q_loss else None
            stats = torch.zeros((L, 4), dtype=torch.int64, device="cuda")
            _lib.check(lib.nat_rvq_encode_f32(handle, x.data_ptr(), _lib.LAYOUT_BCT, 1, N, codes.data_ptr(),
                                              _lib.CODES_I32, q.data_ptr() if q is not None else None,
                                              loss.data_ptr() if loss is not None else None, 0.25, stats.data_ptr(),
                                              ws.data_ptr(), ws_bytes, 0, torch.cuda.current_stream().cuda_stream))
            torch.cuda.synchronize()
            return codes, q, loss, stats

        base = run(False, 2, True)
        assert (base[3][:, :3].sum(dim=1) == N).all()
        for group in (1, 2, 3, 1 << 20):
            for want in (True, False):
                got = run(True, group, want)
                assert torch.equal(got[0], base[0]), f"group {group}: indices differ from the per-layer kernels"
                assert (got[3][:, :3].sum(dim=1) == N).all(), got[3]
                if want:
                    assert torch.equal(got[1], base[1])
                    assert torch.equal(got[2], base[2])
        if N <= 20_000:
            ref = rvq_oracle.rvq_encode(x.cpu(), [c.cpu() for c in cbs])
            _check_codes(x.cpu(), torch.stack([c.cpu() for c in cbs]), np.stack([r.numpy() for r in ref]),
                         base[0].view(L, 1, N).cpu().numpy())
    finally:
        lib.nat_rvq_codebooks_destroy(handle)


def test_random_shapes_tensor_path_equals_exact_scan():
    """Differential sweep over shapes that exercise the CTA-pair schedule (odd / even tile counts, a phantom tile for
    the odd CTA of a pair, partial last tiles, fewer tiles than SMs, more than one tile per CTA), ragged D and K, the
    replayed residual updates (L from 1 to 8) and inputs far from unit scale. Both device paths return the true fp64
    argmin, so they must agree bit for bit; quantised sums and losses must agree as well."""
    rng = np.random.default_rng(20261018)
    dims = [8, 64, 72, 200, 256, 512, 768, 1024]
    for trial in range(40):
        D = int(rng.choice(dims))
        K = int(rng.choice([1, 2, 100, 256, 300, 1024, 2048]))
        L = int(rng.integers(1, 9))                          # the reference's default stack depth is 8
        N = int(rng.choice([1, 2, 127, 128, 129, 255, 257, 300, 1000, 128 * 149, 128 * 148 * 2 + 5, 40001]))
        if trial == 0:
            D, K, L, N = 512, 8192, 8, 700                   # deep stack, large codebook
        if N > 20000 and (K > 1024 or D > 768 or L > 4):
            N = 3001
        scale = float(rng.choice([1e-3, 1.0, 1.0, 37.0]))
        g = torch.Generator().manual_seed(int(rng.integers(1 << 30)))
        cbs = torch.randn(L, K, D, generator=g) * float(rng.choice([0.2, 1.0, 5.0]))
        x = (torch.randn(1, D, N, generator=g) * scale).cuda()
        fast = _dropin(cbs)
        slow = _dropin(cbs, exact_scan=True)
        with torch.no_grad():
            qa, ca, la = fast(x)
            qb, cb_, lb = slow(x)
        for l, (u, v) in enumerate(zip(ca, cb_)):
            assert torch.equal(u, v), (trial, D, K, L, N, scale, l, int((u != v).sum()))
        assert torch.equal(qa, qb), (trial, D, K, L, N)
        assert abs(la["vq_loss"].item() - lb["vq_loss"].item()) <= 1e-6 * max(1.0, abs(lb["vq_loss"].item()))
        enc = fast.encode(x)                                 # codes-only form takes the hot update path
        for u, v in zip(enc, cb_):
            assert torch.equal(u, v), (trial, D, K, L, N, "encode")


def test_chunk_path_is_cuda_graph_capturable():
    """BASELINE config 5 (1 s chunks): encode() issues only stream-ordered work on the current stream, so the chunk
    path can be captured once and replayed; replays follow new input and equal the eager codes."""
    torch.manual_seed(21)
    cbs = torch.randn(4, 1024, 768)
    rvq = _dropin(cbs)
    x = torch.randn(1, 768, 75, device="cuda")
    eager = [c.clone() for c in rvq.encode(x)]                # also builds the codebook handle outside the capture
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            rvq.encode(x)
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = rvq.encode(x)
    g.replay()
    torch.cuda.synchronize()
    assert all(torch.equal(a, b) for a, b in zip(eager, out))
    x.copy_(torch.randn(1, 768, 75, device="cuda"))
    g.replay()
    torch.cuda.synchronize()
    fresh = rvq.encode(x)
    assert all(torch.equal(a, b) for a, b in zip(fresh, out))
