"""-m gpu: GEMM / conv / attention cores through the C-ABI vs plain PyTorch fp32 references."""

import pytest
import torch
import torch.nn.functional as F

from gpu_common import engine, lib, relerr, stream
from depth_pro import _capi

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _gemm(backend, A, W, bias, act):
    M, K = A.shape
    N = W.shape[0]
    C = torch.empty(M, N, device=DEV)
    _capi.check(lib().dp_gemm_test(engine(), backend, A.data_ptr(), W.data_ptr(),
                                   None if bias is None else bias.data_ptr(), C.data_ptr(), M, N, K, act, stream()))
    torch.cuda.synchronize()
    return C


def _ref_gemm(A, W, bias, act):
    C = A.double() @ W.double().t()
    if bias is not None:
        C = C + bias.double()
    if act == _capi.ACT_RELU:
        C = F.relu(C)
    elif act == _capi.ACT_GELU:
        C = F.gelu(C)
    return C


@pytest.mark.parametrize("M,N,K,act", [(128, 128, 64, 0), (300, 384, 256, 2), (1155, 1024, 1024, 0), (577, 32, 1152, 1),
                                       (20195, 3072, 1024, 0)])
def test_gemm_fp32(M, N, K, act):
    g = torch.Generator(device=DEV).manual_seed(M + N)
    A = torch.randn(M, K, device=DEV, generator=g)
    W = torch.randn(N, K, device=DEV, generator=g) / K ** 0.5
    b = torch.randn(N, device=DEV, generator=g)
    got = _gemm(0, A, W, b, act)
    assert relerr(got, _ref_gemm(A, W, b, act)) < 2e-6


@pytest.mark.parametrize("M,N,K,act", [(128, 256, 64, 0), (128, 128, 64, 0), (128, 32, 64, 0), (256, 256, 256, 0),
                                       (300, 384, 256, 2), (1155, 1024, 1024, 1), (577, 32, 1152, 1),
                                       (20195, 3072, 1024, 0), (20195, 1024, 4096, 0)])
def test_gemm_bf16_tcgen05(M, N, K, act):
    g = torch.Generator(device=DEV).manual_seed(M + N + 1)
    A = torch.randn(M, K, device=DEV, generator=g)
    W = torch.randn(N, K, device=DEV, generator=g) / K ** 0.5
    b = torch.randn(N, device=DEV, generator=g)
    got = _gemm(1, A, W, b, act)
    # the kernel rounds A and W to bf16 and accumulates in fp32: compare against exactly that
    ref = _ref_gemm(A.bfloat16().float(), W.bfloat16().float(), b, act)
    assert relerr(got, ref) < 2e-5


# ---- epilogue forms of the tcgen05 GEMM (flags documented at dp_gemm_test in include/depthpro_b200.h)
def _bf16r(t):
    return t.bfloat16().float()


@pytest.mark.parametrize("M,N,K,act", [(300, 256, 128, 0), (21349, 3072, 1024, 0), (21349, 4096, 1024, 2),
                                       (1155, 128, 256, 1)])
def test_gemm_bf16_tma_store_epilogue(M, N, K, act):
    """bf16 row-major output through smem + TMA store (qkv / fc1+GELU path, single CTA and CTA pair)."""
    g = torch.Generator(device=DEV).manual_seed(7 * M + N)
    A = torch.randn(M, K, device=DEV, generator=g)
    W = torch.randn(N, K, device=DEV, generator=g) / K ** 0.5
    b = torch.randn(N, device=DEV, generator=g)
    got = _gemm(1, A, W, b, act | 0x100)
    ref = _ref_gemm(_bf16r(A), _bf16r(W), b, act)
    # output rounded to bf16: half an ulp = 2^-9 relative per element (erf-GELU approximation is 1e-6 abs)
    err = (got - ref).abs()
    assert (err <= ref.abs() * 2.0 ** -8 + 2e-5).all(), float(err.max())


@pytest.mark.parametrize("M,N,K", [(300, 256, 128), (21349, 1024, 1024), (21349, 1024, 4096)])
def test_gemm_bf16_fp32_residual_epilogue(M, N, K):
    """proj / fc2 form: x += gamma * (acc + bias) on the fp32 residual stream, in place."""
    g = torch.Generator(device=DEV).manual_seed(M + 3 * N + K)
    A = torch.randn(M, K, device=DEV, generator=g)
    W = torch.randn(N, K, device=DEV, generator=g) / K ** 0.5
    b = torch.randn(N, device=DEV, generator=g)
    res = torch.randn(M, N, device=DEV, generator=g)
    C = res.clone()
    _capi.check(lib().dp_gemm_test(engine(), 1, A.data_ptr(), W.data_ptr(), b.data_ptr(), C.data_ptr(), M, N, K, 0x200,
                                   stream()))
    torch.cuda.synchronize()
    ref = res.double() + b.double() * (_bf16r(A).double() @ _bf16r(W).double().t() + b.double())
    assert relerr(C, ref) < 2e-5


@pytest.mark.parametrize("M,N,act,offset", [(300, 256, 0, 0.0), (21349, 3072, 0, 0.3), (21349, 4096, 2, -0.5),
                                            (577, 1024, 0, 2.0)])
def test_gemm_bf16_layernorm_folded_consumer(M, N, act, offset):
    """qkv / fc1 with norm1 / norm2 folded in: LN(A) W^T + b computed as rstd * (A (g*W)^T) - rstd * mean *
    colsum(g*W) + (W b_ln + b) from the RAW bf16 rows and their (sum, sum of squares)."""
    K = 1024
    g = torch.Generator(device=DEV).manual_seed(5 * M + N)
    # rows with different scales and a common offset: the mean term must cancel
    A = torch.randn(M, K, device=DEV, generator=g) * (0.5 + torch.rand(M, 1, device=DEV, generator=g) * 4) + offset
    W = torch.randn(N, K, device=DEV, generator=g) / K ** 0.5
    b = torch.randn(N, device=DEV, generator=g)
    got = _gemm(1, A, W, b, act | 0x100 | 0x1000)
    k = torch.arange(K, device=DEV, dtype=torch.float32)
    gam, bet = 1 + 0.25 * torch.sin(0.37 * k), 0.1 * torch.cos(0.11 * k)
    y = F.layer_norm(A.double(), (K,), gam.double(), bet.double(), eps=1e-6)
    ref = _ref_gemm(y, W, b, act)
    # the fp32 LayerNorm + bf16 GEMM path this replaces rounds LN(A) and W to bf16 (2^-9 relative each) before a
    # K = 1024 dot product; the folded form rounds A and g*W instead.  Same error budget: a few 1e-3 of the
    # output scale, plus the bf16 rounding of the output itself.
    err = (got - ref).abs()
    scale = float(ref.abs().mean())
    assert float(err.mean()) < 4e-3 * scale, (float(err.mean()), scale)
    assert float(err.max()) < 6e-2 * float(ref.abs().max()), float(err.max())
    # and against the stand-alone path's own error on the same data
    base = _ref_gemm(_bf16r(y.float()), _bf16r(W), b, act)
    assert float(err.mean()) < 2.5 * float((base - ref).abs().mean()) + 2e-3 * scale


@pytest.mark.parametrize("M,K", [(300, 128), (21349, 1024), (21349, 4096)])
def test_gemm_bf16_layernorm_emitting_producer(M, K):
    """proj / fc2 form that also emits what the next folded GEMM needs: bf16(x) and per-row partial sums."""
    N = 1024
    g = torch.Generator(device=DEV).manual_seed(M + K)
    A = torch.randn(M, K, device=DEV, generator=g)
    W = torch.randn(N, K, device=DEV, generator=g) / K ** 0.5
    b = torch.randn(N, device=DEV, generator=g)
    res = torch.randn(M, N, device=DEV, generator=g) * 3 + 0.7
    C = torch.zeros(2 * M, N, device=DEV)
    C[:M] = res
    _capi.check(lib().dp_gemm_test(engine(), 1, A.data_ptr(), W.data_ptr(), b.data_ptr(), C.data_ptr(), M, N, K,
                                   0x200 | 0x2000, stream()))
    torch.cuda.synchronize()
    ref = res.double() + b.double() * (_bf16r(A).double() @ _bf16r(W).double().t() + b.double())
    assert relerr(C[:M], ref) < 2e-5                      # the residual update itself is unchanged
    ln = F.layer_norm(C[:M].double(), (N,), eps=1e-6)     # statistics of the UPDATED stream
    err = (C[M:].double() - ln).abs()
    # (bf16(x) - mean) * rstd: bf16 rounding of x relative to the row's spread
    assert float(err.max()) < 2.0 ** -8 * float((C[:M].abs().max(dim=1).values / C[:M].std(dim=1)).max()) + 1e-4
    assert float(err.mean()) < 2e-3


@pytest.mark.parametrize("M,K", [(300, 128), (21349, 1024), (21349, 4096)])
def test_gemm_bf16_pair_residual_producer(M, K):
    """The same producer over a residual stream stored as a (hi, lo) pair of 16-bit arrays (x = hi + lo, hi = the next
    folded GEMM's operand; common.cuh GemmOp::ln_xlo): the update must match the fp32-stream form to the pair's
    2^-16 resolution, the emitted statistics are those of the updated stream, and 48 consecutive updates (a whole
    ViT-L worth of proj / fc2) must not drift."""
    N = 1024
    g = torch.Generator(device=DEV).manual_seed(M + K + 1)
    A = torch.randn(M, K, device=DEV, generator=g)
    W = torch.randn(N, K, device=DEV, generator=g) / K ** 0.5
    b = torch.randn(N, device=DEV, generator=g)
    res = torch.randn(M, N, device=DEV, generator=g) * 3 + 0.7
    res[:, 5] *= 200.0                                     # an outlier channel, as DINOv2 has
    C = torch.zeros(2 * M, N, device=DEV)
    C[:M] = res
    flags = 0x200 | 0x2000 | 0x4000
    _capi.check(lib().dp_gemm_test(engine(), 1, A.data_ptr(), W.data_ptr(), b.data_ptr(), C.data_ptr(), M, N, K, flags, stream()))
    torch.cuda.synchronize()
    upd = b.double() * (_bf16r(A).double() @ _bf16r(W).double().t() + b.double())
    ref = res.double() + upd
    # element-wise: the pair resolves 2^-16 of each VALUE (two 8-bit mantissas + the sign of lo), on top of the fp32
    # accumulation-order error of the GEMM itself (the fp32-stream test allows 2e-5 of the tensor's maximum)
    gemm_err = 1e-5 * float(upd.abs().max())
    assert float(((C[:M].double() - ref).abs() - 2.0 ** -15 * ref.abs()).max()) < gemm_err
    ln = F.layer_norm(C[:M].double(), (N,), eps=1e-6)
    err = (C[M:].double() - ln).abs()
    assert float(err.max()) < 2.0 ** -8 * float((C[:M].abs().max(dim=1).values / C[:M].std(dim=1)).max()) + 1e-4
    assert float(err.mean()) < 2e-3
    if M == 300:
        x = C[:M].clone()
        want = ref.clone()
        for _ in range(47):
            C[:M] = x
            _capi.check(lib().dp_gemm_test(engine(), 1, A.data_ptr(), W.data_ptr(), b.data_ptr(), C.data_ptr(), M, N, K, flags,
                                           stream()))
            torch.cuda.synchronize()
            x = C[:M].clone()
            want += upd
        assert float(((x.double() - want).abs() - 48 * 2.0 ** -16 * want.abs()).max()) < 48 * gemm_err


@pytest.mark.parametrize("S,Cout,K,dual", [(32, 64, 128, 0), (96, 256, 256, 0), (192, 256, 256, 1), (32, 128, 64, 1)])
def test_gemm_bf16_convt_pixel_shuffle_epilogue(S, Cout, K, dual):
    """ConvTranspose2d k2 s2 as GEMM (N = (dy, dx, o)) + pixel shuffle through a 5-D TMA store; dual = ReLU twin."""
    g = torch.Generator(device=DEV).manual_seed(S + Cout + K)
    M, N = S * S, 4 * Cout
    A = torch.randn(M, K, device=DEV, generator=g)
    W = torch.randn(N, K, device=DEV, generator=g) / K ** 0.5
    b = torch.randn(Cout, device=DEV, generator=g)
    C = torch.empty(2 * S, 2 * S, Cout, device=DEV)
    flags = 0x400 | (0x800 if dual else 0)
    _capi.check(lib().dp_gemm_test(engine(), 1, A.data_ptr(), W.data_ptr(), b.data_ptr(), C.data_ptr(), M, N, K, flags,
                                   stream()))
    torch.cuda.synchronize()
    y = (_bf16r(A) @ _bf16r(W).t()).view(S, S, 2, 2, Cout) + b          # (y, x, dy, dx, o)
    ref = y.permute(0, 2, 1, 3, 4).reshape(2 * S, 2 * S, Cout)
    if dual:
        ref = ref.clamp_min(0)
    err = (C - ref).abs()
    assert (err <= ref.abs() * 2.0 ** -8 + 2e-5).all(), float(err.max())


@pytest.mark.parametrize("B,H,W,Cin,Cout,dual", [(1, 96, 96, 256, 256, 0), (1, 192, 192, 256, 256, 1), (2, 24, 48, 128, 128, 1)])
def test_conv3x3_bf16_tma_store_epilogue(B, H, W, Cin, Cout, dual):
    """NHWC bf16 conv output through the 4-D TMA store (CTA pair at 192^2), optional (x, relu(x)) dual store."""
    g = torch.Generator(device=DEV).manual_seed(H + Cin + dual)
    x = torch.randn(B, Cin, H, W, device=DEV, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, device=DEV, generator=g) / (9 * Cin) ** 0.5
    b = torch.randn(Cout, device=DEV, generator=g)
    got = _conv(1 | 0x100 | (0x800 if dual else 0), x, w, b)
    ref = F.conv2d(_bf16r(x), _bf16r(w), b, padding=1)
    if dual:
        ref = ref.clamp_min(0)
    err = (got - ref).abs()
    assert (err <= ref.abs() * 2.0 ** -8 + 3e-5).all(), float(err.max())


def _conv(backend, x_nchw, w, bias):
    B, Cin, H, W_ = x_nchw.shape
    Cout = w.shape[0]
    x = x_nchw.permute(0, 2, 3, 1).contiguous()
    y = torch.empty(B, H, W_, Cout, device=DEV)
    _capi.check(lib().dp_conv3x3_test(engine(), backend, x.data_ptr(), w.contiguous().data_ptr(),
                                      None if bias is None else bias.data_ptr(), y.data_ptr(), B, H, W_, Cin, Cout,
                                      stream()))
    torch.cuda.synchronize()
    return y.permute(0, 3, 1, 2)


@pytest.mark.parametrize("backend", [0, 1])
@pytest.mark.parametrize("B,H,W,Cin,Cout", [(1, 8, 16, 64, 128), (2, 24, 48, 128, 32), (1, 48, 48, 1024, 256),
                                            (1, 96, 96, 256, 256)])
def test_conv3x3(backend, B, H, W, Cin, Cout):
    g = torch.Generator(device=DEV).manual_seed(H + Cin)
    x = torch.randn(B, Cin, H, W, device=DEV, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, device=DEV, generator=g) / (9 * Cin) ** 0.5
    b = torch.randn(Cout, device=DEV, generator=g)
    got = _conv(backend, x, w, b)
    if backend == 1:
        ref = F.conv2d(x.bfloat16().double(), w.bfloat16().double(), b.double(), padding=1)
        tol = 2e-5
    else:
        ref = F.conv2d(x.double(), w.double(), b.double(), padding=1)
        tol = 5e-6  # K up to 9216 fp32 accumulations
    assert relerr(got, ref) < tol


def _sdpa_ref(qkv, bf16_inputs=True):
    n = qkv.shape[0]
    src = qkv.bfloat16().double() if bf16_inputs else qkv.double()
    q, k, v = src.reshape(n, 577, 3, 16, 64).permute(2, 0, 3, 1, 4)
    return F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(n, 577, 1024)


# bf16 kernels: the output is rounded to bf16 (half an ulp = 2^-9 = 1.95e-3 of the value) on top of the bf16 rounding of
# P; measured 2.7e-3 .. 3.0e-3 of the tensor's absmax -> tolerance 4e-3 (it was 1.5e-2).
ATTN_TOL_BF16 = 4e-3


@pytest.mark.parametrize("backend,n,tol", [(0, 2, 2e-6), (1, 3, ATTN_TOL_BF16), (1, 37, ATTN_TOL_BF16), (2, 3, ATTN_TOL_BF16)])
def test_attention(backend, n, tol):
    g = torch.Generator(device=DEV).manual_seed(n)
    qkv = torch.randn(n, 577, 3072, device=DEV, generator=g)
    out = torch.empty(n, 577, 1024, device=DEV)
    _capi.check(lib().dp_attention_test(engine(), backend, qkv.data_ptr(), out.data_ptr(), n, stream()))
    torch.cuda.synchronize()
    ref = _sdpa_ref(qkv, backend >= 1)   # 1 = tcgen05, 2 = mma.sync
    assert relerr(out, ref) < tol


def test_fp16_flavour_cores():
    """The fp16 build of the library (16-bit storage = IEEE half, `precision=torch.half`): tcgen05 GEMM, implicit-GEMM
    conv and attention against fp64 references on the FP16-rounded inputs -- proves the operand-format bits of the
    instruction descriptor, the TMA element type and every pack / unpack switched together."""
    L, E = lib("fp16"), engine(_capi.PREC_FP32, "fp16")
    assert L.dp_act_dtype() == b"fp16"
    g = torch.Generator(device=DEV).manual_seed(9)
    M, N, K = 1155, 512, 1024
    A = torch.randn(M, K, device=DEV, generator=g)
    W = torch.randn(N, K, device=DEV, generator=g) / K ** 0.5
    b = torch.randn(N, device=DEV, generator=g)
    C = torch.empty(M, N, device=DEV)
    _capi.check(L.dp_gemm_test(E, 1, A.data_ptr(), W.data_ptr(), b.data_ptr(), C.data_ptr(), M, N, K, 0, stream()), L)
    torch.cuda.synchronize()
    ref16 = A.half().double() @ W.half().double().t() + b.double()
    refbf = A.bfloat16().double() @ W.bfloat16().double().t() + b.double()
    assert relerr(C, ref16) < 2e-5                      # fp32 output: exact up to accumulation order
    assert relerr(C, refbf) > 1e-4                      # ... and it is NOT the bf16 arithmetic
    Cb = torch.empty(M, N, device=DEV)                  # 16-bit output through the TMA-store epilogue: half an fp16 ulp
    _capi.check(L.dp_gemm_test(E, 1, A.data_ptr(), W.data_ptr(), b.data_ptr(), Cb.data_ptr(), M, N, K, 0x100, stream()), L)
    torch.cuda.synchronize()
    assert relerr(Cb, ref16) < 6e-4
    # fp16 has no headroom above 65504: 16-bit outputs saturate (F2FP.SATFINITE) instead of turning into inf
    A2 = A.clone()
    A2[0] *= 3e4
    _capi.check(L.dp_gemm_test(E, 1, A2.data_ptr(), W.data_ptr(), b.data_ptr(), Cb.data_ptr(), M, N, K, 0x100, stream()), L)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(Cb).all()) and float(Cb[0].abs().max()) == 65504.0
    assert relerr(Cb[1:], ref16[1:]) < 6e-4
    x = torch.randn(1, 48, 48, 128, device=DEV, generator=g)
    w = torch.randn(256, 128, 3, 3, device=DEV, generator=g) / (9 * 128) ** 0.5
    y = torch.empty(1, 48, 48, 256, device=DEV)
    _capi.check(L.dp_conv3x3_test(E, 1, x.data_ptr(), w.data_ptr(), None, y.data_ptr(), 1, 48, 48, 128, 256, stream()), L)
    torch.cuda.synchronize()
    refc = F.conv2d(x.permute(0, 3, 1, 2).half().double(), w.half().double(), None, padding=1).permute(0, 2, 3, 1)
    assert relerr(y, refc) < 2e-5
    n = 3
    qkv = torch.randn(n, 577, 3072, device=DEV, generator=g)
    out = torch.empty(n, 577, 1024, device=DEV)
    for backend in (1, 2):                              # tcgen05 and mma.sync attention kernels
        _capi.check(L.dp_attention_test(E, backend, qkv.data_ptr(), out.data_ptr(), n, stream()), L)
        torch.cuda.synchronize()
        q, k, v = qkv.half().double().reshape(n, 577, 3, 16, 64).permute(2, 0, 3, 1, 4)
        ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(n, 577, 1024)
        assert relerr(out, ref) < 1e-3, backend          # fp16 P and output: 8x finer than the bf16 kernel's 4e-3


def _attn_variant(expv, pingpong, qkv):
    """Run the tcgen05 kernel in variant `expv`, then switch back to the default (the switch is process-wide)."""
    n = qkv.shape[0]
    out = torch.empty(n, 577, 1024, device=DEV)
    scratch = torch.empty_like(out)
    try:
        backend = 1 | ((expv + 1) << 8) | ((1 - pingpong) << 16)
        _capi.check(lib().dp_attention_test(engine(), backend, qkv.data_ptr(), out.data_ptr(), n, stream()))
    finally:
        _capi.check(lib().dp_attention_test(engine(), 1 | (0xFF << 8), qkv.data_ptr(), scratch.data_ptr(), n, stream()))
    torch.cuda.synchronize()
    return out, scratch


ATTN_VARIANTS = [(0, 1), (5, 1), (12, 1), (13, 1), (0, 0), (13, 0)]


@pytest.mark.parametrize("expv,pingpong", ATTN_VARIANTS)
@pytest.mark.parametrize("n", [3, 37])
def test_attention_variants(expv, pingpong, n):
    """Every compiled variant of the tcgen05 kernel (scalar / packed exp2 chain, P through shared memory or TMEM, 0 / 25 %
    of the exponentials as a polynomial on the FMA pipe, with and without the MUFU ping-pong) against fp64 SDPA on the
    bf16-rounded inputs.  One sequence has a 9x sharper softmax (near one-hot rows: P's own bf16 rounding is no longer
    averaged away, measured 4.3e-3) -> 6e-3 here, for the default kernel as well."""
    g = torch.Generator(device=DEV).manual_seed(100 + n)
    qkv = torch.randn(n, 577, 3072, device=DEV, generator=g)
    qkv[1] *= 3.0   # sharper softmax: scores far below the row maximum go through the clamped polynomial range
    out, base = _attn_variant(expv, pingpong, qkv)
    ref = _sdpa_ref(qkv)
    assert relerr(out, ref) < 6e-3
    assert relerr(base, ref) < 6e-3
    assert float((out - base).abs().mean() / base.abs().mean()) < 2e-3


def test_attention_variant_switch_changes_the_arithmetic():
    """identical max-errors across variants proved nothing.  The polynomial variants must differ from
    the all-MUFU ones in SOME output bits (different exp2), while the variant that only moves P (smem -> TMEM) must be
    bit-identical to its base."""
    g = torch.Generator(device=DEV).manual_seed(5)
    qkv = torch.randn(3, 577, 3072, device=DEV, generator=g)
    o = {v: _attn_variant(v, 1, qkv)[0] for v in (5, 12, 13)}
    assert torch.equal(o[5], o[12]), "P through TMEM must not change a bit"
    assert not torch.equal(o[12], o[13]), "polynomial share did not change the output"
    frac = float((o[12] != o[13]).float().mean())
    assert 1e-4 < frac < 0.5, frac


@pytest.mark.parametrize("expv", [0, 5, 12, 13])
def test_attention_lazy_rescale_fires_in_every_block(expv):
    """the lazy running maximum only rescales O and l when a row maximum grows by more than 2^8, which
    random inputs never do.  Here q and k share one direction d per head and the keys' component along d steps up by 7
    per 128-key block (q.d ~ 10), so the raw score grows by ~70 per block = 12.6 in the kernel's log2 units (> 8): EVERY
    later key block must take the rescale branch in every softmax warp.  The debug counter proves it did -- exactly
    4 blocks x 4 warps per (sequence, head, query tile) -- and the result still matches fp64 SDPA."""
    n = 2
    g = torch.Generator(device=DEV).manual_seed(11)
    qkv = torch.randn(n, 577, 3072, device=DEV, generator=g)
    q = qkv[..., :1024].reshape(n, 577, 16, 64)
    k = qkv[..., 1024:2048].reshape(n, 577, 16, 64)
    d = torch.nn.functional.normalize(torch.randn(16, 64, device=DEV, generator=g), dim=-1)
    blk = (torch.arange(577, device=DEV) // 128).float()
    q.copy_(0.3 * q + 10.0 * d)
    k.copy_(0.3 * k + (7.0 * blk)[None, :, None, None] * d)
    out = torch.empty(n, 577, 1024, device=DEV)
    scratch = torch.empty_like(out)
    try:
        assert lib().dp_debug_counter(engine(), 0, 1) >= 0          # reset
        _capi.check(lib().dp_attention_test(engine(), 1 | ((expv + 1) << 8), qkv.data_ptr(), out.data_ptr(), n, stream()))
        fired = lib().dp_debug_counter(engine(), 0, 1)
    finally:
        _capi.check(lib().dp_attention_test(engine(), 1 | (0xFF << 8), qkv.data_ptr(), scratch.data_ptr(), n, stream()))
    # 4 softmax warps per query tile, blocks 1..4 each rescale; the only warps that never do are those of the LAST
    # sequence's last tile whose 32 rows all lie beyond the tensor (TMA zero-fill: q = 0): 16 heads x 1 warp x 4 blocks
    assert fired == n * 16 * 5 * 4 * 4 - 16 * 4, fired
    # the softmax is concentrated on the last block's 65 keys: P's own bf16 rounding is averaged over few terms
    # (measured 4.3e-3, like the sharpened case of test_attention_variants)
    assert relerr(out, _sdpa_ref(qkv)) < 6e-3
    # and on benign inputs the branch never runs
    qkv2 = torch.randn(n, 577, 3072, device=DEV, generator=g)
    lib().dp_debug_counter(engine(), 0, 1)
    _capi.check(lib().dp_attention_test(engine(), 1, qkv2.data_ptr(), out.data_ptr(), n, stream()))
    assert lib().dp_debug_counter(engine(), 0, 1) == 0


def test_residual_l2_prefetch_is_a_pure_hint():
    """DEPTHPRO_RES_PREFETCH (cp.async.bulk.prefetch.L2 of the fp32 residual rows by the producer warp) must not
    change a single bit of the proj / fc2 form's output.  Switched through dp_kernel_bench's A/B bits."""
    import ctypes

    M, N, K = 21349, 1024, 1024
    g = torch.Generator(device=DEV).manual_seed(77)
    A = torch.randn(M, K, device=DEV, generator=g)
    W = torch.randn(N, K, device=DEV, generator=g) / K ** 0.5
    b = torch.randn(N, device=DEV, generator=g)
    res = torch.randn(M, N, device=DEV, generator=g)
    ms = ctypes.c_float()
    outs = []
    try:
        for bit in (0x200, 0x100):  # off, on
            _capi.check(lib().dp_kernel_bench(engine(), 2 | bit, 256, 256, 64, 1, ctypes.byref(ms)))
            C = res.clone()
            _capi.check(lib().dp_gemm_test(engine(), 1, A.data_ptr(), W.data_ptr(), b.data_ptr(), C.data_ptr(), M, N, K,
                                           0x200, stream()))
            torch.cuda.synchronize()
            outs.append(C)
    finally:
        _capi.check(lib().dp_kernel_bench(engine(), 2 | 0x200, 256, 256, 64, 1, ctypes.byref(ms)))
    assert torch.equal(outs[0], outs[1])
    ref = res.double() + b.double() * (_bf16r(A).double() @ _bf16r(W).double().t() + b.double())
    assert relerr(outs[1], ref) < 2e-5
