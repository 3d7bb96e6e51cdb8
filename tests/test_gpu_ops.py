"""Operator-level parity (B200, `-m gpu`): the tcgen05 GEMM with every fused epilogue and the attention kernels, called
through the C ABI, against a plain PyTorch fp32/fp64 reference of the same op on the same (bf16-rounded) operands.

Tolerances: a bf16 output carries 2^-9 relative rounding (4e-3 of the tensor scale asserted as 6e-3); the packed-f16
GELU epilogue adds ~5e-4 rms (1e-2 asserted); fp32 outputs (split-K atomics, 3-plane mode, column sums) 2e-5.
"""
import numpy as np
import pytest
import torch

from simxns_b200 import _lib

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _device():
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    torch.cuda.set_device(0)
    _lib.require_device()


def split3(x):
    hi = x.to(torch.bfloat16)
    r = x - hi.float()
    mid = r.to(torch.bfloat16)
    lo = (r - mid.float()).to(torch.bfloat16)
    return torch.stack([hi, mid, lo])


def run_gemm(M, N, K, mn, planes, epi, ksplit, seed, colsum=False):
    lib = _lib.lib()
    torch.manual_seed(seed)
    dev = "cuda"
    A32 = torch.randn(M, K, device=dev) * (0.05 if epi in (_lib.EPI_BIAS_GELU, _lib.EPI_BIAS_GELU_FWD) else 1.0)
    B32 = torch.randn(N, K, device=dev)
    bias = torch.randn(N, device=dev)
    aux32 = torch.randn(M, N, device=dev)
    if planes == 1:
        A, B, aux = (t.to(torch.bfloat16)[None] for t in (A32, B32, aux32))
    else:
        A, B, aux = split3(A32), split3(B32), split3(aux32)
    Af, Bf, auxf = A.float().sum(0), B.float().sum(0), aux.float().sum(0)
    if mn:  # stored transposed: [K, M] / [K, N]
        A_dev, B_dev, lda, ldb = A.transpose(1, 2).contiguous(), B.transpose(1, 2).contiguous(), M, N
    else:
        A_dev, B_dev, lda, ldb = A.contiguous(), B.contiguous(), K, K
    out = torch.full((planes, M, N), float("nan"), device=dev, dtype=torch.bfloat16)
    out2 = torch.full((planes, M, N), float("nan"), device=dev, dtype=torch.bfloat16)
    if epi == _lib.EPI_ATOMIC_F32:
        outf = torch.zeros(M, N, device=dev)
    else:
        outf = torch.full((N,), 0.25, device=dev) if colsum else None  # column sums ACCUMULATE into outf
    ref = Af.double() @ Bf.double().t()
    if epi in (_lib.EPI_BIAS, _lib.EPI_BIAS_GELU, _lib.EPI_BIAS_ADD, _lib.EPI_BIAS_GELU_FWD):
        ref = ref + bias.double()
    if epi == _lib.EPI_BIAS_GELU_FWD:  # inference epilogue: gelu only
        ref = torch.nn.functional.gelu(ref)
    if epi in (_lib.EPI_BIAS_ADD, _lib.EPI_ADD):
        ref = ref + auxf.double()
    ref2 = None
    if epi == _lib.EPI_BIAS_GELU:  # out = gelu'(pre), out2 = gelu(pre)
        ref2 = torch.nn.functional.gelu(ref)
        cdf = 0.5 * (1 + torch.erf(ref / 2 ** 0.5))
        pdf = torch.exp(-0.5 * ref * ref) / (2 * np.pi) ** 0.5
        ref = cdf + ref * pdf
    if epi == _lib.EPI_MUL:
        ref = ref * auxf.double()
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.simx_gemm(M, N, K, A_dev.data_ptr(), lda, A_dev.stride(0), mn, B_dev.data_ptr(), ldb, B_dev.stride(0),
                             mn, planes, epi, bias.data_ptr(), aux.data_ptr(), N, aux.stride(0), out.data_ptr(),
                             out2.data_ptr(), N, out.stride(0), _lib.ptr(outf), N, ksplit, st), "simx_gemm")
    torch.cuda.synchronize()
    got = outf.double() if epi == _lib.EPI_ATOMIC_F32 else out.float().sum(0).double()
    assert not torch.isnan(got).any()
    err = ((got - ref).abs().max() / ref.abs().max()).item()
    err2 = None
    if ref2 is not None:
        got2 = out2.float().sum(0).double()
        err2 = ((got2 - ref2).abs().max() / ref2.abs().max()).item()
    cs_err = None
    if colsum:
        # the column sums are of the STORED (rounded) output, exactly what a separate pass over `out` would add up
        want = 0.25 + out.float().sum(0).double().sum(0)
        cs_err = ((outf.double() - want).abs().max() / want.abs().max()).item()
    return err, err2, cs_err


# (M, N, K, mn_major, planes, epilogue, ksplit): ragged M / N tails, all three tile widths, both operand majors
GEMM_CASES = [
    (128, 256, 64, 0, 1, 0, 1), (1000, 768, 768, 0, 1, 0, 1), (300, 64, 192, 0, 1, 0, 1), (300, 128, 320, 0, 1, 0, 1),
    (520, 328, 200, 0, 1, 0, 1), (512, 2304, 768, 0, 1, 1, 1), (512, 3072, 768, 0, 1, 2, 1), (512, 768, 3072, 0, 1, 3, 1),
    (512, 3072, 768, 0, 1, 4, 1), (512, 768, 2304, 0, 1, 5, 1), (512, 768, 768, 0, 1, 6, 3), (8192, 3072, 768, 0, 1, 2, 1),
    (2304, 768, 4096, 1, 1, 6, 1), (768, 3072, 8192, 1, 1, 6, 8), (768, 768, 1000, 1, 1, 6, 4), (192, 64, 520, 1, 1, 6, 2),
    (384, 768, 768, 0, 3, 1, 1), (256, 3072, 768, 0, 3, 2, 1), (768, 768, 2048, 1, 3, 6, 4),
    (512, 3072, 768, 0, 1, 7, 1), (300, 128, 192, 0, 1, 7, 1), (256, 3072, 768, 0, 3, 7, 1), (1000, 328, 128, 0, 1, 7, 1),
]


@pytest.mark.parametrize("case", GEMM_CASES, ids=lambda c: "x".join(str(v) for v in c))
def test_gemm_epilogues_vs_torch(case):
    M, N, K, mn, planes, epi, ksplit = case
    err, err2, _ = run_gemm(M, N, K, mn, planes, epi, ksplit, seed=M + N + K)
    gelu = epi in (_lib.EPI_BIAS_GELU, _lib.EPI_BIAS_GELU_FWD)
    tol = 1e-5 if planes == 3 else (2e-5 if epi == _lib.EPI_ATOMIC_F32 else (1e-2 if gelu else 6e-3))
    assert err < tol, (case, err)
    if err2 is not None:
        assert err2 < tol, (case, err2)


@pytest.mark.parametrize("M,N,K,planes", [(512, 3072, 768, 1),    # staged epilogue: sums reduced in shared memory
                                          (1000, 192, 128, 1),    # ragged rows, N < tile width (column guard)
                                          (4100, 3072, 256, 1),   # more tiles than CTAs: per-CTA accumulation over tiles
                                          (300, 3200, 128, 1),    # N beyond the smem accumulator -> column-sum kernel
                                          (300, 128, 192, 1),     # narrow tile (register epilogue) -> column-sum kernel
                                          (256, 768, 256, 3)])    # 3-plane mode -> column-sum kernel
def test_gemm_mul_epilogue_column_sums(M, N, K, planes):
    """EPI_MUL + outf: the FFN bias gradient db1 = colsum((dz2 W2) o gelu') from the dgrad GEMM's epilogue"""
    err, _, cs = run_gemm(M, N, K, 0, planes, _lib.EPI_MUL, 1, seed=7, colsum=True)
    assert err < (1e-5 if planes == 3 else 6e-3)
    assert cs < 2e-5, cs


def attention_reference(qkv, mask, dctx, heads):
    """softmax(q k^T / 8 + key mask) v per head, and its gradients, in float64 (LEAD/modeling_bert.py:324-368)"""
    B, S, H3 = qkv.shape
    H = H3 // 3
    x = qkv.double().requires_grad_(True)
    q, k, v = (x[..., i * H:(i + 1) * H].reshape(B, S, heads, 64).transpose(1, 2) for i in range(3))
    s = q @ k.transpose(-1, -2) / 8.0
    if mask is not None:
        s = s + (1.0 - mask.double())[:, None, None, :] * -1e30
    ctx = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B, S, H)
    (ctx * dctx.double()).sum().backward()
    return ctx.detach(), x.grad


@pytest.mark.parametrize("B,S,heads,ragged", [(3, 128, 2, False), (5, 128, 12, True), (4, 32, 3, True), (2, 77, 2, True),
                                              (2, 160, 2, True), (3, 200, 2, True), (2, 256, 3, False), (1, 512, 1, True)])
def test_attention_forward_backward_and_bias_gradient(B, S, heads, ragged):
    lib = _lib.lib()
    H = heads * 64
    g = torch.Generator(device="cuda").manual_seed(S + heads)
    qkv = torch.randn(B, S, 3 * H, device="cuda", generator=g).to(torch.bfloat16)
    dctx = torch.randn(B, S, H, device="cuda", generator=g).to(torch.bfloat16)
    mask = None
    if ragged:
        lens = torch.randint(1, S + 1, (B,), generator=torch.Generator().manual_seed(S))
        mask = (torch.arange(S)[None, :] < lens[:, None]).to(torch.int64).cuda()
    ctx = torch.empty(B, S, H, device="cuda", dtype=torch.bfloat16)
    dqkv = torch.full((B, S, 3 * H), float("nan"), device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(B, heads, S, device="cuda")
    delta = torch.empty(B, heads, S, device="cuda")
    dbias = torch.full((3 * H,), 0.5, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.simx_attention_forward(qkv.data_ptr(), ctx.data_ptr(), _lib.ptr(mask), lse.data_ptr(), B, S, H, heads,
                                          1, 0, 0, st), "simx_attention_forward")
    _lib.check(lib.simx_attention_backward(qkv.data_ptr(), ctx.data_ptr(), dctx.data_ptr(), dqkv.data_ptr(),
                                           _lib.ptr(mask), lse.data_ptr(), delta.data_ptr(), B, S, H, heads, 1, 0, 0,
                                           dbias.data_ptr(), st), "simx_attention_backward")
    torch.cuda.synchronize()
    ctx_ref, dqkv_ref = attention_reference(qkv.float(), mask, dctx.float(), heads)
    assert ((ctx.double() - ctx_ref).abs().max() / ctx_ref.abs().max()).item() < 1e-2
    assert not torch.isnan(dqkv.float()).any()
    # the backward consumes the bf16-rounded ctx it was given (delta = rowsum(dO o O)): 2e-2 of the tensor scale
    assert ((dqkv.double() - dqkv_ref).abs().max() / dqkv_ref.abs().max()).item() < 2e-2
    # bias gradient = column sums of dqkv over all B*S rows, accumulated into the caller's buffer.  S <= 128 and the
    # 3-plane path: sums of exactly the stored values.  S > 128 (ceil(S/128)^2 tcgen05 blocks): sums of the blocks' bf16
    # partial tiles, which the stored (reduce-added, re-rounded) dqkv matches only to bf16 rounding.
    want = 0.5 + dqkv.float().double().sum((0, 1))
    tol = 5e-3 if S > 128 else 2e-5
    assert ((dbias.double() - want).abs().max() / want.abs().max()).item() < tol
    want_exact = 0.5 + dqkv_ref.sum((0, 1))
    assert ((dbias.double() - want_exact).abs().max() / want_exact.abs().max()).item() < 2e-2
