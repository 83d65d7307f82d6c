"""Kernel-level checks of the fused GEMM epilogues (one launch of the production tcgen05 kernel each, through
vavae_test_gemm_epilogue) against an fp32 torch reference of the same op on the SAME bf16-rounded operands, so a
regression points at the epilogue instead of surfacing as "forward rel-L2 7e-3":
  * QKV: + bias, optional qk RMSNorm, RoPE, q/k scattered head-major, V stored transposed with its ones row
    (models/lightningdit.py:66-76, models/pos_embed.py:38-42,135) -- head_dim 64 and 72;
  * SwiGLU on row-interleaved w12 (models/swiglu_ffn.py:31-36);
  * gated residual (models/lightningdit.py:248-249) + the next RMSNorm's operand and per-row partial sums;
  * the consumer side of the fused RMSNorm (1/rms row scale + per-sample shift.W^T + bias vector).
Operands are bf16, accumulation fp32: the only differences from the reference are the accumulation order and the
final bf16 rounding of bf16 outputs, hence bars of 2e-3 (bf16 outputs, 1 ulp = 2^-8 relative) and 1e-5 (fp32)."""
import pytest
import torch

from conftest import rel_l2
from oracle import lightningdit_oracle as O

pytestmark = pytest.mark.gpu


def _dev():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device (run with -m gpu on the B200 box)")
    return torch.device("cuda", 0)


def _rand(*shape, seed, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


def _operands(M, N, K, seed):
    A = _rand(M, K, seed=seed).bfloat16()
    W = _rand(N, K, seed=seed + 1, scale=K ** -0.5).bfloat16()
    bias = _rand(N, seed=seed + 2, scale=0.1)
    return A, W, bias


@pytest.mark.parametrize("H,hd,qknorm", [(6, 64, False), (4, 72, False), (4, 72, True), (6, 64, True)])
@pytest.mark.parametrize("B", [2, 5])
def test_qkv_epilogue(H, hd, qknorm, B):
    from va_vae_b200 import engine as E
    dev = _dev()
    T, D = 256, H * hd
    M, N, K = B * T, 3 * D, D
    A, W, bias = _operands(M, N, K, seed=10 + hd)
    cos, sin = O.rope_tables(hd, 16)
    qn_w = (1 + 0.1 * _rand(hd, seed=3)) if qknorm else None
    kn_w = (1 + 0.1 * _rand(hd, seed=4)) if qknorm else None
    # fp32 reference on the bf16 operands (models/lightningdit.py:68-76)
    qkv = (A.float() @ W.float().t() + bias).reshape(B, T, 3, H, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    if qknorm:
        q, k = O.rms_norm(q, qn_w), O.rms_norm(k, kn_w)
    q = q * cos + O.rotate_half(q) * sin
    k = k * cos + O.rotate_half(k) * sin
    hdp, hdv = (hd + 15) // 16 * 16, E.attention_hdv(hd)
    qo = torch.zeros(B, H, T, hdp, dtype=torch.bfloat16, device=dev)
    ko = torch.zeros_like(qo)
    vt = torch.zeros(B, H, hdv, T, dtype=torch.bfloat16, device=dev)
    vt[:, :, hd, :] = 1.0                                  # the ones row the engine writes once at finalize
    kw = dict(kind=E.EPI_QKV, M=M, N=N, K=K, T=T, A=A.to(dev), W=W.to(dev), bias=bias.to(dev), H=H, hd=hd,
              cos=cos.to(dev).contiguous(), sin=sin.to(dev).contiguous(), q=qo, k=ko, vt=vt)
    if qknorm:
        kw.update(qn_w=qn_w.to(dev), kn_w=kn_w.to(dev))
    E.test_gemm_epilogue(**kw)
    torch.cuda.synchronize()
    assert rel_l2(qo[..., :hd], q) < 2e-3 and rel_l2(ko[..., :hd], k) < 2e-3
    assert rel_l2(vt[:, :, :hd, :], v.transpose(-1, -2)) < 2e-3
    assert torch.all(qo[..., hd:] == 0) and torch.all(ko[..., hd:] == 0)            # K padding stays zero
    assert torch.all(vt[:, :, hd, :] == 1) and torch.all(vt[:, :, hd + 1:, :] == 0)  # ones row / padding untouched
    # element-wise: every output within 1.5 bf16 ulp of the fp32 reference (catches a mis-scattered head or token,
    # which a norm-wise bound over the whole tensor can hide)
    for got, want in ((qo[..., :hd], q), (ko[..., :hd], k), (vt[:, :, :hd, :], v.transpose(-1, -2))):
        diff = (got.float().cpu() - want).abs()
        assert float((diff - 1.5 * 2.0 ** -8 * want.abs()).max()) < 2e-2


@pytest.mark.parametrize("fused_norm", [False, True])
@pytest.mark.parametrize("S", [1024, 170])
def test_swiglu_epilogue(fused_norm, S):
    from va_vae_b200 import engine as E
    dev = _dev()
    T, Bn, K = 256, 3, 384
    M = Bn * T
    N = (2 * S + 3) // 4 * 4
    A, W12, b12 = _operands(M, 2 * S, K, seed=40 + S)
    ref_in = A.float()
    bias_rows = b12.unsqueeze(0).expand(Bn, -1)
    rstd = torch.ones(M)
    kw = {}
    if fused_norm:
        # producer contract: A = un-normalised x * w * (1 + scale), partial sums of x^2 in 5 slots, per-sample vector sp
        part = torch.zeros(M, 16)
        part[:, :5] = _rand(M, 5, seed=7).abs() * K / 5
        rstd = torch.rsqrt(part.sum(1) / K + 1e-6)
        bias_rows = _rand(Bn, 2 * S, seed=8, scale=0.2)
        kw = dict(in_part=part.to(dev), in_nparts=5, in_dim=K, sp_ld=N, part_ld=16)
    U = (ref_in @ W12.float().t()) * rstd[:, None] + bias_rows.repeat_interleave(T, 0)
    want = torch.nn.functional.silu(U[:, :S]) * U[:, S:]                               # swiglu_ffn.py:33-35
    # engine layout: rows (x1_j, x2_j) interleaved, N padded to a multiple of 4 with zero rows
    il = torch.stack([torch.arange(S), torch.arange(S) + S], 1).reshape(-1)
    Wi = torch.zeros(N, K, dtype=torch.bfloat16)
    Wi[: 2 * S] = W12[il]
    bi = torch.zeros(N)
    bi[: 2 * S] = b12[il]
    if fused_norm:
        sp = torch.zeros(Bn, N)
        sp[:, : 2 * S] = bias_rows[:, il]
        kw["sp"] = sp.to(dev)
    Spad = (S + 7) // 8 * 8
    out = torch.zeros(M, Spad, dtype=torch.bfloat16, device=dev)
    E.test_gemm_epilogue(kind=E.EPI_SWIGLU, M=M, N=N, K=K, T=T, A=A.to(dev), W=Wi.to(dev), bias=bi.to(dev),
                         out_bf16=out, ldo=Spad, **kw)
    torch.cuda.synchronize()
    assert rel_l2(out[:, :S], want) < 3e-3
    assert torch.all(out[:, S:] == 0)


@pytest.mark.parametrize("N,K", [(384, 384), (1152, 3072), (768, 2048)])
@pytest.mark.parametrize("next_norm", [False, True])
@pytest.mark.parametrize("T,Bn", [(256, 3), (256, 4), (160, 3)])
def test_gated_residual_epilogue(N, K, next_norm, T, Bn):
    """M = 768 / 1024 rows: whole 256-row tile pairs -> the TMA epilogue (residual tiles moved by cp.async.bulk.tensor);
    M = 480: partial tiles -> the register-prefetch fast path plus the guarded path on the edge tile."""
    from va_vae_b200 import engine as E
    dev = _dev()
    M = Bn * T
    A, W, bias = _operands(M, N, K, seed=70 + N)
    X0 = _rand(M, N, seed=1)
    mod_ld, gate_off, scale_off = 3 * N + 8, 8, 8 + 2 * N
    mod = _rand(Bn, mod_ld, seed=2, scale=0.5)
    gate = mod[:, gate_off: gate_off + N].repeat_interleave(T, 0)
    want = X0 + gate * (A.float() @ W.float().t() + bias)                               # lightningdit.py:248-249
    X = X0.clone().to(dev)
    kw = {}
    if next_norm:
        nx_w = 1 + 0.1 * _rand(N, seed=3)
        nxA = torch.zeros(M, N, dtype=torch.bfloat16, device=dev)
        PL = max(16, N // 4)
        part = torch.full((M, PL), float("nan"), device=dev)
        kw = dict(nx_w=nx_w.to(dev), nx_scale_off=scale_off, nx_A=nxA, nx_part=part, part_ld=PL)
    nparts = E.test_gemm_epilogue(kind=E.EPI_GATED_RESIDUAL, M=M, N=N, K=K, T=T, A=A.to(dev), W=W.to(dev),
                                  bias=bias.to(dev), X=X, mod=mod.to(dev), mod_ld=mod_ld, gate_off=gate_off, **kw)
    torch.cuda.synchronize()
    assert rel_l2(X, want) < 1e-5
    if next_norm:
        scale = mod[:, scale_off: scale_off + N].repeat_interleave(T, 0)
        assert rel_l2(nxA, want * nx_w * (1 + scale)) < 3e-3                           # un-normalised operand of the next GEMM
        fine = E.test_gemm_epilogue.fine_out
        # wide tiles hold whole 128-column groups -> one coarse slot per group; the narrow tiles of a small-M launch
        # write eight fine slots per 32-column chunk instead
        nslots = nparts * 8 if fine else nparts
        assert nparts == (N // 32 if fine else (N + 127) // 128)
        assert not torch.isnan(part[:, :nslots]).any()
        total = _fold_fine(part, nparts).sum(1) if fine else part[:, :nparts].sum(1)
        assert rel_l2(total, (want * want).sum(1)) < 1e-5                              # what the consumer turns into 1/rms
        assert torch.isnan(part[:, nslots:]).all()                                     # nothing beyond is touched
        # the slots are a pure function of the row: a second launch on the same inputs is bit-identical
        X2, part2 = X0.clone().to(dev), torch.zeros(M, PL, device=dev)
        kw2 = dict(kw, nx_part=part2, nx_A=torch.zeros_like(nxA))
        E.test_gemm_epilogue(kind=E.EPI_GATED_RESIDUAL, M=M, N=N, K=K, T=T, A=A.to(dev), W=W.to(dev), bias=bias.to(dev),
                             X=X2, mod=mod.to(dev), mod_ld=mod_ld, gate_off=gate_off, **kw2)
        assert torch.equal(part2[:, :nslots], part[:, :nslots]) and torch.equal(X2, X)


def _fold_fine(fine, nchunks):
    """Coarse (128-column) slots from fine ones ([M, nchunks * 8]: s(c, j) = sum of squares of columns 32c + 4j .. + 3)
    in the kernels' order: fold8_j( ((s(4g, j) + s(4g+1, j)) + s(4g+2, j)) + s(4g+3, j) )."""
    ng = (nchunks + 3) // 4
    out = torch.zeros(fine.shape[0], ng, dtype=torch.float32, device=fine.device)
    s = fine[:, : nchunks * 8].reshape(fine.shape[0], nchunks, 8)
    for g in range(ng):
        t = s[:, 4 * g].clone()
        for c in range(4 * g + 1, min(4 * g + 4, nchunks)):
            t = t + s[:, c]
        out[:, g] = ((t[:, 0] + t[:, 4]) + (t[:, 2] + t[:, 6])) + ((t[:, 1] + t[:, 5]) + (t[:, 3] + t[:, 7]))
    return out


@pytest.mark.parametrize("N,K", [(384, 384), (1152, 3072), (1152, 1152), (96, 384)])
def test_gated_residual_paths_agree_bitwise(N, K):
    """The same rows through the TMA epilogue (M = 768: whole 256-row tile pairs) and through the register-prefetch /
    guarded epilogues (M = 480): residual stream, next operand and partial sums must be bit-identical, or results would
    depend on the batch size.  N = 96 is one narrow tile: fine slots on both paths."""
    from va_vae_b200 import engine as E
    dev = _dev()
    T = 32
    PL = max(16, N // 4)
    res = {}
    for M in (768, 480):
        Bn = M // T
        A, W, bias = _operands(768, N, K, seed=91)
        X0 = _rand(768, N, seed=5)
        mod_ld, gate_off, scale_off = 3 * N + 8, 8, 8 + 2 * N
        mod = _rand(768 // T, mod_ld, seed=6, scale=0.5)
        nx_w = 1 + 0.1 * _rand(N, seed=3)
        X = X0[:M].clone().to(dev)
        nxA = torch.zeros(M, N, dtype=torch.bfloat16, device=dev)
        part = torch.zeros(M, PL, device=dev)
        nparts = E.test_gemm_epilogue(kind=E.EPI_GATED_RESIDUAL, M=M, N=N, K=K, T=T, A=A[:M].contiguous().to(dev),
                                      W=W.to(dev), bias=bias.to(dev), X=X, mod=mod[:Bn].contiguous().to(dev), mod_ld=mod_ld,
                                      gate_off=gate_off, nx_w=nx_w.to(dev), nx_scale_off=scale_off, nx_A=nxA, nx_part=part,
                                      part_ld=PL)
        torch.cuda.synchronize()
        res[M] = (X, nxA, part, nparts, E.test_gemm_epilogue.fine_out)
    (Xa, Na, Pa, na, fa), (Xb, Nb, Pb, nb, fb) = res[768], res[480]
    assert fa == fb and na == nb == (N // 32 if fa else (N + 127) // 128)   # (both launches are small: same tile width)
    assert torch.equal(Xa[:480], Xb) and torch.equal(Na[:480], Nb)
    width = na * 8 if fa else na
    assert torch.equal(Pa[:480, :width], Pb[:, :width])
    want_ss = (Xb.double() ** 2).sum(1)
    got = _fold_fine(Pb, nb).sum(1) if fb else Pb[:, :nb].sum(1)
    assert rel_l2(got.double(), want_ss) < 1e-5


def test_fused_norm_consumer_fine_and_coarse_slots_agree_bitwise():
    """A consumer (mlp.w12 here) fed fine slots folds them into the groups a coarse-slot producer would have written."""
    from va_vae_b200 import engine as E
    dev = _dev()
    T, Bn, K, S = 256, 3, 384, 512
    M, N = Bn * T, 2 * S
    A, W12, b12 = _operands(M, N, K, seed=33)
    nch, ng = K // 32, (K + 127) // 128
    fine = _rand(M, nch * 8, seed=7).abs() * 4
    coarse = _fold_fine(fine, nch)
    sp = _rand(Bn, N, seed=8, scale=0.2)
    outs = []
    for part, nparts, is_fine, pl in ((fine, nch, 1, nch * 8 + 8), (coarse, ng, 0, 16)):
        out = torch.zeros(M, S, dtype=torch.bfloat16, device=dev)
        p = torch.full((M, pl), 1e30)                      # stale values behind the valid slots must be ignored
        p[:, : part.shape[1]] = part
        E.test_gemm_epilogue(kind=E.EPI_SWIGLU, M=M, N=N, K=K, T=T, A=A.to(dev), W=W12.to(dev), bias=b12.to(dev),
                             out_bf16=out, ldo=S, in_part=p.to(dev), in_nparts=nparts, in_fine=is_fine, in_dim=K,
                             sp=sp.to(dev), sp_ld=N, part_ld=pl)
        torch.cuda.synchronize()
        outs.append(out)
    assert torch.equal(outs[0], outs[1])
    rstd = torch.rsqrt(coarse.sum(1) / K + 1e-6)
    U = (A.float() @ W12.float().t()) * rstd[:, None] + sp.repeat_interleave(T, 0)
    want = torch.nn.functional.silu(U[:, 0::2]) * U[:, 1::2]
    assert rel_l2(outs[0], want) < 3e-3


def test_gated_residual_tma_epilogue_single_cta_matches_cta_pairs():
    """The TMA epilogue also runs under the single-CTA kernel (developer knob set_gemm_two_cta(0)): same bits as CTA pairs."""
    from va_vae_b200 import engine as E
    dev = _dev()
    T, M, N, K = 256, 768, 1152, 1152
    A, W, bias = _operands(M, N, K, seed=17)
    X0 = _rand(M, N, seed=1)
    mod_ld, gate_off, scale_off = 3 * N + 8, 8, 8 + 2 * N
    mod = _rand(M // T, mod_ld, seed=2, scale=0.5)
    nx_w = 1 + 0.1 * _rand(N, seed=3)
    res = []
    try:
        for two in (1, 0):
            E.set_gemm_two_cta(two)
            X = X0.clone().to(dev)
            nxA = torch.zeros(M, N, dtype=torch.bfloat16, device=dev)
            part = torch.zeros(M, 16, device=dev)
            n = E.test_gemm_epilogue(kind=E.EPI_GATED_RESIDUAL, M=M, N=N, K=K, T=T, A=A.to(dev), W=W.to(dev),
                                     bias=bias.to(dev), X=X, mod=mod.to(dev), mod_ld=mod_ld, gate_off=gate_off,
                                     nx_w=nx_w.to(dev), nx_scale_off=scale_off, nx_A=nxA, nx_part=part, part_ld=16)
            torch.cuda.synchronize()
            res.append((X, nxA, part[:, :n]))
    finally:
        E.set_gemm_two_cta(2)
    gate = mod[:, gate_off: gate_off + N].repeat_interleave(T, 0)
    want = X0 + gate * (A.float() @ W.float().t() + bias)
    assert rel_l2(res[0][0], want) < 1e-5
    for a, b in zip(res[0], res[1]):
        assert torch.equal(a, b)
