"""Kernel-level parity on the B200, every call going through the C ABI (libbimodal_b200.so).

Checkers: the oracle's restatement of the reference functions run on CUDA tensors (bit-exact
criteria for the integer / fp32 paths) and plain fp32 torch for the floating-point kernels, with
the tolerance written next to each assert.
"""
import ctypes as C
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _stream():
    return torch.cuda.current_stream().cuda_stream


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(1, 3, 336, 336), (1, 3, 224, 224), (1, 3, 28, 28), (1, 3, 7, 5)])
def test_pgd_step_bit_exact(lib, cuda_device, shape):
    from bimodalattack_b200 import native
    from oracle import hotpath
    g = torch.Generator(device="cpu").manual_seed(3)
    img = torch.rand(shape, generator=g).to(cuda_device)
    orig = (img + 0.2 * (torch.rand(shape, generator=g).to(cuda_device) - 0.5)).clamp(0, 1)
    grad = torch.randn(shape, generator=g).to(cuda_device)
    grad.view(-1)[::7] = 0.0          # sign(0) = 0
    grad.view(-1)[3] = float("nan")   # sign(nan) = 0 in torch
    grad.view(-1)[5] = -0.0
    eps, alpha = 64 / 255, 4 / 255
    want = hotpath.pgd_step(img.clone(), eps, alpha, grad, orig).detach()
    got = img.clone()
    native.pgd_step_(got, grad, orig, alpha * eps, eps)
    assert torch.equal(got, want)     # fp32, bit exact


def test_pgd_step_fused_normalize(lib, cuda_device):
    import torchvision.transforms as T
    from bimodalattack_b200 import native, synthetic
    g = torch.Generator(device="cpu").manual_seed(4)
    shape = (1, 3, 336, 336)
    img = torch.rand(shape, generator=g).to(cuda_device)
    orig = img.clone()
    grad = torch.randn(shape, generator=g).to(cuda_device)
    got = img.clone()
    norm = native.pgd_step_(got, grad, orig, 0.01, 0.1, synthetic.CLIP_MEAN, synthetic.CLIP_STD)
    want = T.Normalize(synthetic.CLIP_MEAN, synthetic.CLIP_STD)(got)
    assert torch.equal(norm, want)    # same fp32 subtract and divide


# ------------------------------------------------------------------------------------------------
def test_torch_cuda_philox_contract(cuda_device):
    """The RNG contract the sampler relies on: offset advances by 4 per small draw."""
    torch.zeros(1, device=cuda_device)   # default_generators is filled lazily on CUDA init
    torch.manual_seed(1234)
    gen = torch.cuda.default_generators[cuda_device.index]
    o0 = gen.get_offset()
    torch.rand((512, 20), device=cuda_device)
    o1 = gen.get_offset()
    torch.randint(0, 256, (512, 1, 1), device=cuda_device)
    o2 = gen.get_offset()
    assert (o1 - o0, o2 - o1) == (4, 4)


@pytest.mark.parametrize("V,dtype,sw,ntok,topk,nrep,masked", [
    (32064, torch.float32, 512, 20, 256, 1, False),
    (32064, torch.float32, 512, 20, 256, 1, True),
    (32064, torch.float32, 4096, 20, 256, 2, True),
    (262208, torch.float32, 512, 20, 256, 1, True),
    (1000, torch.float32, 64, 8, 32, 1, False),
    (1000, torch.float32, 33, 5, 1000, 3, False),
])
def test_sample_ids_bit_exact(lib, cuda_device, V, dtype, sw, ntok, topk, nrep, masked):
    """Tie-free gradients + identical (seed, offset): candidate ids equal the reference's
    sample_ids_from_grad run on CUDA, and torch's generator ends at the same offset."""
    from bimodalattack_b200 import native
    from oracle import hotpath
    g = torch.Generator(device="cpu").manual_seed(V + sw)
    # distinct values per row: a random permutation scaled into a sane range
    grad = torch.stack([torch.randperm(V, generator=g) for _ in range(ntok)]).to(torch.float32)
    grad = ((grad - V / 2) / V).to(cuda_device).to(dtype)
    ids = torch.randint(5, V, (ntok,), generator=g).to(cuda_device)
    not_allowed = None
    if masked:
        not_allowed = torch.randperm(V, generator=g)[: int(0.3 * V)].to(cuda_device)
    gen = torch.cuda.default_generators[cuda_device.index]

    torch.manual_seed(77)
    torch.rand(8, device=cuda_device)            # move the offset off zero
    seed, off = gen.initial_seed(), gen.get_offset()
    want = hotpath.sample_ids_from_grad(ids.clone(), grad.clone(), sw, topk, nrep, not_allowed)
    off_after = gen.get_offset()

    mask = native.pack_mask_bits(not_allowed, V, cuda_device)
    got, adv, topk_ids = native.sample_ids(ids, grad, sw, topk, nrep, mask, seed, off, want_topk=True)
    torch.cuda.synchronize()
    gm = grad.clone()
    if not_allowed is not None:
        gm[:, not_allowed] = float("inf")
    want_topk = (-gm).topk(topk, dim=1).indices
    assert torch.equal(topk_ids, want_topk)
    assert off + adv == off_after
    assert torch.equal(got, want)


@pytest.mark.parametrize("V", [32064, 262208])
def test_topk_bf16_with_ties_is_a_valid_topk(lib, cuda_device, V):
    """bf16 gradients are full of ties, where torch.topk's order is unspecified: require the same
    multiset of selected gradient values and a non-increasing order of -grad."""
    from bimodalattack_b200 import native
    g = torch.Generator(device="cpu").manual_seed(9)
    grad = (torch.randn((20, V), generator=g) * 1e-3).to(cuda_device).to(torch.bfloat16)
    ids = torch.randint(5, V, (20,), generator=g).to(cuda_device)
    _, _, topk_ids = native.sample_ids(ids, grad, 8, 256, 1, None, 1, 0, want_topk=True)
    torch.cuda.synchronize()
    picked = torch.gather(-grad.float(), 1, topk_ids)
    want_vals = (-grad.float()).topk(256, dim=1).values
    assert torch.equal(picked, want_vals)
    for r in range(20):
        assert topk_ids[r].unique().numel() == 256


# ------------------------------------------------------------------------------------------------
def _gemm(lib, A, B, res=None, block_n=256):
    M, K = A.shape
    N = B.shape[0]
    Cm = torch.empty((M, N), dtype=torch.bfloat16, device=A.device)
    from bimodalattack_b200 import native
    native.check(lib.bma_gemm_bf16(A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0), Cm.data_ptr(),
                                   Cm.stride(0), None if res is None else res.data_ptr(),
                                   0 if res is None else res.stride(0), M, N, K, block_n, _stream()), "gemm")
    torch.cuda.synchronize()
    return Cm


@pytest.mark.parametrize("M,N,K,bn", [
    (128, 256, 64, 256), (128, 256, 256, 256), (256, 512, 4096, 256), (300, 776, 1088, 256),
    (42, 4096, 4096, 256), (1000, 1024, 512, 128), (4096, 4096, 11008, 256), (21504, 4096, 4096, 256),
    # odd number of 128-row blocks (the second CTA of the last pair is all padding), ragged row count
    (385, 512, 512, 256), (16157, 2560, 2560, 256), (129, 256, 64, 256),
])
def test_gemm_tcgen05(lib, cuda_device, M, N, K, bn):
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    A = (torch.randn((M, K), generator=g) * 0.5).to(cuda_device).to(torch.bfloat16)
    B = (torch.randn((N, K), generator=g) * 0.5).to(cuda_device).to(torch.bfloat16)
    got = _gemm(lib, A, B, block_n=bn).float()
    want = A.float() @ B.float().T
    # fp32 accumulation of exact bf16 products, one rounding to bf16: <= 1 bf16 ulp of the result
    # plus accumulation-order noise
    err = (got - want).abs()
    tol = want.abs() * 2 ** -8 + 1e-2 * math.sqrt(K) * 2 ** -8
    assert bool((err <= tol).all()), f"max err {err.max().item()} at {err.argmax().item()}"


def test_gemm_residual(lib, cuda_device):
    g = torch.Generator(device="cpu").manual_seed(2)
    A = torch.randn((256, 512), generator=g).to(cuda_device).to(torch.bfloat16)
    B = (torch.randn((512, 512), generator=g) * 0.1).to(cuda_device).to(torch.bfloat16)
    R = torch.randn((256, 512), generator=g).to(cuda_device).to(torch.bfloat16)
    got = _gemm(lib, A, B, res=R).float()
    want = ((A.float() @ B.float().T).to(torch.bfloat16).float() + R.float()).to(torch.bfloat16).float()
    assert (got - want).abs().max().item() <= 2 ** -6 * want.abs().max().item()


@pytest.mark.parametrize("ntok,d,V", [(20, 4096, 32064), (20, 2560, 262208), (1, 256, 1000), (40, 512, 4096)])
def test_token_grad(lib, cuda_device, ntok, d, V):
    from bimodalattack_b200 import native
    g = torch.Generator(device="cpu").manual_seed(V)
    G = (torch.randn((ntok, d), generator=g) * 1e-2).to(cuda_device).to(torch.bfloat16)
    W = (torch.randn((V, d), generator=g) * 0.02).to(cuda_device).to(torch.bfloat16)
    got = native.token_grad(G, W).float()
    torch.cuda.synchronize()
    want = G.float() @ W.float().T
    err = (got - want).abs()
    tol = want.abs() * 2 ** -8 + 1e-6
    assert bool((err <= tol).all()), f"max err {err.max().item()}"


# ------------------------------------------------------------------------------------------------
def test_rmsnorm(lib, cuda_device):
    from bimodalattack_b200 import native
    g = torch.Generator(device="cpu").manual_seed(5)
    x = torch.randn((333, 4096), generator=g).to(cuda_device).to(torch.bfloat16)
    w = (1 + 0.1 * torch.randn((4096,), generator=g)).to(cuda_device).to(torch.bfloat16)
    y = torch.empty_like(x)
    native.check(lib.bma_rmsnorm(x.data_ptr(), x.stride(0), w.data_ptr(), y.data_ptr(), y.stride(0),
                                 333, 4096, 1e-5, 0, _stream()), "rmsnorm")
    torch.cuda.synchronize()
    xf = x.float()
    ref = w * (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-5)).to(torch.bfloat16)   # HF LlamaRMSNorm
    # same rounding points; the fp32 mean may differ in the last bit -> at most 1 bf16 ulp
    assert (y.float() - ref.float()).abs().max().item() <= 2 ** -7 * ref.float().abs().max().item()


def _attn_reference(q, kt, vt, kp, vp, S, scale):
    """fp32: every query row sees all prefix keys and the causal part of its own candidate."""
    M, H, D = q.shape
    P = kp.shape[0]
    out = torch.empty((M, H, D), dtype=torch.float32, device=q.device)
    B = M // S
    qf, ktf, vtf, kpf, vpf = (t.float() for t in (q, kt, vt, kp, vp))
    causal = torch.tril(torch.ones((S, S), dtype=torch.bool, device=q.device))
    for b in range(B):
        sl = slice(b * S, (b + 1) * S)
        k = torch.cat([kpf, ktf[sl]], 0)            # [P+S, H, D]
        v = torch.cat([vpf, vtf[sl]], 0)
        s = torch.einsum("shd,khd->hsk", qf[sl], k) * scale
        mask = torch.cat([torch.ones((S, P), dtype=torch.bool, device=q.device), causal], 1)
        s = s.masked_fill(~mask, float("-inf"))
        p = torch.softmax(s, dim=-1)
        out[sl] = torch.einsum("hsk,khd->shd", p, v)
    return out


@pytest.mark.parametrize("B,S,P,H,G,D", [
    (7, 42, 596, 4, 1, 128), (3, 42, 0, 2, 1, 128), (1, 596, 0, 2, 1, 128), (5, 130, 70, 2, 1, 128),
    (64, 42, 24, 2, 1, 128), (512, 41, 596, 1, 1, 128), (6, 41, 300, 4, 2, 128), (2, 300, 17, 2, 1, 128),
    # head_dim 256 (Gemma-3: 64-key blocks, 256-column O accumulator), GQA 2
    (7, 41, 20, 4, 2, 256), (3, 300, 4, 2, 2, 256), (1, 70, 0, 2, 1, 256), (40, 41, 200, 2, 2, 256),
    # head_dim 64 (the tiny CPU-sized model of BASELINE.json config 1)
    (5, 42, 24, 4, 1, 64), (3, 130, 0, 2, 2, 64)])
def test_prefix_tail_attention(lib, cuda_device, B, S, P, H, G, D):
    """H query heads, H // G kv heads (G = GQA group). Values are handed over transposed
    ([kv_heads*D, rows]) the way the fused QKV epilogue writes them."""
    from bimodalattack_b200 import native
    Hkv = H // G
    g = torch.Generator(device="cpu").manual_seed(B * S + P)
    M = B * S
    q = torch.randn((M, H, D), generator=g).to(cuda_device).to(torch.bfloat16)
    kt = torch.randn((M, Hkv, D), generator=g).to(cuda_device).to(torch.bfloat16)
    vt = torch.randn((M, Hkv, D), generator=g).to(cuda_device).to(torch.bfloat16)
    kp = torch.randn((max(P, 1), Hkv, D), generator=g).to(cuda_device).to(torch.bfloat16)[:P]
    vp = torch.randn((max(P, 1), Hkv, D), generator=g).to(cuda_device).to(torch.bfloat16)[:P]
    ld_t = (M + 7) // 8 * 8 + 8
    ld_p = (max(P, 1) + 7) // 8 * 8
    vt_T = torch.full((Hkv * D, ld_t), float("nan"), dtype=torch.bfloat16, device=cuda_device)   # padding must not leak
    vt_T[:, :M] = vt.reshape(M, Hkv * D).T
    vp_T = torch.zeros((Hkv * D, ld_p), dtype=torch.bfloat16, device=cuda_device)
    vp_T[:, :P] = vp.reshape(P, Hkv * D).T
    kp_buf = torch.zeros((max(P, 1), Hkv * D), dtype=torch.bfloat16, device=cuda_device)
    kp_buf[:P] = kp.reshape(P, Hkv * D)
    out = torch.zeros((M, H * D), dtype=torch.bfloat16, device=cuda_device)
    scale = D ** -0.5
    q2, k2 = q.reshape(M, H * D), kt.reshape(M, Hkv * D)
    native.check(lib.bma_attention(q2.data_ptr(), q2.stride(0), k2.data_ptr(), k2.stride(0),
                                   vt_T.data_ptr(), vt_T.stride(0), kp_buf.data_ptr(), kp_buf.stride(0),
                                   vp_T.data_ptr(), vp_T.stride(0), out.data_ptr(), out.stride(0),
                                   M, S, P, H, G, D, scale, _stream()), "attention")
    torch.cuda.synchronize()
    rep = lambda t: t.repeat_interleave(G, dim=1)
    want = _attn_reference(q, rep(kt), rep(vt), rep(kp), rep(vp), S, scale).reshape(M, H * D)
    assert not bool(torch.isnan(out.float()).any())
    # bf16 P and bf16 output: 2^-7 relative to the output scale (|out| <~ 1 for unit-variance V)
    assert (out.float() - want).abs().max().item() <= 2.5e-2


@pytest.mark.parametrize("L,H,D", [(577, 16, 64), (257, 4, 64), (50, 2, 64), (577, 4, 128), (130, 2, 128), (64, 1, 128)])
def test_attention_no_tail_is_plain_self_attention(lib, cuda_device, L, H, D):
    """S = 0: every row attends to the P "prefix" keys only. With the sequence's own K / V^T handed in as
    the prefix (inside a row-major Q | K | V buffer, the way the vision tower calls it) that is plain
    non-causal self-attention (reference modeling_clip eager attention: softmax(Q K^T / sqrt(d)) V)."""
    from bimodalattack_b200 import native
    g = torch.Generator(device="cpu").manual_seed(L * H + D)
    qkv = torch.randn((L, 3 * H * D), generator=g).to(cuda_device).to(torch.bfloat16)
    Lp = (L + 7) // 8 * 8
    vT = torch.full((H * D, Lp), float("nan"), dtype=torch.bfloat16, device=cuda_device)   # padding must not leak
    vT[:, :L] = qkv[:, 2 * H * D:].T
    out = torch.zeros((L, H * D), dtype=torch.bfloat16, device=cuda_device)
    kp = qkv.data_ptr() + 2 * H * D
    scale = D ** -0.5
    native.check(lib.bma_attention(qkv.data_ptr(), qkv.stride(0), kp, qkv.stride(0), vT.data_ptr(), Lp,
                                   kp, qkv.stride(0), vT.data_ptr(), Lp, out.data_ptr(), out.stride(0),
                                   L, 0, L, H, 1, D, scale, _stream()), "attention(no tail)")
    torch.cuda.synchronize()
    q, k, v = (qkv[:, i * H * D:(i + 1) * H * D].float().view(L, H, D) for i in range(3))
    p_ = torch.softmax(torch.einsum("qhd,khd->hqk", q, k) * scale, dim=-1)
    want = torch.einsum("hqk,khd->qhd", p_, v).reshape(L, H * D)
    assert not bool(torch.isnan(out.float()).any())
    assert (out.float() - want).abs().max().item() <= 2.5e-2
