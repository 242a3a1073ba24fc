"""Native gradient step (bimodalattack_b200/gradient.py + csrc/phase_a.cu) against PyTorch: every
row kernel and GEMM mode against the same op in fp32 torch, then the whole forward + input-gradient
backward against autograd through the HF model (the reference's compute_gradient, :953-1028).

Tolerances: single ops are compared with an fp32 torch evaluation of the SAME bf16 inputs, to 1 bf16
ulp of the row scale (outputs are bf16). Whole-model gradients are compared with an fp32 autograd run
on the same bf16-valued weights; the reference's own bf16 autograd path is measured against that
fp32 truth in the same test and the native path must be no farther from it than 1.5x the reference."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def _k(dev):
    from bimodalattack_b200.gradient import _Lib
    return _Lib(dev)


def _close(got, want, rel=2 ** -7, what=""):
    scale = float(want.float().abs().max()) + 1e-12
    err = float((got.float() - want.float()).abs().max())
    assert err <= rel * scale, f"{what}: max err {err} vs scale {scale}"


@pytest.mark.parametrize("M,N,K,tile", [(638, 4096, 512, (160, 1)), (638, 1000, 256, (128, 1)), (300, 638, 128, (128, 1)),
                                        (638, 4096, 1024, (256, 2)), (638, 4096, 1024, (256, 1)), (77, 64, 64, (64, 1))])
def test_gemm_ex_tiles_bias_residual_fp32(cuda_device, M, N, K, tile):
    k = _k(cuda_device)
    g = torch.Generator().manual_seed(M + N)
    A = (torch.randn(M, K, generator=g) * 0.5).to(cuda_device).bfloat16()
    B = (torch.randn(N, K, generator=g) * 0.5).to(cuda_device).bfloat16()
    bias = torch.randn(N, generator=g).to(cuda_device).bfloat16()
    want = A.float() @ B.float().T
    Np = (N + 7) // 8 * 8                   # output rows are written with 16-byte stores: padded leading dimension
    out = torch.zeros(M, Np, dtype=torch.bfloat16, device=cuda_device)[:, :N]
    k.gemm(A, B, out=out, M=M, N=N, K=K, tile=tile)
    _close(out, want, what="plain")
    k.gemm(A, B, out=out, M=M, N=N, K=K, tile=tile, bias=bias)
    _close(out, want + bias.float(), what="bias")
    o32 = torch.zeros(M, Np, dtype=torch.float32, device=cuda_device)[:, :N]
    k.gemm(A, B, out32=o32, M=M, N=N, K=K, tile=tile)
    assert float((o32 - want).abs().max()) <= 1e-3 * float(want.abs().max())
    with pytest.raises(RuntimeError):       # an unaligned leading dimension is refused, not mis-stored
        k.gemm(A, B, out=torch.zeros(M, N + 1, dtype=torch.bfloat16, device=cuda_device)[:, :N], M=M, N=N, K=K, tile=tile)
    if N % 8 == 0:
        R = torch.randn(M, N, generator=g).to(cuda_device).bfloat16()
        k.gemm(A, B, out=out, M=M, N=N, K=K, tile=tile, R=R, ldr=N)
        _close(out, want.bfloat16().float() + R.float(), what="residual")


def test_gemm_ex_batched_heads(cuda_device):
    """Per-head products of one sequence: heads side by side along the columns (Q K^T) and stacked
    along the rows (P V with V^T), ragged key count 150 (not a multiple of 64)."""
    k = _k(cuda_device)
    H, L, hd = 4, 150, 128
    Lp = (L + 7) // 8 * 8
    g = torch.Generator().manual_seed(2)
    Q = (torch.randn(L, H * hd, generator=g) * 0.3).to(cuda_device).bfloat16()
    Kk = (torch.randn(L, H * hd, generator=g) * 0.3).to(cuda_device).bfloat16()
    S = torch.zeros(H, L, Lp, dtype=torch.float32, device=cuda_device)
    k.gemm(Q, Kk, out32=S, M=L, N=L, K=hd, ldc=Lp, tile=(128, 1), batch=H, a_col=hd, b_col=hd, c_batch=L * Lp)
    want = torch.einsum("qhd,khd->hqk", Q.float().view(L, H, hd), Kk.float().view(L, H, hd))
    assert float((S[:, :, :L] - want).abs().max()) <= 1e-3 * float(want.abs().max())
    P = torch.zeros(H, L, Lp, dtype=torch.bfloat16, device=cuda_device)
    P[:, :, :L] = torch.softmax(want, -1).bfloat16()
    Vt = torch.zeros(H * hd, Lp, dtype=torch.bfloat16, device=cuda_device)
    V = (torch.randn(L, H * hd, generator=g) * 0.3).to(cuda_device).bfloat16()
    Vt[:, :L] = V.T
    O = torch.zeros(L, H * hd, dtype=torch.bfloat16, device=cuda_device)
    k.gemm(P, Vt, out=O, M=L, N=hd, K=L, lda=Lp, ldb=Lp, ldc=H * hd, tile=(128, 1), batch=H, a_row=L, b_row=hd, c_batch=hd)
    wantO = torch.einsum("hqk,khd->qhd", P[:, :, :L].float(), V.float().view(L, H, hd)).reshape(L, H * hd)
    _close(O, wantO, what="PV")


@pytest.mark.parametrize("causal", [1, 0])
def test_softmax_rows_and_backward(cuda_device, causal):
    from bimodalattack_b200 import native
    lib = native.load()
    H, L = 3, 77
    Lp = 80
    g = torch.Generator().manual_seed(4)
    S = (torch.randn(H, L, Lp, generator=g) * 3).to(cuda_device)
    dP = torch.randn(H, L, Lp, generator=g).to(cuda_device)
    P = torch.full((H, L, Lp), 7.0, dtype=torch.bfloat16, device=cuda_device)
    st = native.stream_ptr(cuda_device)
    scale = 0.25
    native.check(lib.bma_softmax_rows(S.data_ptr(), Lp, P.data_ptr(), Lp, H, L, scale, causal, st), "softmax")
    mask = torch.ones(L, L, dtype=torch.bool, device=cuda_device).tril() if causal else torch.ones(L, L, dtype=torch.bool, device=cuda_device)
    x = (S[:, :, :L] * scale).masked_fill(~mask, float("-inf"))
    want = torch.softmax(x, -1)
    _close(P[:, :, :L], want, what="softmax")
    assert float(P[:, :, L:].float().abs().max()) == 0.0
    dS = torch.full((H, L, Lp), 7.0, dtype=torch.bfloat16, device=cuda_device)
    native.check(lib.bma_softmax_bwd_rows(P.data_ptr(), Lp, dP.data_ptr(), Lp, dS.data_ptr(), Lp, H, L, scale, causal, st), "bwd")
    Pf = P[:, :, :L].float()
    delta = (Pf * dP[:, :, :L]).sum(-1, keepdim=True)
    wantdS = Pf * (dP[:, :, :L] - delta) * scale
    _close(dS[:, :, :L], wantdS, what="dS")
    assert float(dS[:, :, L:].float().abs().max()) == 0.0


def test_transpose_rope_glu_act_ce_norm_backwards(cuda_device):
    from bimodalattack_b200 import native
    lib = native.load()
    st = native.stream_ptr(cuda_device)
    dev = cuda_device
    g = torch.Generator().manual_seed(8)
    bf = torch.bfloat16
    # batched transpose
    src = torch.randn(3, 70, 50, generator=g).to(dev).to(bf)
    dst = torch.zeros(3, 50, 72, dtype=bf, device=dev)
    native.check(lib.bma_transpose_bf16(src.data_ptr(), 50, 70 * 50, dst.data_ptr(), 72, 50 * 72, 70, 50, 3, st), "T")
    assert torch.equal(dst[:, :, :70], src.transpose(1, 2))
    # rmsnorm backward (+ residual gradient) vs autograd of the HF formula
    rows, d = 37, 256
    x = torch.randn(rows, d, generator=g).to(dev).to(bf)
    w = (1 + 0.1 * torch.randn(d, generator=g)).to(dev).to(bf)
    dy = torch.randn(rows, d, generator=g).to(dev).to(bf)
    dres = torch.randn(rows, d, generator=g).to(dev).to(bf)
    xf = x.float().requires_grad_()
    y = w.float() * (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-5))
    want = torch.autograd.grad(y, xf, dy.float())[0] + dres.float()
    dx = torch.zeros(rows, d, dtype=bf, device=dev)
    native.check(lib.bma_rmsnorm_bwd(dy.data_ptr(), d, x.data_ptr(), d, w.data_ptr(), dres.data_ptr(), d, dx.data_ptr(), d,
                                     rows, d, C.c_float(1e-5), 0, st), "rmsnorm_bwd")
    _close(dx, want, rel=2 ** -6, what="rmsnorm_bwd")
    # layernorm forward / backward
    b = (0.1 * torch.randn(d, generator=g)).to(dev).to(bf)
    y2 = torch.zeros(rows, d, dtype=bf, device=dev)
    native.check(lib.bma_layernorm(x.data_ptr(), d, w.data_ptr(), b.data_ptr(), y2.data_ptr(), d, rows, d, C.c_float(1e-5), st), "ln")
    xf = x.float().requires_grad_()
    ln = torch.nn.functional.layer_norm(xf, (d,), w.float(), b.float(), 1e-5)
    _close(y2, ln, what="layernorm")
    want = torch.autograd.grad(ln, xf, dy.float())[0] + dres.float()
    native.check(lib.bma_layernorm_bwd(dy.data_ptr(), d, x.data_ptr(), d, w.data_ptr(), dres.data_ptr(), d, dx.data_ptr(), d,
                                       rows, d, C.c_float(1e-5), st), "ln_bwd")
    _close(dx, want, rel=2 ** -6, what="layernorm_bwd")
    # rotary backward: adjoint of y1 = x1 c - x2 s, y2 = x2 c + x1 s
    n_heads, hd, pos0 = 3, 128, 5
    cos = torch.cos(torch.rand(64, hd // 2, generator=g) * 6).to(dev)
    sin = torch.sin(torch.rand(64, hd // 2, generator=g) * 6).to(dev)
    gr = torch.randn(rows, n_heads * hd, generator=g).to(dev).to(bf)
    g4 = gr.float().view(rows, n_heads, 2, hd // 2)
    c = cos[pos0:pos0 + rows].view(rows, 1, hd // 2)
    s_ = sin[pos0:pos0 + rows].view(rows, 1, hd // 2)
    want = torch.stack([g4[:, :, 0] * c + g4[:, :, 1] * s_, g4[:, :, 1] * c - g4[:, :, 0] * s_], 2).reshape(rows, -1)
    buf = gr.clone()
    native.check(lib.bma_rope_bwd(buf.data_ptr(), n_heads * hd, cos.data_ptr(), sin.data_ptr(), rows, n_heads, hd, pos0, st), "rope")
    _close(buf, want, rel=2 ** -6, what="rope_bwd")
    # GLU (packed gate|up per 256 columns) forward / backward, SiLU and tanh-GELU
    dff = 256
    gu = torch.randn(rows, 2 * dff, generator=g).to(dev).to(bf)
    dact = torch.randn(rows, dff, generator=g).to(dev).to(bf)
    for act, fn in ((0, torch.nn.functional.silu), (1, lambda t: torch.nn.functional.gelu(t, approximate="tanh"))):
        guf = gu.float().requires_grad_()
        blocks = guf.view(rows, dff // 128, 2, 128)
        ref = (fn(blocks[:, :, 0]) * blocks[:, :, 1]).reshape(rows, dff)
        out = torch.zeros(rows, dff, dtype=bf, device=dev)
        native.check(lib.bma_glu(gu.data_ptr(), 2 * dff, out.data_ptr(), dff, rows, dff, act, st), "glu")
        _close(out, ref, rel=2 ** -6, what="glu")
        want = torch.autograd.grad(ref, guf, dact.float())[0]
        dgu = torch.zeros(rows, 2 * dff, dtype=bf, device=dev)
        native.check(lib.bma_glu_bwd(gu.data_ptr(), 2 * dff, dact.data_ptr(), dff, dgu.data_ptr(), 2 * dff, rows, dff, act, st), "glu_bwd")
        _close(dgu, want, rel=2 ** -6, what="glu_bwd")
    # pointwise activations
    xs = torch.randn(rows * d, generator=g).to(dev).to(bf)
    gs = torch.randn(rows * d, generator=g).to(dev).to(bf)
    for kind, fn in ((0, lambda t: t * torch.sigmoid(1.702 * t)), (1, torch.nn.functional.gelu)):
        xf = xs.float().requires_grad_()
        ref = fn(xf)
        out = torch.zeros_like(xs)
        native.check(lib.bma_act(xs.data_ptr(), None, out.data_ptr(), xs.numel(), kind, st), "act")
        _close(out, ref, what="act")
        want = torch.autograd.grad(ref, xf, gs.float())[0]
        native.check(lib.bma_act(xs.data_ptr(), gs.data_ptr(), out.data_ptr(), xs.numel(), kind, st), "act_bwd")
        _close(out, want, rel=2 ** -6, what="act_bwd")
    # cross-entropy gradient over T rows
    T, V = 5, 1000
    logits = (torch.randn(T, V, generator=g) * 3).to(dev).to(bf)
    tgt = torch.randint(0, V, (T,), generator=g).to(dev)
    lf = logits.float().requires_grad_()
    loss = torch.nn.functional.cross_entropy(lf, tgt)
    want = torch.autograd.grad(loss, lf)[0]
    dl = torch.zeros(T, V, dtype=bf, device=dev)
    lrow = torch.zeros(T, dtype=torch.float32, device=dev)
    native.check(lib.bma_ce_grad(logits.data_ptr(), V, tgt.data_ptr(), dl.data_ptr(), V, lrow.data_ptr(), T, V, st), "ce")
    _close(dl, want, what="ce_grad")
    assert abs(float(lrow.mean()) - float(loss)) <= 1e-3 * abs(float(loss))


def _grad_setup(device, size, px, mode):
    import torchvision.transforms as T
    from bimodalattack_b200 import BimodalAttack, BimodalAttackConfig, synthetic as syn
    tok = syn.make_tokenizer(1024 if size == "tiny128" else 4096)
    proc = syn.make_llava_processor(tok, px)
    model = syn.make_llava(size, dtype=torch.bfloat16, device=device, head_gain=20.0, attn_implementation="sdpa")
    norm = T.Normalize(syn.CLIP_MEAN, syn.CLIP_STD)
    image = torch.rand((1, 3, px, px), generator=torch.Generator().manual_seed(5)).to(device)
    flags = dict(pgd_attack=True, gcg_attack=True, joint_eval=True) if mode == "joint" else dict(pgd_attack=False, gcg_attack=True)
    cfg = BimodalAttackConfig(num_steps=1, search_width=16, topk=32, seed=1, allow_non_ascii=True, alpha=4 / 255,
                              eps=64 / 255, optim_str_init="x x x x x x x x x x", verbosity="ERROR", images_folder=None, **flags)
    atk = BimodalAttack(model, tok, proc, cfg, normalize=norm)
    atk._prepare([{"role": "user", "content": "w10 w11 w12 w13 w14 w15"}], "w20 w21 w22 w23 w24")
    return atk, model, tok, proc, norm, image, cfg


@pytest.mark.parametrize("rc_px", [0, 42])
def test_native_vision_tower_forward_and_backward(cuda_device, rc_px):
    """CLIP tower + projector on the native kernels (head_dim 64, quick_gelu, LayerNorm, biases, patch
    conv as GEMM, Normalize folded in) against HF: features vs ``get_image_features`` and the pixel
    gradient vs fp32 autograd, with HF's own bf16 autograd as the error yardstick. rc_px = 42: a second
    tower swapped in at its own image size (LLaVA-RC, BASELINE config 3)."""
    import copy
    from bimodalattack_b200 import synthetic as syn
    from bimodalattack_b200.attack import image_features
    from bimodalattack_b200.gradient import NativeVisionTower
    atk, model, tok, proc, norm, image, cfg = _grad_setup(cuda_device, "small128v", 56, "joint")
    if rc_px:
        syn.swap_in_robust_clip(model, image_size=rc_px)
        image = torch.rand((1, 3, rc_px, rc_px), generator=torch.Generator().manual_seed(6)).to(cuda_device)
    assert NativeVisionTower.supported(model, norm) is None
    vt = NativeVisionTower(model, norm)
    got = vt.forward(image).clone()
    want_bf16 = image_features(model, norm(image), False)
    twin = copy.deepcopy(model).float()
    img32 = image.clone().requires_grad_()
    want = image_features(twin, norm(img32), False)
    assert got.shape == want.shape

    def rel(a, b):
        return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-30))

    e_n, e_h = rel(got, want), rel(want_bf16, want)
    print(f"features rel err vs fp32: native {e_n:.4f}  hf-bf16 {e_h:.4f}")
    assert e_n <= 1.5 * e_h + 2e-3
    # scoring-phase forward (no backward follows): one flash-attention launch per layer instead of the products
    got_flash = vt.forward(image, keep=False).clone()
    e_f = rel(got_flash, want)
    print(f"features rel err vs fp32: native flash {e_f:.4f}")
    assert e_f <= 1.5 * e_h + 2e-3
    assert bool(torch.equal(vt.forward(image).clone(), got))      # ... and leaves the keep=True state usable
    d_feats = (torch.randn(want.shape, generator=torch.Generator().manual_seed(1)).to(cuda_device) * 1e-2).bfloat16()
    g_true = torch.autograd.grad(want, img32, d_feats.float())[0]
    img16 = image.clone().requires_grad_()
    g_hf = torch.autograd.grad(image_features(model, norm(img16), False), img16, d_feats)[0]
    g_nat = vt.backward(d_feats)
    e_n, e_h = rel(g_nat, g_true), rel(g_hf, g_true)
    print(f"pixel grad rel err vs fp32: native {e_n:.4f}  hf-bf16 {e_h:.4f}")
    assert e_n <= 1.5 * e_h + 5e-3
    assert float((torch.sign(g_nat) == torch.sign(g_true)).float().mean()) >= float((torch.sign(g_hf) == torch.sign(g_true)).float().mean()) - 0.02


@pytest.mark.parametrize("size,px,mode", [("tiny128", 28, "joint"), ("tiny128", 28, "gcg_only"), ("small128", 56, "joint"),
                                          ("small128v", 56, "joint")])
def test_native_gradient_matches_autograd(cuda_device, size, px, mode):
    """compute_gradient on the native kernels vs autograd: the fp32 gradient of the same model is the
    truth; the reference's bf16 autograd path (HF, sdpa) sets the error scale."""
    import copy
    from bimodalattack_b200 import BimodalAttack
    atk, model, tok, proc, norm, image, cfg = _grad_setup(cuda_device, size, px, mode)
    assert atk._native_gradient() is not None, atk._native_grad_why
    assert (atk._native_tower() is not None) == (size == "small128v"), atk._native_vt_why
    optim_ids = tok(cfg.optim_str_init, add_special_tokens=False, return_tensors="pt")["input_ids"].to(cuda_device)
    joint = cfg.pgd_attack
    atk.graph_gradient = False
    img = image.clone().requires_grad_() if joint else None
    n_tok, n_img = atk._gradient_eager(optim_ids, img)
    img = image.clone().requires_grad_() if joint else None
    h_tok, h_img = atk._gradient_hf(optim_ids, img)
    twin = BimodalAttack(copy.deepcopy(model).float(), tok, proc, cfg, normalize=norm, engine=atk.engine)
    twin.segment_ids, twin.target_ids = atk.segment_ids, atk.target_ids
    twin.segment_embeds = {k: twin.embedding_layer(v) for k, v in atk.segment_ids.items()}
    img = image.clone().requires_grad_() if joint else None
    W32 = twin.embedding_layer.weight
    # fp32 truth, written out (the twin's token_grad GEMM is bf16-only): d loss / d one-hot = g_embeds . W^T
    e = twin.segment_embeds
    oe = W32.detach()[optim_ids].clone().requires_grad_()
    if joint:
        from bimodalattack_b200.attack import image_features
        feats = image_features(twin.model, twin._normalize_pixels(img), False)
        seq = [e["before_img"], feats, e["before_suffix"], oe, e["after"], e["target"]]
    else:
        seq = [e["before"], oe, e["after"], e["target"]]
    x = torch.cat(seq, 1)
    T = atk.target_ids.shape[1]
    logits = twin.model(inputs_embeds=x).logits[:, -T - 1:-1]
    loss = torch.nn.functional.cross_entropy(logits.reshape(-1, logits.shape[-1]), atk.target_ids.view(-1))
    if joint:
        g_e, t_img = torch.autograd.grad(loss, (oe, img))
    else:
        g_e, t_img = torch.autograd.grad(loss, oe)[0], None
    t_tok = (g_e[0] @ W32.detach().T).unsqueeze(0)

    def rel(a, b):
        return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-30))

    e_native, e_hf = rel(n_tok, t_tok), rel(h_tok, t_tok)
    print(f"token grad rel err vs fp32: native {e_native:.4f}  hf-bf16 {e_hf:.4f}")
    assert e_native <= 1.5 * e_hf + 5e-3
    assert abs(float(atk._native_grad.loss) - float(loss)) <= 2e-2 * abs(float(loss)) + 1e-3
    # the top-k sets the sampler draws from overlap like the reference's own do
    k = 16
    top = lambda t: set((-t.float()[0, 3]).topk(k).indices.tolist())   # noqa: E731
    assert len(top(n_tok) & top(t_tok)) >= len(top(h_tok) & top(t_tok)) - 2
    if joint:
        ei_native, ei_hf = rel(n_img, t_img), rel(h_img, t_img)
        print(f"image grad rel err vs fp32: native {ei_native:.4f}  hf-bf16 {ei_hf:.4f}")
        assert ei_native <= 1.5 * ei_hf + 5e-3
        agree = float((torch.sign(n_img) == torch.sign(t_img)).float().mean())
        agree_hf = float((torch.sign(h_img) == torch.sign(t_img)).float().mean())
        assert agree >= agree_hf - 0.02


def test_gemm_ex_causal_modes(cuda_device):
    """Products of one causal sequence: tiles above the diagonal skipped (mode 1), K ranges cut at the
    diagonal (modes 2 and 3). Same numbers as the dense products on causally masked operands."""
    k = _k(cuda_device)
    H, L, hd = 3, 300, 128
    Lp = (L + 7) // 8 * 8
    g = torch.Generator().manual_seed(5)
    Q = (torch.randn(L, H * hd, generator=g) * 0.3).to(cuda_device).bfloat16()
    Kk = (torch.randn(L, H * hd, generator=g) * 0.3).to(cuda_device).bfloat16()
    tril = torch.ones(L, L, dtype=torch.bool, device=cuda_device).tril()
    S = torch.full((H, L, Lp), float("nan"), dtype=torch.float32, device=cuda_device)
    k.gemm(Q, Kk, out32=S, M=L, N=L, K=hd, ldc=Lp, tile=(128, 1), batch=H, a_col=hd, b_col=hd, c_batch=L * Lp, causal=1)
    want = torch.einsum("qhd,khd->hqk", Q.float().view(L, H, hd), Kk.float().view(L, H, hd))
    got = S[:, :, :L]
    assert float((got - want)[:, tril].abs().max()) <= 1e-3 * float(want.abs().max())     # the lower triangle is all there
    assert bool(torch.isnan(got[:, :128, 128:]).all())                                    # tiles above the diagonal: untouched
    # mode 2: O = P V with a lower-triangular P
    P = torch.zeros(H, L, Lp, dtype=torch.bfloat16, device=cuda_device)
    P[:, :, :L] = (torch.rand(H, L, L, generator=g).to(cuda_device) * tril).bfloat16()
    V = (torch.randn(L, H * hd, generator=g) * 0.3).to(cuda_device).bfloat16()
    Vt = torch.zeros(H * hd, Lp, dtype=torch.bfloat16, device=cuda_device)
    Vt[:, :L] = V.T
    O = torch.zeros(L, H * hd, dtype=torch.bfloat16, device=cuda_device)
    k.gemm(P, Vt, out=O, M=L, N=hd, K=L, lda=Lp, ldb=Lp, ldc=H * hd, tile=(128, 1), batch=H, a_row=L, b_row=hd, c_batch=hd, causal=2)
    _close(O, torch.einsum("hqk,khd->qhd", P[:, :, :L].float(), V.float().view(L, H, hd)).reshape(L, H * hd), what="causal PV")
    # mode 3: dK = dS^T Q with dS^T upper-triangular in (key, query)
    dSt = torch.zeros(H, L, Lp, dtype=torch.bfloat16, device=cuda_device)
    dSt[:, :, :L] = (torch.randn(H, L, L, generator=g).to(cuda_device) * tril.T).bfloat16()
    Qt = torch.zeros(H * hd, Lp, dtype=torch.bfloat16, device=cuda_device)
    Qt[:, :L] = Q.T
    dK = torch.zeros(L, H * hd, dtype=torch.bfloat16, device=cuda_device)
    k.gemm(dSt, Qt, out=dK, M=L, N=hd, K=L, lda=Lp, ldb=Lp, ldc=H * hd, tile=(128, 1), batch=H, a_row=L, b_row=hd, c_batch=hd, causal=3)
    _close(dK, torch.einsum("hkq,qhd->khd", dSt[:, :, :L].float(), Q.float().view(L, H, hd)).reshape(L, H * hd), what="causal dK")
