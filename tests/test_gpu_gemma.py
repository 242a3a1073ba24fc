"""Gemma-3 decoder on the B200 scorer (BASELINE.json config 4):
head_dim 256 attention, GQA, q/k-norm, (1+w) RMSNorm with four norms per layer, GeGLU, scaled
embeddings, tied LM head, local + global rotary tables -- against the oracle (the reference's
dataflow) on a random-init Gemma-3 with the 4b attention geometry in miniature.

Tolerances are the ones of tests/test_gpu_attack.py: 2 bf16 ulps against the reference's own bf16
loss and 2e-2 relative against an fp32 recomputation on the same (bf16-valued) weights.
"""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu

PROMPT = "w10 w11 w12 w13 w14 w15"
TARGET = "w20 w21 w22 w23 w24"
INIT = "x x x x x x x x x x"

MODES = {
    "joint": dict(pgd_attack=True, gcg_attack=True, joint_eval=True),
    "gcg_only": dict(pgd_attack=False, gcg_attack=True),
    "pgd_gcg_nojoint": dict(pgd_attack=True, gcg_attack=True, joint_eval=False),
}


def _setup(device, **cfg_kw):
    import torchvision.transforms as T
    from bimodalattack_b200 import BimodalAttackConfig, synthetic as syn
    tok = syn.make_tokenizer(1024, image_token="<image_soft_token>",
                             extra_tokens=["<start_of_image>", "<end_of_image>", "<start_of_turn>", "<end_of_turn>"])
    proc = syn.Gemma3Processor(tok)
    model = syn.make_gemma3("tiny256", dtype=torch.bfloat16, device=device, embed_gain=4.0, randomize_norms=True)
    model.config._attn_implementation = "sdpa"
    norm = T.Normalize(syn.GEMMA_MEAN, syn.GEMMA_STD)
    image = torch.rand((1, 3, 56, 56), generator=torch.Generator().manual_seed(5)).to(device)
    kw = dict(num_steps=3, search_width=48, topk=32, seed=1, allow_non_ascii=True, alpha=4 / 255, eps=64 / 255,
              optim_str_init=INIT, verbosity="ERROR", images_folder=None, filter_ids=True)
    kw.update(cfg_kw)
    return model, tok, proc, norm, image, BimodalAttackConfig(**kw)


@pytest.mark.parametrize("mode", list(MODES))
def test_gemma_candidate_losses_match_oracle(cuda_device, mode):
    from bimodalattack_b200 import BimodalAttack
    from bimodalattack_b200.attack import image_features
    from oracle import hotpath
    model, tok, proc, norm, image, cfg = _setup(cuda_device, **MODES[mode])
    atk = BimodalAttack(model, tok, proc, cfg, normalize=norm)
    assert atk.engine.gemma and atk.engine.head_dim == 256
    atk._prepare([{"role": "user", "content": PROMPT}], TARGET)
    ora = hotpath.OracleAttack(model, tok, proc, cfg, normalize=norm)
    ora32 = hotpath.OracleAttack(copy.deepcopy(model).float(), tok, proc, cfg, normalize=norm)
    n_tok = len(tok(INIT, add_special_tokens=False)["input_ids"])
    sampled = torch.randint(12, 1000, (48, n_tok), generator=torch.Generator().manual_seed(17)).to(cuda_device)
    with torch.no_grad():
        feats = image_features(model, norm(image), True) if cfg.pgd_attack else None
        if mode == "joint":         # Gemma order: before_img, optim, before_suffix, image, after, target
            orders = [("pgd", dict(single=True), True)]
        elif mode == "gcg_only":
            orders = [("gcg", dict(no_joint_eval=True), False)]
        else:
            orders = [("gcg", dict(single=True), False), ("gcg_pgd", dict(single=True), True)]
        for omode, flags, with_img in orders:
            batch = atk._build_input_embeds(sampled, image=feats if with_img else None, search_width=48,
                                            mode=omode, **flags)
            got = atk._compute_candidates_loss_original(48, batch)
            key_loss, key_idx = atk._winner()
            for o in (ora, ora32):
                o.target_ids = atk.target_ids
                o.seg_embeds = {k: o.emb(v) for k, v in atk.segment_ids.items()}
            order = hotpath.segment_order(omode, "gemma3", **flags)
            assert order == batch.order
            if with_img:
                assert order.index("optim") < order.index("image")       # the image rides in the tail
            want_bf16, _ = hotpath.candidates_loss(
                model, hotpath.assemble(ora._parts(sampled, feats if with_img else None), order, 48), atk.target_ids, 16)
            f32 = feats.float() if feats is not None else None
            want_fp32, _ = hotpath.candidates_loss(
                ora32.model, hotpath.assemble(ora32._parts(sampled, f32 if with_img else None), order, 48),
                atk.target_ids, 16)
            want_bf16, want_fp32 = want_bf16.float(), want_fp32.float()
            ulp = 2.0 ** (torch.floor(torch.log2(want_bf16.abs().clamp_min(1e-3))) - 7)
            assert bool(((got - want_bf16).abs() <= 2 * ulp + 1e-3).all()), (got - want_bf16).abs().max()
            assert bool(((got - want_fp32).abs() <= 2e-2 * want_fp32.abs() + 1e-3).all()), (got - want_fp32).abs().max()
            assert key_idx == int(got.argmin()) and abs(key_loss - float(got.min())) == 0.0
            assert float(want_fp32[key_idx]) <= float(want_fp32.min()) + 2e-2 * abs(float(want_fp32.min())) + 1e-3
            assert float(want_fp32.max() - want_fp32.min()) > 0.1


def test_gemma_gradient_uses_unscaled_table(cuda_device):
    """The reference multiplies the one-hot by the RAW embedding table (:968), i.e. without
    Gemma's sqrt(d) scale; the token gradient must do the same."""
    from bimodalattack_b200 import BimodalAttack
    from oracle import hotpath
    model, tok, proc, norm, image, cfg = _setup(cuda_device, **MODES["gcg_only"])
    atk = BimodalAttack(model, tok, proc, cfg, normalize=norm)
    atk._prepare([{"role": "user", "content": PROMPT}], TARGET)
    ora = hotpath.OracleAttack(model, tok, proc, cfg, normalize=norm)
    ora.target_ids = atk.target_ids
    ora.seg_embeds = {k: ora.emb(v) for k, v in atk.segment_ids.items()}
    optim_ids = tok(INIT, add_special_tokens=False, return_tensors="pt")["input_ids"].to(cuda_device)
    g_tok, _ = atk.compute_gradient(optim_ids, None)
    w_tok, _ = ora.gradient(optim_ids, None)
    scale = w_tok.float().abs().max()
    assert float(scale) > 0
    assert float((g_tok.float() - w_tok.float()).abs().max()) <= 2 ** -7 * float(scale)


@pytest.mark.parametrize("mode", ["joint", "gcg_only", "pgd_only"])
def test_gemma_run_end_to_end(cuda_device, mode):
    import bimodalattack_b200 as bma
    flags = dict(MODES.get(mode, dict(pgd_attack=True, gcg_attack=False)))
    model, tok, proc, norm, image, cfg = _setup(cuda_device, **flags)
    cfg.num_steps = 4
    res = bma.run(model, tok, proc, [{"role": "user", "content": PROMPT}], "goal", TARGET,
                  image.clone() if cfg.pgd_attack else None, cfg, normalize=norm)
    assert len(res.losses) == 4 and all(l == l and l > 0 for l in res.losses)
    assert res.best_loss == min(res.losses)
    assert res.losses[-1] < res.losses[0] + 0.05


def test_gemma_first_step_matches_oracle(cuda_device):
    import bimodalattack_b200 as bma
    from oracle import hotpath
    model, tok, proc, norm, image, cfg = _setup(cuda_device, **MODES["gcg_only"])
    cfg.num_steps = 1
    atk = bma.BimodalAttack(model, tok, proc, cfg, normalize=norm)
    res = atk.run([{"role": "user", "content": PROMPT}], "goal", TARGET, None)
    log = hotpath.OracleAttack(model, tok, proc, cfg, normalize=norm).run(
        [{"role": "user", "content": PROMPT}], TARGET, None)
    assert abs(res.losses[0] - log.losses[0]) <= 0.0625 * 2 + 2e-2 * abs(log.losses[0])


def test_add_norm_kernel(lib, cuda_device):
    """x <- bf16(x + norm(y)); xn <- norm(x) in the Gemma (1+w) form, xn aliasing y."""
    from bimodalattack_b200 import native
    rows, d, eps = 301, 2560, 1e-6
    g = torch.Generator().manual_seed(3)
    y = torch.randn((rows, d), generator=g).to(cuda_device).bfloat16()
    x = torch.randn((rows, d), generator=g).to(cuda_device).bfloat16()
    w1 = (torch.rand(d, generator=g) - 0.5).to(cuda_device).bfloat16()
    w2 = (torch.rand(d, generator=g) - 0.5).to(cuda_device).bfloat16()

    def norm(t, w):
        f = t.float()
        return (f * torch.rsqrt(f.pow(2).mean(-1, keepdim=True) + eps) * (1.0 + w.float())).to(torch.bfloat16)

    want_x = x + norm(y, w1)
    want_xn = norm(want_x, w2)
    buf = y.clone()
    xs = x.clone()
    native.check(lib.bma_add_norm(buf.data_ptr(), d, w1.data_ptr(), xs.data_ptr(), d, w2.data_ptr(),
                                  buf.data_ptr(), d, rows, d, eps, 1, torch.cuda.current_stream().cuda_stream),
                 "add_norm")
    torch.cuda.synchronize()
    # identical rounding points; only the fp32 summation order of the mean differs (<= 1 bf16 ulp)
    assert float((xs.float() - want_x.float()).abs().max()) <= 2 ** -7 * float(want_x.float().abs().max())
    assert float((xs != want_x).float().mean()) < 1e-3
    assert float((buf.float() - want_xn.float()).abs().max()) <= 2 ** -6 * float(want_xn.float().abs().max())


def test_gemma_suffix_head_sharing(cuda_device):
    """Gemma joint order puts the image AFTER the suffix, so the tail is long (optim | before_suffix
    | image | after | target) and only its first rows can be shared with the base suffix."""
    from bimodalattack_b200 import BimodalAttack
    from bimodalattack_b200.attack import image_features
    model, tok, proc, norm, image, cfg = _setup(cuda_device, **MODES["joint"])
    atk = BimodalAttack(model, tok, proc, cfg, normalize=norm)
    atk._prepare([{"role": "user", "content": PROMPT}], TARGET)
    base = tok(INIT, add_special_tokens=False, return_tensors="pt")["input_ids"].to(cuda_device).reshape(-1)
    g = torch.Generator().manual_seed(4)
    cand = base.cpu().repeat(40, 1)
    for i in range(1, 40):
        cand[i, int(torch.randint(0, base.numel(), (1,), generator=g))] = int(torch.randint(12, 1000, (1,), generator=g))
    cand = cand.to(cuda_device)
    with torch.no_grad():
        feats = image_features(model, norm(image), True)
        batch = atk._build_input_embeds_pgd(cand, feats, search_width=40, single=True)
        full = atk._compute_candidates_loss_original(40, batch).clone()
        rows_full = atk.engine.last_rows
        atk._base = (cand, base.to(torch.int64).contiguous())
        shared = atk._compute_candidates_loss_original(40, batch).clone()
        rows_shared = atk.engine.last_rows
    assert rows_shared < rows_full + atk._plan(batch).n_tail
    assert float((shared - full).abs().max()) <= 1e-2 * float(full.abs().max())
    assert float(shared[0]) == float(shared[0]) and float((shared[1:] - shared[0]).abs().max()) > 0
