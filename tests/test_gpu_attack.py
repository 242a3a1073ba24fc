"""End-to-end parity of the B200 path against the oracle (the reference's dataflow restated in
torch) on the GPU: gradient, candidate scoring in every segment order, winner selection, and the
public ``run()``. Model: random-init LLaVA-architecture decoder with head_dim 128 in bf16.

Tolerances (stated per assert): the reference computes the per-candidate loss in bf16 (ulp 2^-5 =
0.031 at loss ~4-8, 2^-4 above 8); the B200 path keeps the same bf16 rounding points inside the
decoder but accumulates the CE in fp32. Both are compared with an fp32 run of the oracle on the
same (bf16-valued) weights.
"""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu

PROMPT = "w10 w11 w12 w13 w14 w15"
TARGET = "w20 w21 w22 w23 w24"
INIT = "x x x x x x x x x x"


def _setup(device, size="tiny128", image_px=28, **cfg_kw):
    import torchvision.transforms as T
    from bimodalattack_b200 import BimodalAttackConfig, synthetic as syn
    tok = syn.make_tokenizer(1024 if size == "tiny128" else 4096)
    proc = syn.make_llava_processor(tok, image_px)
    model = syn.make_llava(size, dtype=torch.bfloat16, device=device, head_gain=20.0, attn_implementation="sdpa")
    norm = T.Normalize(syn.CLIP_MEAN, syn.CLIP_STD)
    g = torch.Generator().manual_seed(5)
    image = torch.rand((1, 3, image_px, image_px), generator=g).to(device)
    kw = dict(num_steps=3, search_width=48, topk=32, seed=1, allow_non_ascii=True, alpha=4 / 255, eps=64 / 255,
              optim_str_init=INIT, verbosity="ERROR", images_folder=None, filter_ids=True)
    kw.update(cfg_kw)
    return model, tok, proc, norm, image, BimodalAttackConfig(**kw)


def _fp32_twin(model):
    twin = copy.deepcopy(model).float()
    return twin


MODES = {
    "joint": dict(pgd_attack=True, gcg_attack=True, joint_eval=True),
    "gcg_only": dict(pgd_attack=False, gcg_attack=True),
    "pgd_gcg_nojoint": dict(pgd_attack=True, gcg_attack=True, joint_eval=False),
}


@pytest.mark.parametrize("mode", list(MODES))
def test_candidate_losses_match_oracle(cuda_device, mode):
    from bimodalattack_b200 import BimodalAttack
    from bimodalattack_b200.attack import image_features
    from oracle import hotpath
    model, tok, proc, norm, image, cfg = _setup(cuda_device, **MODES[mode])
    atk = BimodalAttack(model, tok, proc, cfg, normalize=norm)
    atk._prepare([{"role": "user", "content": PROMPT}], TARGET)
    ora = hotpath.OracleAttack(model, tok, proc, cfg, normalize=norm)
    ora32 = hotpath.OracleAttack(_fp32_twin(model), tok, proc, cfg, normalize=norm)
    g = torch.Generator().manual_seed(17)
    n_tok = len(tok(INIT, add_special_tokens=False)["input_ids"])
    sampled = torch.randint(8, 1000, (48, n_tok), generator=g).to(cuda_device)
    with torch.no_grad():
        feats = image_features(model, norm(image), False) if cfg.pgd_attack else None
        orders = []
        if mode == "joint":
            orders.append(("pgd", dict(single=True), True))
        elif mode == "gcg_only":
            orders.append(("gcg", dict(no_joint_eval=True), False))
        else:
            orders.append(("gcg", dict(single=True), False))
            orders.append(("gcg_pgd", dict(single=True), True))
        for omode, flags, with_img in orders:
            batch = atk._build_input_embeds(sampled, image=feats if with_img else None, search_width=48,
                                            mode=omode, **flags)
            got = atk._compute_candidates_loss_original(48, batch)
            key_loss, key_idx = atk._winner()
            for o, dtype_tol in ((ora, None), (ora32, None)):
                o.seg_ids = atk.segment_ids
                o.target_ids = atk.target_ids
                o.seg_embeds = {k: o.emb(v) for k, v in atk.segment_ids.items()}
            order = hotpath.segment_order(omode, "llava", **flags)
            assert order == batch.order
            f16 = feats
            want_bf16, _ = hotpath.candidates_loss(model, hotpath.assemble(ora._parts(sampled, f16 if with_img else None), order, 48),
                                                   atk.target_ids, 16)
            f32 = feats.float() if feats is not None else None
            want_fp32, _ = hotpath.candidates_loss(ora32.model, hotpath.assemble(ora32._parts(sampled, f32 if with_img else None), order, 48),
                                                   atk.target_ids, 16)
            want_bf16, want_fp32 = want_bf16.float(), want_fp32.float()
            # (1) against the reference's own bf16 value: 2 bf16 ulps at the loss magnitude
            ulp = 2.0 ** (torch.floor(torch.log2(want_bf16.abs().clamp_min(1e-3))) - 7)
            assert bool(((got - want_bf16).abs() <= 2 * ulp + 1e-3).all()), (got - want_bf16).abs().max()
            # (2) against the fp32 recomputation: bf16 activation noise through the decoder
            assert bool(((got - want_fp32).abs() <= 2e-2 * want_fp32.abs() + 1e-3).all()), (got - want_fp32).abs().max()
            # (3) winner: packed key == fp32 argmin of the returned losses, and tie-aware vs the oracle
            assert key_idx == int(got.argmin()) and abs(key_loss - float(got.min())) == 0.0
            assert float(want_fp32[key_idx]) <= float(want_fp32.min()) + 2e-2 * abs(float(want_fp32.min())) + 1e-3
            # the scores must really differ between candidates, or argmin parity would be vacuous
            assert float(want_fp32.max() - want_fp32.min()) > 0.1


@pytest.mark.parametrize("mode", ["joint", "gcg_only"])
def test_gradient_matches_oracle(cuda_device, mode):
    from bimodalattack_b200 import BimodalAttack
    from oracle import hotpath
    model, tok, proc, norm, image, cfg = _setup(cuda_device, **MODES[mode])
    atk = BimodalAttack(model, tok, proc, cfg, normalize=norm)
    atk._prepare([{"role": "user", "content": PROMPT}], TARGET)
    ora = hotpath.OracleAttack(model, tok, proc, cfg, normalize=norm)
    ora.seg_ids, ora.target_ids = atk.segment_ids, atk.target_ids
    ora.seg_embeds = {k: ora.emb(v) for k, v in atk.segment_ids.items()}
    optim_ids = tok(INIT, add_special_tokens=False, return_tensors="pt")["input_ids"].to(cuda_device)
    img_a = image.clone().requires_grad_() if cfg.pgd_attack else None
    img_b = image.clone().requires_grad_() if cfg.pgd_attack else None
    # this test pins the HF-autograd path (same backward graph as the reference) to the oracle; the
    # native forward/backward has its own parity test against fp32 autograd (test_gpu_gradient.py)
    atk._native_grad, atk._native_grad_why = None, "HF path under test"
    g_tok, g_img = atk.compute_gradient(optim_ids, img_a)
    w_tok, w_img = ora.gradient(optim_ids, img_b)
    assert g_tok.shape == w_tok.shape and g_tok.dtype == w_tok.dtype == torch.bfloat16
    # same bf16 backward up to the embedding gradient; the final contraction differs only in
    # accumulation order -> 1 bf16 ulp relative to the row scale
    scale = w_tok.float().abs().max()
    assert float((g_tok.float() - w_tok.float()).abs().max()) <= 2 ** -7 * float(scale)
    if cfg.pgd_attack:
        assert torch.allclose(g_img, w_img, rtol=0, atol=1e-6 * float(w_img.abs().max()) + 1e-12) or \
            float((g_img - w_img).abs().max()) <= 2e-2 * float(w_img.abs().max())
        assert torch.equal(torch.sign(g_img), torch.sign(w_img)) or \
            float((torch.sign(g_img) != torch.sign(w_img)).float().mean()) < 0.02


@pytest.mark.parametrize("mode", ["joint", "gcg_only", "pgd_gcg_nojoint", "pgd_only"])
def test_run_end_to_end(cuda_device, mode, tmp_path):
    import bimodalattack_b200 as bma
    flags = dict(MODES.get(mode, dict(pgd_attack=True, gcg_attack=False)))
    model, tok, proc, norm, image, cfg = _setup(cuda_device, **flags)
    cfg.images_folder = str(tmp_path / "img") if cfg.pgd_attack else None
    cfg.num_steps = 4
    res = bma.run(model, tok, proc, [{"role": "user", "content": PROMPT}], "goal", TARGET,
                  image.clone() if cfg.pgd_attack else None, cfg, normalize=norm)
    assert len(res.losses) == 4 and len(res.strings) == 4 and len(res.adversarial_suffixes) == 4
    assert all(l == l and l > 0 for l in res.losses)
    assert res.best_loss == min(res.losses)
    assert res.losses[-1] < res.losses[0] + 0.05       # the attack makes progress (or at least holds)
    if cfg.pgd_attack:
        assert (tmp_path / "img" / "3.png").exists()


def test_first_step_matches_oracle_trajectory(cuda_device):
    """Same seed, same inputs: step 0 of run() picks a candidate whose oracle loss is (tie-aware)
    the oracle's minimum, and the PGD image equals the oracle's image bit for bit when the image
    gradients have the same sign pattern."""
    import bimodalattack_b200 as bma
    from oracle import hotpath
    model, tok, proc, norm, image, cfg = _setup(cuda_device, **MODES["joint"])
    cfg.num_steps = 1
    atk = bma.BimodalAttack(model, tok, proc, cfg, normalize=norm)
    res = atk.run([{"role": "user", "content": PROMPT}], "goal", TARGET, image.clone())
    ora = hotpath.OracleAttack(model, tok, proc, cfg, normalize=norm)
    log = ora.run([{"role": "user", "content": PROMPT}], TARGET, image.clone())
    same_pixels = float((atk.final_image.detach() == ora.final_image.detach()).float().mean())
    assert same_pixels > 0.98
    cand = log.candidate_losses[0]
    sampled = log.sampled[0]
    mine = atk.final_optim_ids.squeeze(0).cpu()
    hits = (sampled == mine).all(dim=1).nonzero()
    if hits.numel() > 0:        # identical gradients -> identical candidate set
        assert float(cand[int(hits[0])]) <= float(cand.min()) + 0.0625 * 2
    assert abs(res.losses[0] - log.losses[0]) <= 0.0625 * 2 + 2e-2 * abs(log.losses[0])


def test_llava_rc_pgd_only(cuda_device, tmp_path):
    """BASELINE config 3: LLaVA with a swapped-in robust CLIP tower at its own image size,
    PGD-only (eps 64/255, alpha 4/255): vision-tower gradient -> fused PGD kernel -> batch-1
    scoring. Every step must stay inside the eps-ball and [0,1], and match the oracle's first step."""
    import bimodalattack_b200 as bma
    from bimodalattack_b200 import synthetic as syn
    from oracle import hotpath
    model, tok, proc, norm, _, cfg = _setup(cuda_device, pgd_attack=True, gcg_attack=False, search_width=0)
    syn.swap_in_robust_clip(model, image_size=42)            # 3x3 patches = 9 image tokens
    image = torch.rand((1, 3, 42, 42), generator=torch.Generator().manual_seed(6)).to(cuda_device)
    cfg.num_steps = 3
    atk = bma.BimodalAttack(model, tok, proc, cfg, normalize=norm)
    res = atk.run([{"role": "user", "content": PROMPT}], "goal", TARGET, image.clone())
    final = atk.final_image.detach()
    assert float((final - image).abs().max()) <= cfg.eps + 1e-6
    assert float(final.min()) >= 0.0 and float(final.max()) <= 1.0
    assert len(res.losses) == 3 and res.losses[-1] <= res.losses[0] + 0.05
    cfg1 = bma.BimodalAttackConfig(**{**cfg.__dict__, "num_steps": 1})
    log = hotpath.OracleAttack(model, tok, proc, cfg1, normalize=norm).run(
        [{"role": "user", "content": PROMPT}], TARGET, image.clone())
    assert abs(res.losses[0] - log.losses[0]) <= 0.0625 * 2 + 2e-2 * abs(log.losses[0])


def test_small_model_joint_scoring_scale(cuda_device):
    """A deeper/wider decoder (3 layers, d=512, 4 heads, 16 image tokens, sw=200 -> several
    128-row tiles, ragged last tile): losses vs the fp32 oracle and the winner."""
    from bimodalattack_b200 import BimodalAttack
    from bimodalattack_b200.attack import image_features
    from oracle import hotpath
    model, tok, proc, norm, image, cfg = _setup(cuda_device, size="small128", image_px=56, search_width=200,
                                                **MODES["joint"])
    atk = BimodalAttack(model, tok, proc, cfg, normalize=norm)
    atk._prepare([{"role": "user", "content": PROMPT}], TARGET)
    ora32 = hotpath.OracleAttack(_fp32_twin(model), tok, proc, cfg, normalize=norm)
    ora32.seg_embeds = {k: ora32.emb(v) for k, v in atk.segment_ids.items()}
    n_tok = len(tok(INIT, add_special_tokens=False)["input_ids"])
    sampled = torch.randint(8, 4000, (200, n_tok), generator=torch.Generator().manual_seed(3)).to(cuda_device)
    with torch.no_grad():
        feats = image_features(model, norm(image), False)
        batch = atk._build_input_embeds_pgd(sampled, feats, search_width=200, single=True)
        got = atk._compute_candidates_loss_original(200, batch)
        _, idx = atk._winner()
        order = hotpath.segment_order("pgd", "llava", single=True)
        want, _ = hotpath.candidates_loss(ora32.model, hotpath.assemble(ora32._parts(sampled, feats.float()), order, 200),
                                          atk.target_ids, 50)
    assert bool(((got - want).abs() <= 2e-2 * want.abs() + 1e-3).all()), (got - want).abs().max()
    assert idx == int(got.argmin())
    assert float(want[idx]) <= float(want.min()) + 2e-2 * abs(float(want.min())) + 1e-3


def _related_candidates(base, n, vocab_lo, vocab_hi, n_replace, seed, device):
    """n candidates that differ from ``base`` ([L]) at n_replace random positions, plus the edge
    cases: identical to the base, first position replaced, last position replaced."""
    g = torch.Generator().manual_seed(seed)
    L = base.numel()
    cand = base.cpu().repeat(n, 1)
    for i in range(n):
        pos = torch.randperm(L, generator=g)[:n_replace]
        cand[i, pos] = torch.randint(vocab_lo, vocab_hi, (n_replace,), generator=g)
    cand[0] = base.cpu()
    cand[1] = base.cpu(); cand[1, 0] = vocab_lo + 1
    cand[2] = base.cpu(); cand[2, L - 1] = vocab_lo + 2
    if n > 12:                               # exact duplicates: of another candidate and of the base
        cand[9] = cand[4]
        cand[11] = base.cpu()
    return cand.to(device)


@pytest.mark.parametrize("mode,n,n_replace", [("joint", 300, 1), ("gcg_only", 64, 2), ("joint", 4300, 1)])
def test_suffix_head_sharing_matches_full_tails(cuda_device, mode, n, n_replace):
    """Rows a candidate shares with the suffix it was sampled from are computed once (SURVEY 8(f)
    rank 3). Same losses as scoring every full tail (only the attention's key-block order differs)
    and as the fp32 oracle; fewer decoder rows. n = 4300 exceeds one scorer chunk (4096)."""
    from bimodalattack_b200 import BimodalAttack
    from bimodalattack_b200.attack import image_features
    from oracle import hotpath
    model, tok, proc, norm, image, cfg = _setup(cuda_device, search_width=n, **MODES[mode])
    atk = BimodalAttack(model, tok, proc, cfg, normalize=norm)
    atk._prepare([{"role": "user", "content": PROMPT}], TARGET)
    base = tok(INIT, add_special_tokens=False, return_tensors="pt")["input_ids"].to(cuda_device).reshape(-1)
    cand = _related_candidates(base, n, 8, 1000, n_replace, seed=n, device=cuda_device)
    with torch.no_grad():
        feats = image_features(model, norm(image), False) if cfg.pgd_attack else None
        if mode == "joint":
            batch = atk._build_input_embeds_pgd(cand, feats, search_width=n, single=True)
            order = hotpath.segment_order("pgd", "llava", single=True)
        else:
            batch = atk._build_input_embeds_gcg(cand, search_width=n, no_joint_eval=True)
            order = hotpath.segment_order("gcg", "llava", no_joint_eval=True)
        full = atk._compute_candidates_loss_original(n, batch).clone()
        rows_full = atk.engine.last_rows
        _, idx_full = atk._winner()
        atk._base = (cand, base.to(torch.int64).contiguous())
        shared = atk._compute_candidates_loss_original(n, batch).clone()
        rows_shared = atk.engine.last_rows
        key_loss, idx_shared = atk._winner()
    S = atk._plan(batch).n_tail - 1
    L = base.numel()
    first_diff = torch.where((cand != base).any(1), (cand != base).float().argmax(1), torch.tensor(L, device=cuda_device))
    first_diff = first_diff.clamp(max=S - atk.target_ids.numel())     # scored rows are never shared
    n_chunks = (n + 4095) // 4096          # the engine sizes the workspace for 4096 candidates (+ the base) per chunk
    assert rows_full == n * S
    # identical candidates (within a chunk) and copies of the base own no rows; + the base, once per chunk
    want_rows = n_chunks * S
    for lo in range(0, n, 4096):
        chunk = cand[lo:lo + 4096]
        uniq, first = [], {}
        for i, row in enumerate(chunk.tolist()):
            if tuple(row) not in first and row != base.tolist():
                first[tuple(row)] = i
                uniq.append(lo + i)
        want_rows += int((S - first_diff[torch.tensor(uniq, device=cuda_device)]).sum())
    assert rows_shared == want_rows
    assert float(shared[9]) == float(shared[4]) and float(shared[11]) == float(shared[0])   # exact copies
    assert rows_shared < 0.85 * rows_full
    assert float((shared - full).abs().max()) <= 1e-2 * float(full.abs().max())
    assert idx_shared == int(shared.argmin()) and key_loss == float(shared.min())
    assert float(full[idx_shared]) <= float(full.min()) + 1e-2 * float(full.abs().max())
    if n <= 300:
        ora32 = hotpath.OracleAttack(_fp32_twin(model), tok, proc, cfg, normalize=norm)
        ora32.target_ids = atk.target_ids
        ora32.seg_embeds = {k: ora32.emb(v) for k, v in atk.segment_ids.items()}
        f32 = feats.float() if feats is not None else None
        want, _ = hotpath.candidates_loss(ora32.model, hotpath.assemble(ora32._parts(cand, f32), order, n),
                                          atk.target_ids, 50)
        assert bool(((shared - want).abs() <= 2e-2 * want.abs() + 1e-3).all()), (shared - want).abs().max()


@pytest.mark.parametrize("mode", ["joint", "gcg_only"])
def test_graphed_gradient_equals_eager(cuda_device, mode):
    """The batch-1 gradient step is replayed from a CUDA graph after two eager calls: same kernels
    on static buffers, so the results must equal the eager path for fresh inputs."""
    from bimodalattack_b200 import BimodalAttack
    model, tok, proc, norm, image, cfg = _setup(cuda_device, **MODES[mode])
    atk = BimodalAttack(model, tok, proc, cfg, normalize=norm)
    atk._prepare([{"role": "user", "content": PROMPT}], TARGET)
    n_tok = len(tok(INIT, add_special_tokens=False)["input_ids"])
    g = torch.Generator().manual_seed(11)
    for call in range(5):
        ids = torch.randint(8, 1000, (1, n_tok), generator=g).to(cuda_device)
        img = (torch.rand(image.shape, generator=g).to(cuda_device).requires_grad_()) if cfg.pgd_attack else None
        got_tok, got_img = atk.compute_gradient(ids, img)
        img2 = img.detach().clone().requires_grad_() if img is not None else None
        want_tok, want_img = atk._gradient_eager(ids, img2)
        assert torch.equal(got_tok, want_tok), call
        if cfg.pgd_attack:
            assert torch.equal(got_img, want_img), call
    assert atk._grad_graph["graph"] is not None and not atk._grad_graph["failed"]


def test_score_host_entry_and_degenerate_batches(cuda_device):
    """The host-buffer entry (pinned ids in, losses + key out, with the sharing base) equals the
    device entry; B = 1, all candidates identical to the base, and a base unrelated to the
    candidates (nothing shared) all give the full-tail losses."""
    from bimodalattack_b200 import BimodalAttack
    from bimodalattack_b200.engine import unpack_key
    model, tok, proc, norm, image, cfg = _setup(cuda_device, **MODES["joint"])
    atk = BimodalAttack(model, tok, proc, cfg, normalize=norm)
    atk._prepare([{"role": "user", "content": PROMPT}], TARGET)
    base = tok(INIT, add_special_tokens=False, return_tensors="pt")["input_ids"].to(cuda_device).reshape(-1).contiguous()
    cand = _related_candidates(base, 37, 8, 1000, 1, seed=9, device=cuda_device)
    with torch.no_grad():
        feats = atk.scoring_features(image)
        batch = atk._build_input_embeds_pgd(cand, feats, search_width=37, single=True)
        plan = atk._plan(batch)
        # size the workspace once, so both entries below run the same (non-fused) ragged pass on the same
        # prefix K/V and can be compared bit for bit
        atk.engine.reserve(40 * plan.n_tail + plan.n_prefix, plan.n_prefix + 64)
        atk.engine.prefill(plan.prefix, deferred=False)
        full = atk.engine.score(cand, plan, atk.target_ids).clone()
        shared = atk.engine.score(cand, plan, atk.target_ids, base_ids=base).clone()
        key_dev = int(atk.engine.key.item())
        loss_h = torch.empty(37, dtype=torch.float32).pin_memory()
        key_h = torch.empty(1, dtype=torch.int64).pin_memory()
        atk.engine.score_host(cand.cpu().pin_memory(), plan, atk.target_ids, loss_h, key_h,
                              base_ids_host=base.cpu().pin_memory())
        assert torch.equal(loss_h, shared.cpu()) and int(key_h.item()) == key_dev
        assert unpack_key(key_dev)[1] == int(shared.argmin())
        tol = 1e-2 * float(full.abs().max())
        assert float((shared - full).abs().max()) <= tol
        one = atk.engine.score(cand[5:6], plan, atk.target_ids, base_ids=base)
        assert abs(float(one[0]) - float(full[5])) <= tol
        same = atk.engine.score(base.repeat(6, 1), plan, atk.target_ids, base_ids=base)
        assert float((same - full[0]).abs().max()) <= tol and float(same.max() - same.min()) == 0.0
        unrelated = torch.randint(8, 1000, base.shape, generator=torch.Generator().manual_seed(2)).to(cuda_device)
        none_shared = atk.engine.score(cand, plan, atk.target_ids, base_ids=unrelated.contiguous())
        # no shared suffix head; only the two exact duplicates inside the batch (rows 9 and 11) are saved,
        # and the (unrelated) base rides along as one extra sequence
        assert atk.engine.last_rows == (37 - 2 + 1) * (plan.n_tail - 1)
        assert float((none_shared - full).abs().max()) <= tol


def test_config1_tiny_model(cuda_device):
    """BASELINE.json config 1 -- the reference's CPU-sized case (2 layers, hidden 256, 4 heads of 64,
    V = 1000, 4 image tokens; joint GCG+PGD, 20 steps, search_width 64) -- on the native path in
    bf16: candidate losses against the oracle, then the full 20-step run."""
    import torchvision.transforms as T
    import bimodalattack_b200 as bma
    from bimodalattack_b200 import synthetic as syn
    from oracle import hotpath
    tok = syn.make_tokenizer(1000)
    proc = syn.make_llava_processor(tok, 28)
    model = syn.make_llava("tiny", dtype=torch.bfloat16, device=cuda_device, head_gain=20.0, attn_implementation="sdpa")
    norm = T.Normalize(syn.CLIP_MEAN, syn.CLIP_STD)
    image = torch.rand((1, 3, 28, 28), generator=torch.Generator().manual_seed(5)).to(cuda_device)
    cfg = bma.BimodalAttackConfig(num_steps=20, search_width=64, topk=32, seed=1, allow_non_ascii=True, alpha=4 / 255,
                                  eps=64 / 255, pgd_attack=True, gcg_attack=True, joint_eval=True,
                                  optim_str_init="x x x x x x x x", verbosity="ERROR", images_folder=None)
    msgs = [{"role": "user", "content": "w10 w11 w12 w13"}]
    atk = bma.BimodalAttack(model, tok, proc, cfg, normalize=norm)
    assert atk.engine.head_dim == 64
    atk._prepare(msgs, "w20 w21 w22 w23 w24")
    cand = torch.randint(8, 990, (64, 8), generator=torch.Generator().manual_seed(2)).to(cuda_device)
    with torch.no_grad():
        feats = atk.scoring_features(image)
        got = atk._compute_candidates_loss_original(64, atk._build_input_embeds_pgd(cand, feats, 64, single=True))
        ora32 = hotpath.OracleAttack(_fp32_twin(model), tok, proc, cfg, normalize=norm)
        ora32.target_ids = atk.target_ids
        ora32.seg_embeds = {k: ora32.emb(v) for k, v in atk.segment_ids.items()}
        order = hotpath.segment_order("pgd", "llava", single=True)
        want, _ = hotpath.candidates_loss(ora32.model, hotpath.assemble(ora32._parts(cand, feats.float()), order, 64),
                                          atk.target_ids, 64)
    assert bool(((got - want).abs() <= 2e-2 * want.abs() + 1e-3).all()), (got - want).abs().max()
    res = bma.BimodalAttack(model, tok, proc, cfg, normalize=norm).run(msgs, "goal", "w20 w21 w22 w23 w24", image.clone())
    assert len(res.losses) == 20 and res.best_loss == min(res.losses) and res.losses[-1] < res.losses[0]
