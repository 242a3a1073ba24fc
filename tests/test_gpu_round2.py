"""GPU parity tests added in round 2: the fused (one row space) prefill, the multi-rank split of the
prefix (emulated ranks on one device), the early-stop flag, ``filter_ids`` actually dropping rows,
``run()`` in the option combinations round 1 never exercised, and sampled ids on bf16 gradients.
All through the C ABI; the oracle (oracle/hotpath.py, pinned to the reference by tests/golden) is the
checker."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu

PROMPT = "w10 w11 w12 w13 w14 w15"
TARGET = "w20 w21 w22 w23 w24"
INIT = "x x x x x x x x x x"
JOINT = dict(pgd_attack=True, gcg_attack=True, joint_eval=True)


def _setup(device, size="tiny128", image_px=28, **cfg_kw):
    import torchvision.transforms as T
    from bimodalattack_b200 import BimodalAttackConfig, synthetic as syn
    tok = syn.make_tokenizer(1024 if size == "tiny128" else 4096)
    proc = syn.make_llava_processor(tok, image_px)
    model = syn.make_llava(size, dtype=torch.bfloat16, device=device, head_gain=20.0, attn_implementation="sdpa")
    norm = T.Normalize(syn.CLIP_MEAN, syn.CLIP_STD)
    image = torch.rand((1, 3, image_px, image_px), generator=torch.Generator().manual_seed(5)).to(device)
    kw = dict(num_steps=3, search_width=48, topk=32, seed=1, allow_non_ascii=True, alpha=4 / 255, eps=64 / 255,
              optim_str_init=INIT, verbosity="ERROR", images_folder=None, filter_ids=True)
    kw.update(cfg_kw)
    return model, tok, proc, norm, image, BimodalAttackConfig(**kw)


def _related(base, n, lo, hi, seed, device):
    g = torch.Generator().manual_seed(seed)
    cand = base.cpu().repeat(n, 1)
    pos = torch.randint(0, base.numel(), (n,), generator=g)
    cand[torch.arange(n), pos] = torch.randint(lo, hi, (n,), generator=g)
    cand[0] = base.cpu()
    return cand.to(device)


def _oracle32_losses(model, tok, proc, cfg, norm, atk, cand, feats, order):
    from oracle import hotpath
    ora = hotpath.OracleAttack(copy.deepcopy(model).float(), tok, proc, cfg, normalize=norm)
    ora.target_ids = atk.target_ids
    ora.seg_embeds = {k: ora.emb(v) for k, v in atk.segment_ids.items()}
    f32 = None if feats is None else feats.float()
    want, _ = hotpath.candidates_loss(ora.model, hotpath.assemble(ora._parts(cand, f32), order, cand.shape[0]),
                                      atk.target_ids, 64)
    return want.float()


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("size,px,n", [("tiny128", 28, 70), ("small128", 56, 333)])
def test_fused_prefill_matches_separate_prefill(cuda_device, size, px, n):
    """The prefix positions ride in the first chunk of the score call as extra rows ([prefix | base |
    candidates], one pass) instead of a separate prefill pass: same losses up to the attention's
    key-block order, same winner, and the fp32 oracle agrees with both."""
    from bimodalattack_b200 import BimodalAttack
    from bimodalattack_b200.attack import image_features
    from oracle import hotpath
    model, tok, proc, norm, image, cfg = _setup(cuda_device, size=size, image_px=px, search_width=n, **JOINT)
    atk = BimodalAttack(model, tok, proc, cfg, normalize=norm)
    atk._prepare([{"role": "user", "content": PROMPT}], TARGET)
    base = tok(INIT, add_special_tokens=False, return_tensors="pt")["input_ids"].to(cuda_device).reshape(-1).contiguous()
    cand = _related(base, n, 8, 1000, seed=n, device=cuda_device)
    eng = atk.engine
    with torch.no_grad():
        feats = image_features(model, norm(image), False)
        batch = atk._build_input_embeds_pgd(cand, feats, search_width=n, single=True)
        plan = atk._plan(batch)
        eng.prefill(plan.prefix, deferred=False)
        sep = eng.score(cand, plan, atk.target_ids, base_ids=base).clone()
        rows_sep = eng.last_rows
        key_sep = int(eng.key.item())
        eng.prefill(plan.prefix, deferred=True)
        fused = eng.score(cand, plan, atk.target_ids, base_ids=base).clone()
        rows_fused = eng.last_rows
        key_fused = int(eng.key.item())
        # the K/V the fused pass published serve later calls exactly like a prefill pass would
        again = eng.score(cand, plan, atk.target_ids, base_ids=base).clone()
    assert rows_sep == rows_fused                     # tail rows only; the prefix rows are not counted
    tol = 1e-2 * float(sep.abs().max())
    assert float((fused - sep).abs().max()) <= tol
    assert float((again - fused).abs().max()) <= tol     # (row offsets move the attention's key-block boundaries)
    from bimodalattack_b200.engine import unpack_key
    assert unpack_key(key_fused)[1] == int(fused.argmin())
    assert float(sep[unpack_key(key_fused)[1]]) <= float(sep.min()) + tol and key_sep == key_sep
    order = hotpath.segment_order("pgd", "llava", single=True)
    want = _oracle32_losses(model, tok, proc, cfg, norm, atk, cand[:64], feats, order)
    assert bool(((fused[:64] - want).abs() <= 2e-2 * want.abs() + 1e-3).all()), (fused[:64] - want).abs().max()


@pytest.mark.parametrize("peers", [True, False])
@pytest.mark.parametrize("ranks,n", [(2, 48), (3, 50), (4, 3), (8, 130)])
def test_split_prefill_over_emulated_ranks(cuda_device, ranks, n, peers):
    """Multi-rank scoring on ONE device (bma_scorer_group_score): the candidates are sharded like
    dist.shard_bounds, the prefix positions are split over the ranks and every layer's K/V slices are
    exchanged -- peers=True: the peer-memory kernels (scatter into every scorer's prefix buffers + flag
    signal / wait, plain pointers standing in for the cudaIpc mappings); peers=False: the pack / unpack
    kernels around local copies standing in for the NCCL all-gather. Every row is computed from the same
    K/V as on a single rank (only its offset inside the row space, hence the attention's key-block
    boundaries, differs), so losses agree with the single-rank fused pass to bf16 re-association noise,
    including ranks that own no candidate at all (n < ranks)."""
    from bimodalattack_b200 import BimodalAttack
    from bimodalattack_b200.attack import image_features
    from bimodalattack_b200.engine import ScoringEngine, group_score, unpack_key
    model, tok, proc, norm, image, cfg = _setup(cuda_device, search_width=n, **JOINT)
    atk = BimodalAttack(model, tok, proc, cfg, normalize=norm)
    atk._prepare([{"role": "user", "content": PROMPT}], TARGET)
    base = tok(INIT, add_special_tokens=False, return_tensors="pt")["input_ids"].to(cuda_device).reshape(-1).contiguous()
    cand = _related(base, n, 8, 1000, seed=ranks, device=cuda_device)
    with torch.no_grad():
        feats = image_features(model, norm(image), False)
        plan = atk._plan(atk._build_input_embeds_pgd(cand, feats, search_width=n, single=True))
        atk.engine.prefill(plan.prefix, deferred=True)
        single = atk.engine.score(cand, plan, atk.target_ids, base_ids=base).clone()
        key_single = int(atk.engine.key.item())
        group = [ScoringEngine(model) for _ in range(ranks)]
        for e in group:
            e.prefill(plan.prefix, deferred=True)
        split = group_score(group, cand, plan, atk.target_ids, base, use_peers=peers).clone()
        key_split = int(group[0].key.item())
        # afterwards every emulated rank holds the complete prefix K/V: a plain score call on any of them agrees
        later = group[-1].score(cand, plan, atk.target_ids, base_ids=base).clone()
    tol = 1e-2 * float(single.abs().max())
    assert float((split - single).abs().max()) <= tol
    assert unpack_key(key_split)[1] == int(split.argmin()) and unpack_key(key_split)[0] == float(split.min())
    assert float(single[unpack_key(key_split)[1]]) <= float(single.min()) + tol
    assert unpack_key(key_single)[1] == int(single.argmin())
    assert float((later - single).abs().max()) <= tol
    for e in group:
        e.close()


@pytest.mark.parametrize("ranks,n,size,px", [(1, 70, "tiny128", 28), (2, 48, "tiny128", 28), (8, 130, "tiny128", 28),
                                             (3, 200, "small128", 56)])
def test_fused_publish_is_bit_identical_to_separate_kernels(cuda_device, ranks, n, size, px):
    """The QKV GEMM's epilogue stores the prefix rows' and the base rows' K/V into the per-layer prefix
    buffers of every (emulated) rank itself; the separate append / scatter kernels it replaces copy the
    same bf16 values afterwards. Same row space, same values: losses and winner keys are bit-identical."""
    from bimodalattack_b200 import BimodalAttack
    from bimodalattack_b200.attack import image_features
    from bimodalattack_b200.engine import ScoringEngine, group_score
    model, tok, proc, norm, image, cfg = _setup(cuda_device, size=size, image_px=px, search_width=n, **JOINT)
    atk = BimodalAttack(model, tok, proc, cfg, normalize=norm)
    atk._prepare([{"role": "user", "content": PROMPT}], TARGET)
    base = tok(INIT, add_special_tokens=False, return_tensors="pt")["input_ids"].to(cuda_device).reshape(-1).contiguous()
    cand = _related(base, n, 8, 1000, seed=ranks, device=cuda_device)
    out = {}
    with torch.no_grad():
        feats = image_features(model, norm(image), False)
        plan = atk._plan(atk._build_input_embeds_pgd(cand, feats, search_width=n, single=True))
        for fused in (True, False):
            group = [ScoringEngine(model) for _ in range(ranks)]
            for e in group:
                e.fused_publish = fused
                e.prefill(plan.prefix, deferred=True)
            if ranks == 1:
                loss = group[0].score(cand, plan, atk.target_ids, base_ids=base).clone()
                again = group[0].score(cand, plan, atk.target_ids, base_ids=base).clone()   # base rows only
            else:
                loss = group_score(group, cand, plan, atk.target_ids, base, use_peers=True).clone()
                again = group[-1].score(cand, plan, atk.target_ids, base_ids=base).clone()
            out[fused] = (loss, again, int(group[0].key.item()))
            for e in group:
                e.close()
    assert torch.equal(out[True][0], out[False][0])
    assert torch.equal(out[True][1], out[False][1])
    assert out[True][2] == out[False][2]


@pytest.mark.parametrize("ranks,n,size,px", [(2, 48, "tiny128", 28), (8, 130, "tiny128", 28), (3, 200, "small128", 56)])
def test_early_signal_is_bit_identical(cuda_device, ranks, n, size, px):
    """Opt-in early signal (bma_scorer_set_early_signal): the QKV GEMM walks its publishing tiles first and raises the
    rank's flag from its own epilogue, the exchange launch only waits. Tile ORDER changes, values do not: losses and
    the winner key equal the default path's bit for bit; the wait accumulator and the bare barrier are callable on a
    scorer without peers (0 ms, no-op)."""
    from bimodalattack_b200 import BimodalAttack
    from bimodalattack_b200.attack import image_features
    from bimodalattack_b200.engine import ScoringEngine, group_score
    model, tok, proc, norm, image, cfg = _setup(cuda_device, size=size, image_px=px, search_width=n, **JOINT)
    atk = BimodalAttack(model, tok, proc, cfg, normalize=norm)
    atk._prepare([{"role": "user", "content": PROMPT}], TARGET)
    base = tok(INIT, add_special_tokens=False, return_tensors="pt")["input_ids"].to(cuda_device).reshape(-1).contiguous()
    cand = _related(base, n, 8, 1000, seed=ranks, device=cuda_device)
    out = {}
    with torch.no_grad():
        feats = image_features(model, norm(image), False)
        plan = atk._plan(atk._build_input_embeds_pgd(cand, feats, search_width=n, single=True))
        for early in (True, False):
            group = [ScoringEngine(model) for _ in range(ranks)]
            for e in group:
                e.early_signal = early
                e.prefill(plan.prefix, deferred=True)
            loss = group_score(group, cand, plan, atk.target_ids, base, use_peers=True).clone()
            again = group[-1].score(cand, plan, atk.target_ids, base_ids=base).clone()
            out[early] = (loss, again, int(group[0].key.item()))
            assert group[0].exchange_wait_ms(reset=True) >= 0.0
            group[0].peers_barrier(3)            # no peers attached any more: a no-op
            for e in group:
                e.close()
    assert torch.equal(out[True][0], out[False][0])
    assert torch.equal(out[True][1], out[False][1])
    assert out[True][2] == out[False][2]


# ------------------------------------------------------------------------------------------------
def _greedy_target(model, atk, cand_row, feats, n_tokens):
    """Target ids the model itself predicts greedily after [prefix, cand_row, after]: with them as the
    target, that candidate matches everywhere (reference early_stop, :1300-1306)."""
    emb = atk.embedding_layer
    e = atk.segment_embeds
    seq = torch.cat([e["before_img"], feats, e["before_suffix"], emb(cand_row.view(1, -1)), e["after"]], dim=1)
    out = []
    for _ in range(n_tokens):
        logits = model(inputs_embeds=seq).logits[0, -1]
        nxt = int(torch.argmax(logits))
        out.append(nxt)
        seq = torch.cat([seq, emb(torch.tensor([[nxt]], device=seq.device))], dim=1)
    return out


def test_early_stop_flag_matches_reference_argmax(cuda_device):
    """early_stop: the flag is raised iff some candidate's argmax equals the target at every scored
    position (torch.argmax's first-maximum rule), here against the oracle on the same bf16 model."""
    from bimodalattack_b200 import BimodalAttack
    from bimodalattack_b200.attack import image_features
    from oracle import hotpath
    model, tok, proc, norm, image, cfg = _setup(cuda_device, search_width=40, early_stop=True, **JOINT)
    atk = BimodalAttack(model, tok, proc, cfg, normalize=norm)
    atk._prepare([{"role": "user", "content": PROMPT}], TARGET)
    base = tok(INIT, add_special_tokens=False, return_tensors="pt")["input_ids"].to(cuda_device).reshape(-1).contiguous()
    cand = _related(base, 40, 8, 1000, seed=4, device=cuda_device)
    order = hotpath.segment_order("pgd", "llava", single=True)
    with torch.no_grad():
        feats = image_features(model, norm(image), False)
        greedy = _greedy_target(model, atk, cand[17], feats, atk.target_ids.numel())
        ora = hotpath.OracleAttack(model, tok, proc, cfg, normalize=norm)
        ora.seg_embeds = {k: ora.emb(v) for k, v in atk.segment_ids.items()}
        for target, expect in ((greedy, True), ([(t + 1) % 1000 for t in greedy], False)):
            tgt = torch.tensor([target], device=cuda_device)
            atk.target_ids = tgt
            atk.segment_ids["target"] = tgt
            atk.segment_embeds["target"] = atk.embedding_layer(tgt)
            ora.target_ids = tgt
            ora.seg_embeds["target"] = ora.emb(tgt)
            atk._plans.clear()
            atk.stop_flag = False
            for share in (False, True):
                atk._base = (cand, base) if share else None
                batch = atk._build_input_embeds_pgd(cand, feats, search_width=40, single=True)
                got = atk._compute_candidates_loss_original(40, batch)
                _, matched = hotpath.candidates_loss(model, hotpath.assemble(ora._parts(cand, feats), order, 40),
                                                     tgt, 40, early_stop=True)
                assert bool(matched) == expect
                assert atk.stop_flag == expect, (share, expect)
                assert int(atk.engine.flag.item()) == int(expect)
                assert got.shape[0] == 40
                atk.stop_flag = False


class _MergingTokenizer:
    """Tokenizer whose decode -> encode round trip FAILS for ids in ``bad`` (they decode to the text of
    the next id), like BPE merges do for real vocabularies (reference filter_ids, :166-186)."""

    def __init__(self, tok, bad):
        self._tok, self._bad = tok, set(bad)

    def __getattr__(self, name):
        return getattr(self._tok, name)

    def __call__(self, *a, **k):
        return self._tok(*a, **k)

    def _fix(self, ids):
        return [int(i) + 1 if int(i) in self._bad else int(i) for i in ids]

    def decode(self, ids, **k):
        return self._tok.decode(self._fix(ids.tolist() if hasattr(ids, "tolist") else ids), **k)

    def batch_decode(self, seqs, **k):
        seqs = seqs.tolist() if hasattr(seqs, "tolist") else seqs
        return [self._tok.decode(self._fix(s), **k) for s in seqs]


def test_filter_ids_drops_rows_and_scorer_takes_any_width(cuda_device):
    """filter_ids really drops candidates (sw' < sw, odd count): same surviving rows as the oracle's
    per-row loop, and the scorer (ragged pass, sharing base) scores exactly those."""
    from bimodalattack_b200 import BimodalAttack
    from bimodalattack_b200.attack import filter_ids, image_features
    from oracle import hotpath
    model, tok0, proc, norm, image, cfg = _setup(cuda_device, search_width=61, **JOINT)
    tok = _MergingTokenizer(tok0, bad=range(100, 1000, 3))
    atk = BimodalAttack(model, tok, proc, cfg, normalize=norm)
    atk._prepare([{"role": "user", "content": PROMPT}], TARGET)
    base = tok(INIT, add_special_tokens=False, return_tensors="pt")["input_ids"].to(cuda_device).reshape(-1).contiguous()
    cand = _related(base, 61, 8, 1000, seed=8, device=cuda_device)
    kept = filter_ids(cand, tok)
    want_kept = hotpath.filter_ids(cand, tok)
    assert torch.equal(kept, want_kept)
    assert 0 < kept.shape[0] < 61
    with torch.no_grad():
        feats = image_features(model, norm(image), False)
        atk._base = (kept, base)
        batch = atk._build_input_embeds_pgd(kept, feats, search_width=kept.shape[0], single=True)
        got = atk._compute_candidates_loss_original(kept.shape[0], batch)
        order = hotpath.segment_order("pgd", "llava", single=True)
        want = _oracle32_losses(model, tok, proc, cfg, norm, atk, kept, feats, order)
    assert got.shape[0] == kept.shape[0]
    assert bool(((got - want).abs() <= 2e-2 * want.abs() + 1e-3).all())
    # nothing survives -> the reference's RuntimeError
    only_bad = torch.full((4, base.numel()), 100, device=cuda_device)
    with pytest.raises(RuntimeError):
        filter_ids(only_bad, tok)


# ------------------------------------------------------------------------------------------------
RUN_CASES = {
    # the oracle golden's "joint_masked_dynamic": non-ASCII mask + shrinking search width
    "joint_masked_dynamic": dict(JOINT, allow_non_ascii=False, dynamic_search=True, min_search_width=16),
    "buffer4": dict(JOINT, buffer_size=4),
    "pgd_after_gcg": dict(JOINT, pgd_after_gcg=True),
    "gcg_buffer3_nofilter": dict(pgd_attack=False, gcg_attack=True, buffer_size=3, filter_ids=False),
    "early_stop_on": dict(JOINT, early_stop=True),
}


@pytest.mark.parametrize("case", list(RUN_CASES))
def test_run_option_combinations(cuda_device, case, tmp_path):
    import bimodalattack_b200 as bma
    from bimodalattack_b200.utils import get_nonascii_toks
    model, tok, proc, norm, image, cfg = _setup(cuda_device, **RUN_CASES[case])
    cfg.num_steps = 5
    atk = bma.BimodalAttack(model, tok, proc, cfg, normalize=norm)
    widths = []
    real = atk.candidate_sampling

    def spy(*a, **k):
        out = real(*a, **k)
        widths.append(int(out[1]))
        if not cfg.allow_non_ascii and cfg.gcg_attack:
            banned = get_nonascii_toks(tok, device=cuda_device)
            changed = out[0][(out[0] != a[2].reshape(1, -1))]
            assert not bool(torch.isin(changed, banned).any())
        return out
    atk.candidate_sampling = spy
    res = atk.run([{"role": "user", "content": PROMPT}], "goal", TARGET, image.clone() if cfg.pgd_attack else None)
    assert len(res.losses) == 5 and all(l == l and l > 0 for l in res.losses)
    assert res.best_loss == min(res.losses)
    assert res.losses[-1] <= res.losses[0] + 0.5
    assert len(res.gradient_times) >= 5 and len(res.loss_times) >= 5
    assert all(t > 0 for t in res.gradient_times + res.loss_times)        # device times from CUDA events
    if cfg.dynamic_search:
        assert widths[0] == 48 and widths[-1] < widths[0] and min(widths) >= cfg.min_search_width
    if cfg.pgd_attack:
        final = atk.final_image.detach()
        assert float((final - image).abs().max()) <= cfg.eps + 1e-6 and 0.0 <= float(final.min()) and float(final.max()) <= 1.0
    # the model's own forward and torch's global flag are back as the caller left them
    assert not any(hasattr(m, "_bma_orig_forward") for m in model.modules())


@pytest.mark.parametrize("case", ["buffer4", "gcg_buffer3_nofilter"])
def test_initial_buffer_matches_oracle(cuda_device, case):
    """buffer_size > 1 (:826-906): the extra entries are drawn with ONE torch.randint on the default generator (which
    shifts every later draw), all entries are scored and the run starts from the best one. The oracle's buffer path
    is pinned to the unmodified reference by the golden cases joint_buffer3 / gcg_buffer2; here the product must
    draw the same entries, score them to bf16 tolerance and start from the (tie-aware) best."""
    import bimodalattack_b200 as bma
    from oracle import hotpath
    model, tok, proc, norm, image, cfg = _setup(cuda_device, **RUN_CASES[case])
    cfg.num_steps = 1
    atk = bma.BimodalAttack(model, tok, proc, cfg, normalize=norm)
    seen = {}
    real = atk.init_buffer

    def spy(img):
        buf = real(img)
        seen["entries"] = [(float(l), ids.reshape(-1).tolist()) for l, ids in buf.buffer]
        return buf
    atk.init_buffer = spy
    res = atk.run([{"role": "user", "content": PROMPT}], "goal", TARGET, image.clone() if cfg.pgd_attack else None)
    ora = hotpath.OracleAttack(model, tok, proc, cfg, normalize=norm)
    log = ora.run([{"role": "user", "content": PROMPT}], TARGET, image.clone() if cfg.pgd_attack else None)
    want_rows = sorted(tuple(ids.reshape(-1).tolist()) for _, ids in ora.init_buffer)
    got_rows = sorted(tuple(ids) for _, ids in seen["entries"])
    assert got_rows == want_rows and len(want_rows) == cfg.buffer_size       # the same random rows (repeats included)
    want = {tuple(ids.reshape(-1).tolist()): float(l) for l, ids in ora.init_buffer}
    for l, ids in seen["entries"]:
        assert abs(l - want[tuple(ids)]) <= 0.0625 * 2 + 2e-2 * abs(want[tuple(ids)])
    assert want[tuple(seen["entries"][0][1])] <= min(want.values()) + 0.0625 * 2     # starts from the best entry
    # (the step itself is covered elsewhere; HF's backward for this tiny tower is not run-to-run deterministic, so the
    # sampled sets may differ by a tie: a loose bound only)
    assert abs(res.losses[0] - log.losses[0]) <= 0.0625 * 2 + 5e-2 * abs(log.losses[0])


def test_run_restores_process_state(cuda_device):
    import bimodalattack_b200 as bma
    model, tok, proc, norm, image, cfg = _setup(cuda_device, **JOINT)
    cfg.num_steps = 1
    before = torch.utils.deterministic.fill_uninitialized_memory
    bma.run(model, tok, proc, [{"role": "user", "content": PROMPT}], "goal", TARGET, image.clone(), cfg, normalize=norm)
    assert torch.utils.deterministic.fill_uninitialized_memory == before
    assert not any(hasattr(m, "_bma_orig_forward") for m in model.modules())
    torch.use_deterministic_algorithms(False)


# ------------------------------------------------------------------------------------------------
def test_sample_ids_bit_exact_on_tie_free_bf16(cuda_device):
    """The product feeds the sampler bf16 gradients. With tie-free bf16 rows (32 064 distinct finite
    bf16 values per row) the candidate ids equal the reference's on CUDA, bit for bit."""
    from bimodalattack_b200 import native
    from oracle import hotpath
    V, ntok, sw, topk = 32064, 20, 512, 256
    g = torch.Generator().manual_seed(3)
    # all finite bf16 bit patterns (sign x exponent < 255 x mantissa), -0 removed, shuffled per row
    pats = torch.arange(0, 1 << 16, dtype=torch.int32)
    pats = pats[((pats >> 7) & 0xFF) != 0xFF]
    pats = pats[pats != 0x8000]
    rows = []
    for _ in range(ntok):
        pick = pats[torch.randperm(pats.numel(), generator=g)[:V]]
        rows.append(pick.to(torch.int16).view(torch.bfloat16))
    grad = torch.stack(rows).to(cuda_device)
    assert all(grad[r].float().unique().numel() == V for r in range(0, ntok, 7))
    ids = torch.randint(5, V, (ntok,), generator=g).to(cuda_device)
    not_allowed = torch.randperm(V, generator=g)[: V // 4].to(cuda_device)
    gen = torch.cuda.default_generators[cuda_device.index]
    torch.manual_seed(5)
    seed, off = gen.initial_seed(), gen.get_offset()
    want = hotpath.sample_ids_from_grad(ids.clone(), grad.clone(), sw, topk, 1, not_allowed)
    got, adv, _ = native.sample_ids(ids, grad, sw, topk, 1, native.pack_mask_bits(not_allowed, V, cuda_device), seed, off)
    assert torch.equal(got, want) and off + adv == gen.get_offset()


@pytest.mark.parametrize("V", [32064, 262208])
def test_sample_ids_on_tied_bf16_rows_agree_outside_tie_groups(cuda_device, V):
    """Real bf16 gradients are full of ties, where torch.topk's order is unspecified (the kernel breaks
    them by lower token id). The RNG stream is the reference's, so every candidate replaces the SAME
    position with the token of the SAME top-k rank: ids must be equal wherever that rank's gradient
    value is unique in its row, and otherwise come from the same tie group."""
    from bimodalattack_b200 import native
    from oracle import hotpath
    ntok, sw, topk = 20, 512, 256
    g = torch.Generator().manual_seed(11)
    grad = (torch.randn((ntok, V), generator=g) * 1e-3).to(cuda_device).to(torch.bfloat16)
    ids = torch.randint(5, V, (ntok,), generator=g).to(cuda_device)
    gen = torch.cuda.default_generators[cuda_device.index]
    torch.manual_seed(6)
    seed, off = gen.initial_seed(), gen.get_offset()
    want = hotpath.sample_ids_from_grad(ids.clone(), grad.clone(), sw, topk, 1, None)
    got, adv, _ = native.sample_ids(ids, grad, sw, topk, 1, None, seed, off)
    assert off + adv == gen.get_offset()
    # the reference changed exactly one position per candidate (or re-drew the current token)
    pos_w = (want != ids).float().argmax(1)
    changed_w = (want != ids).any(1)
    diff = (got != want)
    assert bool((diff.sum(1) <= 1).all())
    rows = torch.arange(sw, device=cuda_device)
    tok_w, tok_g = want[rows, pos_w], got[rows, pos_w]
    gw, gg = grad[pos_w, tok_w].float(), grad[pos_w, tok_g].float()
    assert torch.equal(gw[changed_w], gg[changed_w])                    # same gradient value: same tie group
    n_equal = (grad[pos_w].float() == gw.unsqueeze(1)).sum(1)           # size of that tie group in the row
    unique_rank = changed_w & (n_equal == 1)
    assert bool(unique_rank.any())
    assert torch.equal(got[unique_rank], want[unique_rank])
    # wherever the ids differ, it is inside a tie group, at the position the reference replaced
    bad = diff.any(1)
    assert bool((n_equal[bad] > 1).all())


# ------------------------------------------------------------------------------------------------
def test_fp16_model_is_accepted_and_scored_to_bf16_tolerance(cuda_device):
    """The reference loads LLaVA / LLaVA-RC in float16 (utils/experiments_utils.py:91-100). The library
    computes in bf16: an fp16 model object is accepted as it is (bf16 copies of the decoder weights, HF
    autograd + HF tower for the gradient step), per-candidate losses agree with the reference dataflow run
    on the fp16 model itself and with an fp32 twin to bf16 tolerance, the winner is the fp32 argmin up to
    that tolerance, and run() completes with finite, non-increasing best losses."""
    import copy
    import torchvision.transforms as T
    from bimodalattack_b200 import BimodalAttack, BimodalAttackConfig, synthetic as syn
    from bimodalattack_b200.attack import image_features
    from oracle import hotpath
    tok = syn.make_tokenizer(1024)
    proc = syn.make_llava_processor(tok, 28)
    model = syn.make_llava("tiny128", dtype=torch.float16, device=cuda_device, head_gain=20.0, attn_implementation="sdpa")
    assert next(model.parameters()).dtype == torch.float16
    norm = T.Normalize(syn.CLIP_MEAN, syn.CLIP_STD)
    image = torch.rand((1, 3, 28, 28), generator=torch.Generator().manual_seed(5)).to(cuda_device)
    cfg = BimodalAttackConfig(num_steps=3, search_width=48, topk=32, seed=1, allow_non_ascii=True, alpha=4 / 255,
                              eps=64 / 255, optim_str_init=INIT, verbosity="ERROR", images_folder=None, **JOINT)
    atk = BimodalAttack(model, tok, proc, cfg, normalize=norm)
    atk._prepare([{"role": "user", "content": PROMPT}], TARGET)
    assert atk.engine.model_dtype == torch.float16 and atk.engine.dtype == torch.bfloat16
    assert atk._native_gradient() is None and atk._native_tower() is None       # HF paths for the fp16 object
    base = tok(INIT, add_special_tokens=False, return_tensors="pt")["input_ids"].to(cuda_device).reshape(-1).contiguous()
    cand = _related(base, 48, 8, 1000, seed=3, device=cuda_device)
    with torch.no_grad():
        feats = image_features(model, norm(image), False)
        assert feats.dtype == torch.float16
        batch = atk._build_input_embeds_pgd(cand, feats, search_width=48, single=True)
        got = atk._compute_candidates_loss_original(48, batch).clone()
        key_loss, key_idx = atk._winner()
        order = hotpath.segment_order("pgd", "llava", single=True)
        ora16 = hotpath.OracleAttack(model, tok, proc, cfg, normalize=norm)
        ora16.target_ids = atk.target_ids
        ora16.seg_embeds = {k: ora16.emb(v) for k, v in atk.segment_ids.items()}
        want16, _ = hotpath.candidates_loss(model, hotpath.assemble(ora16._parts(cand, feats), order, 48), atk.target_ids, 16)
        want32 = _oracle32_losses(model, tok, proc, cfg, norm, atk, cand, feats, order)
    want16 = want16.float()
    assert bool(torch.isfinite(got).all())
    assert bool(((got - want32).abs() <= 2e-2 * want32.abs() + 1e-3).all()), (got - want32).abs().max()
    assert bool(((got - want16).abs() <= 2e-2 * want16.abs() + 2e-3).all()), (got - want16).abs().max()
    assert key_idx == int(got.argmin())
    assert float(want32[key_idx]) <= float(want32.min()) + 2e-2 * abs(float(want32.min())) + 1e-3
    res = BimodalAttack(model, tok, proc, cfg, normalize=norm).run([{"role": "user", "content": PROMPT}], "goal", TARGET,
                                                                   image.clone())
    assert len(res.losses) == 3 and all(l == l and abs(l) != float("inf") for l in res.losses)
    assert next(model.parameters()).dtype == torch.float16                      # the caller's object is untouched
