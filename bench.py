#!/usr/bin/env python
"""Headline benchmark: GCG+PGD candidates scored/sec (and steps/sec), LLaVA-1.5-7B architecture,
search_width 512, joint_eval (BASELINE.json config 2). Prints ONE JSON line on rank 0.

  python bench.py --gpus 1 --steps 10 --warmup 3            # this repo's B200 path
  torchrun ... bench.py --gpus N ...                         # candidates sharded over N ranks
  python bench.py --impl reference --steps 2 --warmup 1      # the reference's dataflow on host cores

What a "step" is: the scoring phase of one attack step -- from "candidate
ids and the perturbed image are ready" to "the winner index is known on the host": vision tower +
projector on the new image, shared-prefix KV build, tail-only decoder pass over every candidate,
fused LM-head cross-entropy, argmin, and (N > 1) the min all-reduce.
  value  candidates/s with ids and image already resident in HBM
  e2e    same through the host-buffer C-ABI call: pinned ids + image copied in, losses + key out
Also reported: full attack steps/s (gradient + PGD + sampling + scoring through ``run()``), the
roofline of the dominant kernel class (tcgen05 GEMMs, CUDA events inside the timed steps), and the
oracle (reference dataflow) timed on the host CPU for a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# LM-head scale of the random-init models: logit std ~ 0.02 * gain * sqrt(d), chosen so per-candidate
# losses are spread over a few nats instead of all being ln(V) (SURVEY.md section 7 "Ties")
HEAD_GAIN = {"7b": 2.0, "small128": 10.0, "tiny128": 20.0, "tiny": 20.0}
METRIC = "gcg_pgd_candidates_scored_per_sec"
UNIT = "candidates/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--model", default="7b", choices=["7b", "small128", "tiny128", "tiny", "gemma4b", "gemma-tiny256", "gemma4b-joint",
                                                      "gemma-tiny256-joint"],
                    help="7b = BASELINE.json config 2 (the headline); gemma4b = config 4 (Gemma-3-4b GCG-only); "
                         "tiny = the CPU-sized model of config 1 (use --search-width 64)")
    ap.add_argument("--search-width", type=int, default=512)
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong (default, BASELINE.json north_star): search_width candidates in TOTAL, each of the N "
                         "GPUs scores search_width/N of them; weak: search_width candidates PER GPU")
    ap.add_argument("--full-steps", type=int, default=3, help="full attack steps timed for steps/s (0 = skip)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--multi-gpu-ab", action="store_true",
                    help="N > 1: also time the two opt-in multi-GPU features in this process (early signal from inside the "
                         "QKV GEMM, straggler-aware candidate shares) and report early_signal_ab / adaptive_shares_ab")
    ap.add_argument("--no-hbm-kernels", action="store_true", help="skip the HBM-bound kernel rooflines (roofline_hbm)")
    ap.add_argument("--no-suffix-sharing", action="store_true",
                    help="score every candidate's full tail (the north-star dataflow) instead of computing the "
                         "rows it shares with the current suffix once")
    ap.add_argument("--cpu-sample", type=int, default=0,
                    help="candidates per CPU-baseline step (0 = 3 for the in-run cpu_baseline, 1 per step for --impl reference)")
    return ap.parse_args()


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"bf16_burst": p["bf16_tflops"], "bf16_sustained": p["bf16_tflops_sustained"],
                "hbm_gbs": p["hbm_gbs"], "source": "measured"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.proc = gpu_index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        busy = [c for c in sm if c > 0]
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def algorithmic_flops(d, d_ff, n_layers, heads, head_dim, vocab, P, S, T, n_cand, kv_heads=None):
    """linear layers 2*params*rows, attention 4*rows*(keys)*d, LM head on
    the T target rows; prefix counted once (its last layer stops after the K/V projection)."""
    q = heads * head_dim
    kv = (kv_heads or heads) * head_dim
    per_row = 2 * (d * (q + 2 * kv) + q * d + 3 * d * d_ff)
    tail = n_cand * (S * per_row * n_layers + 4 * S * (P + S / 2) * q * n_layers + 2 * T * d * vocab)
    prefix = P * per_row * n_layers + 4 * P * (P / 2) * q * n_layers
    return tail + prefix, tail / max(n_cand, 1)


def oracle_scoring_fn(model, tok, proc, norm, cfg, messages, target, image, joint, segment_ids=None,
                      target_ids=None):
    """The reference's scoring dataflow (oracle/hotpath.py, pinned to the reference by tests/golden)
    for n candidates: sw-fold repeat of the whole prompt, full sequence, full-vocab logits."""
    from oracle import hotpath
    ora = hotpath.OracleAttack(model, tok, proc, cfg, normalize=norm)
    if segment_ids is None:
        msgs = hotpath.build_messages(messages, joint)
        prompt = proc.apply_chat_template(msgs, add_generation_prompt=True)
        segment_ids = hotpath.split_prompt(prompt, tok, joint, ora.gemma_proc, "cpu")
        target_ids = tok(target, add_special_tokens=False, return_tensors="pt")["input_ids"]
        segment_ids["target"] = target_ids
    ora.target_ids = target_ids.cpu()
    ora.seg_embeds = {k: ora.emb(v.cpu()) for k, v in segment_ids.items()}
    mtype = getattr(model.config, "model_type", "")
    order = (hotpath.segment_order("pgd", mtype, single=True) if joint
             else hotpath.segment_order("gcg", mtype, no_joint_eval=True))

    def score(cand):
        n = cand.shape[0]
        with torch.no_grad():
            feats = hotpath.image_features(model, norm(image), ora.gemma_proc) if joint else None
            emb = hotpath.assemble(ora._parts(cand, feats), order, n)
            loss, _ = hotpath.candidates_loss(model, emb, ora.target_ids, n)
        return loss, emb.shape[1]
    return score


def run_reference(args):
    """--impl reference: the reference's dataflow on the host cores. Each step scores
    ``cpu_sample`` candidates the reference's way: full sequence, full-vocabulary logits, CE in
    model dtype."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from bimodalattack_b200 import BimodalAttackConfig
    # torchrun exports OMP_NUM_THREADS=1 for nproc > 1: the CPU arm must still use every host core
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    t0 = time.perf_counter()
    model, tok, proc, norm, image, messages, target, init, joint = build_case(args, torch.float32, None, fast_init=True)
    build_s = time.perf_counter() - t0
    cfg = BimodalAttackConfig(num_steps=1, search_width=args.search_width, pgd_attack=joint, gcg_attack=True,
                              joint_eval=joint, allow_non_ascii=True, optim_str_init=init, seed=1,
                              alpha=4 / 255, eps=64 / 255, verbosity="ERROR", images_folder=None)
    score = oracle_scoring_fn(model, tok, proc, norm, cfg, messages, target, image, joint)
    n = args.cpu_sample or 1
    g = torch.Generator().manual_seed(3)
    n_tok = len(tok(init, add_special_tokens=False)["input_ids"])
    vocab = model.get_input_embeddings().num_embeddings

    def step():
        sampled = torch.randint(16, min(vocab, 32000), (n, n_tok), generator=g)
        loss, seq_len = score(sampled)
        return seq_len, int(loss.argmin())

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        seq_len, _ = step()
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    cores = torch.get_num_threads()
    sample = (f"{n} candidate(s) per step through the reference dataflow (full {seq_len}-token sequence, "
              f"full-vocab logits, fp32), model {args.model}; search_width-{args.search_width} step extrapolates linearly")
    T = len(tok(target, add_special_tokens=False)["input_ids"])
    tail = n_tok + (6 + T if joint else seq_len - 20 - n_tok)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(args, seq_len - tail, tail),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "host_cpus": os.cpu_count(), "model_build_s": build_s},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


GEMMA = {"gemma4b": "4b", "gemma-tiny256": "tiny256", "gemma4b-joint": "4b", "gemma-tiny256-joint": "tiny256"}
LLAVA_PX = {"7b": 336, "small128": 56, "tiny128": 28, "tiny": 28}
LLAVA_VOCAB = {"7b": 32000, "small128": 4000, "tiny128": 1000, "tiny": 990}


def workload_config(args, P, S):
    sw = args.search_width * (args.gpus if args.scaling == "weak" else 1)
    if args.model in GEMMA:
        if args.model.endswith("-joint"):
            what = (f"Gemma-3-{GEMMA[args.model]} arch random-init, joint GCG+PGD scoring phase (Gemma order "
                    f"before_img|optim|before_suffix|image|after|target: the image tokens follow the suffix, so they "
                    f"are part of every candidate's tail), search_width {sw} ({args.scaling} scaling), topk 256, n_replace 1")
        else:
            what = (f"Gemma-3-{GEMMA[args.model]} arch random-init, GCG-only scoring phase (text-only order "
                    f"before|optim|after|target, 262k-vocab LM head fused with CE), search_width {sw} "
                    f"({args.scaling} scaling), topk 256, n_replace 1")
        l2 = "no flush needed: every step streams the 7.8 GB of decoder weights + 1.3 GB LM head (>> 126 MB L2)"
    else:
        px = LLAVA_PX[args.model]
        what = (f"LLaVA-1.5-{args.model} arch random-init, joint GCG+PGD scoring phase, "
                f"search_width {sw} ({args.scaling} scaling), {px}x{px} synthetic image, topk 256, n_replace 1")
        l2 = "no flush needed: every step streams ~23 GB of weights (>> 126 MB L2)"
    return {"workload": what, "search_width": sw, "prefix_rows": P, "tail_rows": S,
            "parallelism": f"candidate-sharded dp{args.gpus}, replicated model", "l2": l2}


def build_case(args, dtype, device, fast_init=False):
    """(model, tok, proc, norm, image_cpu, messages, target, init, joint) of the selected workload."""
    from bimodalattack_b200 import synthetic as syn
    if args.model in GEMMA:
        size = GEMMA[args.model]
        tok, proc, norm, image, messages, target, init = syn.gemma_case(262208 if size == "4b" else 1024,
                                                                        896 if size == "4b" else 56)
        model = syn.make_gemma3(size, dtype=dtype, device=device, seed=0, embed_gain=4.0 if size != "4b" else 1.5,
                                fast_init=fast_init and size == "4b")
        model.config._attn_implementation = "sdpa"
        return model, tok, proc, norm, image, messages, target, init, args.model.endswith("-joint")
    tok, proc, norm, image, messages, target, init = syn.headline_case(LLAVA_VOCAB[args.model], LLAVA_PX[args.model])
    kw = dict(fast_init=True) if fast_init and args.model == "7b" else {}
    model = syn.make_llava(args.model, dtype=dtype, device=device, seed=0, head_gain=HEAD_GAIN[args.model],
                           attn_implementation=None if device is None else "sdpa", **kw)
    return model, tok, proc, norm, image, messages, target, init, True


def hbm_bound_kernels(atk, model, sampled, optim_ids, image, original, cfg, peaks):
    """Roofline entries of the HBM-bound kernels of the path: token-gradient GEMM, sampler
    (top-k + replacement), PGD step, RMSNorm. Each is timed alone with CUDA events, L2 flushed (a 512 MB
    write) between iterations; bytes are the ALGORITHMIC bytes per launch. `traffic` comes from the
    committed ncu captures (profiles/ncu_traffic.json) when present."""
    import ctypes as C
    from bimodalattack_b200 import native
    dev = image.device if image is not None else sampled.device
    lib = native.load()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get("hbm_kernels", {})

    def timeit(fn, iters=10, warm=3, back_to_back=False):
        for _ in range(warm):
            fn()
        if back_to_back:      # working set > L2 (126 MB): consecutive launches cannot reuse it, no flush needed,
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)   # and launch latency amortises
            torch.cuda.synchronize()
            a.record()
            for _ in range(iters):
                fn()
            b.record()
            b.synchronize()
            return a.elapsed_time(b) * 1e-3 / iters
        ts = []
        for _ in range(iters):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            b.synchronize()
            ts.append(a.elapsed_time(b))
        return statistics.median(ts) * 1e-3

    def entry(name, kernel, secs, nbytes, launches=1):
        gbs = nbytes / secs / 1e9
        return {"kernel": kernel, "bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": gbs / peaks["hbm_gbs"], "us": secs * 1e6, "algorithmic_bytes": nbytes, "launches": launches,
                "traffic": traffic.get(name), "peak_source": peaks["source"] + " (copy bandwidth; kernel timed alone, L2 flushed)"}

    out = {}
    W = model.get_input_embeddings().weight.detach()
    V, d = W.shape
    n_tok = optim_ids.shape[1]
    g = (torch.randn(n_tok, d, device=dev) * 1e-3).to(torch.bfloat16)
    secs = timeit(lambda: native.token_grad(g, W), back_to_back=True)
    out["token_grad"] = entry("token_grad", "gemm_bf16_tn_kernel<*, EPI_STORE_T> (W streamed once as the A operand)", secs,
                              2 * (V * d + n_tok * d + n_tok * V))
    grad = native.token_grad(g, W)
    ids = optim_ids.reshape(-1)
    secs = timeit(lambda: native.sample_ids(ids, grad, cfg.search_width, cfg.topk, cfg.n_replace, None, 1, 0))
    two = grad.shape[1] > 65536
    out["sampler"] = entry("sampler", ("topk_chunk + topk_merge" if two else "topk_row16 (one launch)") + " + sample_replace", secs,
                           n_tok * V * 2 + n_tok * cfg.topk * 8 + cfg.search_width * n_tok * 8, launches=3 if two else 2)
    # BASELINE config 2, second variant: a random 30 % of the vocabulary not allowed (the mask is a V-bit set folded
    # into the top-k read; the scoring phase does not depend on it)
    gm = torch.Generator().manual_seed(30)
    banned = torch.randperm(V, generator=gm)[: int(0.3 * V)].to(dev)
    mask = native.pack_mask_bits(banned, V, dev)
    secs = timeit(lambda: native.sample_ids(ids, grad, cfg.search_width, cfg.topk, cfg.n_replace, mask, 1, 0))
    out["sampler_masked_30pct"] = entry("sampler", "same, 30 % of the vocabulary masked out", secs,
                                        n_tok * V * 2 + V // 8 + n_tok * cfg.topk * 8 + cfg.search_width * n_tok * 8,
                                        launches=3 if two else 2)
    if image is not None:
        x = image.clone()
        gi = torch.randn_like(x)
        secs = timeit(lambda: native.pgd_step_(x, gi, original, cfg.alpha * cfg.eps, cfg.eps))
        out["pgd_step"] = entry("pgd_step", "pgd_step_kernel (sign/step/project/clamp fused)", secs, 4 * x.numel() * 4)
    rows = 16384
    xr = torch.randn(rows, d, device=dev).to(torch.bfloat16)
    yr = torch.empty_like(xr)
    wn = torch.ones(d, dtype=torch.bfloat16, device=dev)
    st = native.stream_ptr(dev)
    secs = timeit(lambda: native.check(lib.bma_rmsnorm(xr.data_ptr(), d, wn.data_ptr(), yr.data_ptr(), d, rows, d,
                                                       C.c_float(1e-5), 0, st), "rmsnorm"), back_to_back=True)
    out["rmsnorm"] = entry("rmsnorm", f"rmsnorm_kernel ({rows} rows x {d})", secs, 2 * rows * d * 2)
    return out


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as td
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        td.init_process_group("nccl", device_id=dev)

    from bimodalattack_b200 import BimodalAttack, BimodalAttackConfig, native, synthetic as syn
    from bimodalattack_b200.attack import image_features
    from bimodalattack_b200 import dist as bdist
    lib = native.load()
    peaks = load_peaks()

    model, tok, proc, norm, image_cpu, messages, target, init, joint = build_case(args, torch.bfloat16, dev)
    gemma = args.model in GEMMA
    sw = args.search_width * (world if args.scaling == "weak" else 1)
    cfg = BimodalAttackConfig(num_steps=1, search_width=sw, topk=256, n_replace=1, pgd_attack=joint,
                              gcg_attack=True, joint_eval=joint, allow_non_ascii=True, filter_ids=False,
                              optim_str_init=init, seed=1, alpha=4 / 255, eps=64 / 255, verbosity="ERROR",
                              images_folder=None)
    atk = BimodalAttack(model, tok, proc, cfg, normalize=norm)
    atk._prepare(messages, target)
    w = atk.world
    image = image_cpu.to(dev)
    original = image.clone()

    # one real gradient (+ PGD) + sampling on rank 0 gives this step's inputs
    torch.manual_seed(1)
    optim_ids = tok(init, add_special_tokens=False, return_tensors="pt")["input_ids"].to(dev)
    n_tok = optim_ids.shape[1]
    sampled = None
    if rank == 0:
        img = image.clone().requires_grad_() if joint else None
        g_tok, g_img = atk.compute_gradient(optim_ids, img)
        if joint:
            image = atk.perform_pgd_step(img, cfg.eps, cfg.alpha, g_img, original).detach()
        sampled, _, _, _ = atk.candidate_sampling(cfg, 0, optim_ids, g_tok, tok, [], 0.0)
    sampled = bdist.broadcast_candidates(sampled, n_tok, dev, w)
    if joint:
        image = bdist.broadcast_tensor(image.contiguous(), w)
    base_ids = optim_ids.reshape(-1).to(torch.int64).contiguous()
    atk.engine.share_suffix_head = not args.no_suffix_sharing
    atk._base = (sampled, base_ids)          # the suffix the candidates were sampled from (all ranks)
    lo, hi = bdist.shard_bounds(sw, w)       # (replaced below by the row-balanced slice the scorer itself uses)

    def build_batch(img):
        if joint:
            feats = atk.scoring_features(img)      # normalise + vision tower + projector, as run() does
            return atk._build_input_embeds_pgd(sampled, feats, search_width=sw, single=True)
        return atk._build_input_embeds_gcg(sampled, search_width=sw, no_joint_eval=True)

    def scoring_step():
        with torch.no_grad():
            atk.engine.prefill_tag = None        # every step rebuilds the shared-prefix K/V (as a real step does)
            atk._compute_candidates_loss_original(sw, build_batch(image))
        return atk._winner()                     # one 8-byte D2H: the step's result

    def barrier():
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    # ---- value: device-resident inputs (profiler OFF) ---------------------------------------------
    for _ in range(args.warmup):
        scoring_step()
    P = atk.engine.prefix_len
    with torch.no_grad():
        plan = atk._plan(build_batch(image))
    S, T = plan.n_tail, atk.target_ids.numel()
    lo, hi, shard_hint = atk.shard(sampled, base_ids, plan)   # this rank's slice (row-balanced when N > 1)
    clocks = ClockSampler(local)
    atk.engine.exchange_wait_ms(reset=True)
    barrier()
    launches0 = lib.bma_launch_count()
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    torch.cuda.nvtx.range_push("bench_timed")        # lets ncu restrict itself to the timed steps
    for _ in range(args.steps):
        win_loss, win_idx = scoring_step()
    torch.cuda.nvtx.range_pop()
    e1.record()
    barrier()
    clock_info = clocks.stop()
    launches = lib.bma_launch_count() - launches0
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        td.all_reduce(ms, op=td.ReduceOp.MAX)
    ms_per_step = float(ms.item()) / args.steps
    value = sw / (ms_per_step * 1e-3)
    # per rank: time spent waiting for other ranks in the per-layer prefix exchange (the straggler shows ~0)
    exchange_wait, exchange_floor_us, early_signal_ab, adaptive_ab = None, None, None, None
    if world > 1:
        wt = torch.zeros(world, device=dev)
        wt[rank] = atk.engine.exchange_wait_ms(reset=True) / args.steps
        td.all_reduce(wt)
        exchange_wait = [round(float(x), 3) for x in wt.tolist()]
        # latency floor of one exchange: 200 data-free exchanges back to back, nothing else on the GPUs
        atk.engine.peers_barrier(20)
        barrier()
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record()
        atk.engine.peers_barrier(200)
        b1.record()
        b1.synchronize()
        exchange_floor_us = round(b0.elapsed_time(b1) / 200 * 1e3, 2)
        atk.engine.exchange_wait_ms(reset=True)

        if args.multi_gpu_ab:
            # A/B of the early signal (the QKV GEMM publishes first and raises the flags itself) in the same process
            def timed_ms(n):
                barrier()
                t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0.record()
                for _ in range(n):
                    scoring_step()
                t1.record()
                barrier()
                t = torch.tensor([t0.elapsed_time(t1)], device=dev)
                td.all_reduce(t, op=td.ReduceOp.MAX)
                return float(t.item()) / n
            ab = {}
            for name, flag in (("off", False), ("on", True), ("off2", False), ("on2", True)):
                atk.engine.set_early_signal(flag)
                scoring_step()
                ms_ab = timed_ms(args.steps)
                wt = torch.zeros(world, device=dev)
                wt[rank] = atk.engine.exchange_wait_ms(reset=True) / (args.steps + 1)
                td.all_reduce(wt)
                ab[name] = {"candidates_per_s": round(sw / (ms_ab * 1e-3), 1), "ms_per_step": round(ms_ab, 3),
                            "exchange_wait_ms_per_step": [round(float(x), 3) for x in wt.tolist()]}
            early_signal_ab = ab
            atk.engine.set_early_signal(False)

            # straggler-aware shares (BimodalAttack.rebalance): 10 untimed steps move the candidate shares towards the
            # ranks' measured speeds, then the same timed loop
            atk.adapt_shares = True
            atk._rebalance_calls = 2
            w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            atk.engine.exchange_wait_ms(reset=True)
            for _ in range(10):
                w0.record()
                scoring_step()
                w1.record()
                w1.synchronize()
                atk.rebalance(w0.elapsed_time(w1))
            atk.adapt_shares = False             # frozen cuts for the timed steps (and everything after)
            ms_ad = timed_ms(args.steps)
            wt = torch.zeros(world, device=dev)
            wt[rank] = atk.engine.exchange_wait_ms(reset=True) / args.steps
            td.all_reduce(wt)
            adaptive_ab = {"candidates_per_s": round(sw / (ms_ad * 1e-3), 1), "ms_per_step": round(ms_ad, 3),
                           "rank_share": [round(x, 4) for x in (atk.rank_share or [])],
                           "exchange_wait_ms_per_step": [round(float(x), 3) for x in wt.tolist()]}
            atk.rank_share = None
    rows_run = atk.engine.last_rows          # tail (+ fused prefix) rows of this rank's shard that went through the decoder

    # ---- per-class profile: a SEPARATE pass with an event pair around every launch -----------------
    prof_steps = max(1, min(args.steps, 5))
    atk.engine.profile(True)
    for _ in range(prof_steps):
        scoring_step()
    torch.cuda.synchronize()
    prof = atk.engine.profile_read()
    atk.engine.profile(False)

    # ---- e2e: host buffers through the C ABI ---------------------------------------------------
    ids_host = sampled.cpu().pin_memory()
    base_host = base_ids.cpu().pin_memory()
    img_host = image.cpu().pin_memory()
    loss_host = torch.empty(sw, dtype=torch.float32).pin_memory()
    key_host = torch.empty(1, dtype=torch.int64).pin_memory()
    img_dev = torch.empty_like(image)
    key_dev = torch.empty(1, dtype=torch.int64, device=dev)

    def e2e_step():
        with torch.no_grad():
            if joint:
                img_dev.copy_(img_host, non_blocking=True)
            pl = atk._plan(build_batch(img_dev))
            atk.engine.prefill(pl.prefix)
            # multi-rank: the library reduces the winner key (one NCCL all-reduce) before it leaves the device
            atk.engine.score_host(ids_host[lo:hi], pl, atk.target_ids, loss_host[lo:hi], key_host, idx_base=lo,
                                  base_ids_host=base_host, batch_hint=shard_hint)
        return int(key_host.item())

    for _ in range(max(1, args.warmup // 2)):
        e2e_step()
    barrier()
    e0.record()
    for _ in range(args.steps):
        e2e_key = e2e_step()
    e1.record()
    barrier()
    ms2 = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        td.all_reduce(ms2, op=td.ReduceOp.MAX)
    e2e_value = sw / (float(ms2.item()) / args.steps * 1e-3)
    h2d = ids_host[lo:hi].numel() * 8 + base_host.numel() * 8 + (img_host.numel() * 4 if joint else 0)
    d2h = (hi - lo) * 4 + 8

    # ---- full attack steps through run() ---------------------------------------------------------
    full = None
    if args.full_steps > 0:
        import dataclasses
        warm_full = 4          # eager warm-up calls and the one-time CUDA-graph captures of phase A / vision tower
        cfg2 = dataclasses.replace(cfg, num_steps=args.full_steps + warm_full)
        atk2 = BimodalAttack(model, tok, proc, cfg2, normalize=norm, engine=atk.engine)
        marks = []

        def hook(i):
            torch.cuda.synchronize()
            marks.append(time.perf_counter())
        atk2.step_hook = hook
        res = atk2.run(messages, "goal", target, image_cpu.to(dev) if joint else None)
        if len(marks) > warm_full:
            dt = torch.tensor([(marks[-1] - marks[warm_full - 1]) / (len(marks) - warm_full)], device=dev)
            if world > 1:
                td.all_reduce(dt, op=td.ReduceOp.MAX)
            full = {"steps_per_sec": 1.0 / float(dt.item()), "ms_per_step": 1e3 * float(dt.item()),
                    "timed_steps": len(marks) - warm_full, "losses": res.losses,
                    "phase_s_last": {"gradient": res.gradient_times[-1] if res.gradient_times else None,
                                     "pgd": res.pgd_times[-1] if res.pgd_times else None,
                                     "sampling": res.sampling_times[-1] if res.sampling_times else None,
                                     "scoring": res.loss_times[-1]}}

    # ---- multi-rank parity: rank 0 scores ALL candidates alone; the sharded winner must be that winner ----
    unsharded = None
    if world > 1:
        barrier()
        if rank == 0:
            with torch.no_grad():
                eng = atk.engine
                eng.attach_comm(None)
                pl = atk._plan(build_batch(image))
                eng.prefill(pl.prefix, deferred=False)
                all_loss = eng.score(sampled, pl, atk.target_ids, base_ids=base_ids)
                u_loss, u_idx = atk._winner()        # (the communicator stays detached: this is the last scoring call)
            # tie-aware: bf16 activation noise (row offsets move the attention's key-block boundaries) is
            # 2 bf16 ulps of the loss; the sharded winner must be within that of the unsharded minimum
            import math
            ulp = 2.0 ** (math.floor(math.log2(max(abs(u_loss), 1e-3))) - 7)
            w_unsh = float(all_loss[win_idx])
            unsharded = {"loss": u_loss, "index": u_idx, "same_index": bool(u_idx == win_idx),
                         "sharded_winner_loss_unsharded": w_unsh, "tolerance_2ulp": 2 * ulp,
                         "matches_sharded": bool(u_idx == win_idx or w_unsh <= u_loss + 2 * ulp)}
        barrier()

    if rank != 0:
        if world > 1:
            td.destroy_process_group()
        return

    hbm = hbm_bound_kernels(atk, model, sampled, optim_ids, image, original, cfg, peaks) if not args.no_hbm_kernels else None

    # ---- roofline of the dominant kernel class (tcgen05 GEMMs), from in-run CUDA events ----------
    tc = model.config.text_config
    gemm_classes = [k for k in prof if k.startswith("gemm")]
    g_ms = sum(prof[k]["ms"] for k in gemm_classes)
    g_fl = sum(prof[k]["flops"] for k in gemm_classes)
    total_prof_ms = sum(v["ms"] for v in prof.values())
    achieved = g_fl / (g_ms * 1e-3) / 1e12 if g_ms > 0 else 0.0
    total_flops, per_cand = algorithmic_flops(tc.hidden_size, tc.intermediate_size, tc.num_hidden_layers,
                                              tc.num_attention_heads, atk.engine.head_dim,
                                              model.get_output_embeddings().weight.shape[0], P, S, T, hi - lo,
                                              kv_heads=atk.engine.n_kv)
    # the dominant kernel: the gate/up projection with the SwiGLU (GeGLU) epilogue
    top = prof["gemm_gate_up_glu"]
    top_tf = top["flops"] / (top["ms"] * 1e-3) / 1e12 if top["ms"] > 0 else 0.0
    rows_top = top["flops"] / max(top["launches"], 1) / (2.0 * 2 * tc.intermediate_size * tc.hidden_size)
    alg_bytes = 2.0 * (2 * tc.intermediate_size * tc.hidden_size + rows_top * tc.hidden_size
                       + rows_top * tc.intermediate_size)        # W + A + C, bf16, per launch
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(args.model, {}).get("gemm_gate_up_glu")
    roofline = {"bound": "tensor",
                "kernel": f"gemm_bf16_tn_kernel<256, EPI_GLU, 128, {1 if os.environ.get('BMA_GEMM_CTAS') == '1' else 2}> "
                          "(gate/up projection + GLU epilogue, tcgen05)",
                "achieved": top_tf, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                "frac": top_tf / peaks["bf16_sustained"], "frac_of_burst_peak": top_tf / peaks["bf16_burst"],
                "peak_source": peaks["source"] + " (sustained cuBLAS bf16; kernel timed inside a long step)",
                "launches_per_step": top["launches"] / prof_steps,
                "avg_launch_ms": top["ms"] / max(top["launches"], 1),
                "algorithmic_flops_per_launch": top["flops"] / max(top["launches"], 1),
                "algorithmic_bytes_per_launch": alg_bytes,
                "all_tcgen05_gemms": {"achieved": achieved, "frac": achieved / peaks["bf16_sustained"],
                                      "frac_of_burst_peak": achieved / peaks["bf16_burst"]},
                # dram__bytes_read.sum + dram__bytes_write.sum of one launch of this kernel from the committed
                # `ncu --set full` capture (profiles/ncu_traffic.json); not measured in this run
                "traffic": traffic, "share_of_step": top["ms"] / total_prof_ms if total_prof_ms else None,
                "share_of_scoring_kernels": g_ms / total_prof_ms if total_prof_ms else None,
                "per_class": {k: {"ms_per_step": v["ms"] / prof_steps, "launches_per_step": v["launches"] / prof_steps,
                                  "tflops": (v["flops"] / (v["ms"] * 1e-3) / 1e12) if v["ms"] > 0 and v["flops"] > 0 else None}
                              for k, v in prof.items()},
                # SURVEY 8(d) numerator (every candidate's full tail) over the step time; with suffix-head
                # sharing fewer rows are executed, so this is an EFFECTIVE rate, not a tensor-pipe rate:
                "step_tflops_algorithmic": total_flops / (ms_per_step * 1e-3) / 1e12,
                "gflop_per_candidate": per_cand / 1e9,
                "suffix_head_sharing": {"enabled": bool(atk.engine.share_suffix_head), "tail_rows_full": (hi - lo) * S,
                                        "tail_rows_executed": int(rows_run),
                                        "note": "rows a candidate shares with the current suffix (before its first "
                                                "replaced position) are computed once; SURVEY 8(f) rank 3"},
                "gemm_tflops_executed_per_step": g_fl / prof_steps / 1e12,
                "profile_pass": f"{prof_steps} extra steps with an event pair around every launch (value and e2e are timed without)"}

    # ---- CPU baseline: the oracle (reference dataflow) on the host cores, bounded sample ------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        import copy
        t0 = time.perf_counter()
        twin = copy.deepcopy(model).to("cpu").float()
        score = oracle_scoring_fn(twin, tok, proc, norm, cfg, messages, target, image.cpu(), joint,
                                  segment_ids=atk.segment_ids, target_ids=atk.target_ids)
        n = args.cpu_sample or 3
        t1 = time.perf_counter()
        cl, seq_len = score(sampled[:n].cpu())
        dt = time.perf_counter() - t1
        cpu = {"value": n / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
               "host_cpus": os.cpu_count(),
               "sample": f"{n} of {sw} candidates through the reference dataflow (full {seq_len}-token "
                         f"sequence, full-vocab logits, fp32) in {dt:.1f}s; linear extrapolation to a full step",
               "loss_vs_gpu_abs_diff": float((cl - loss_host[:n]).abs().max()) if lo == 0 else None,
               "setup_s": t1 - t0}

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": workload_config(args, P, S),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches), "clocks": clock_info, "roofline": roofline, "cpu_baseline": cpu,
        "roofline_hbm": hbm,
        "exchange_wait_ms_per_step": exchange_wait, "exchange_floor_us": exchange_floor_us,
        "early_signal_ab": early_signal_ab, "adaptive_shares_ab": adaptive_ab, "full_step": full, "winner": {"loss": win_loss, "index": win_idx, "e2e_key": e2e_key, "unsharded_rank0": unsharded},
        "a100_reference_published": (
            {"cand_per_s": 12.1, "s_per_iter": 43.9353,
             "note": "BASELINE.md: Gemma-3-4b-it joint_eval sw 512, A100-80GB, bf16"} if gemma and joint else
            {"cand_per_s": 128.0, "s_per_iter": 5.7641,
             "note": "BASELINE.md: Gemma-3-4b-it GCG-only sw 512, A100-80GB, bf16"} if gemma else
            {"cand_per_s": 10.7, "s_per_iter": 49.965,
             "note": "BASELINE.md: LLaVA-1.5-7B joint_eval sw 512, A100-80GB, fp16"}),
    }
    print(json.dumps(out))
    if world > 1:
        td.destroy_process_group()


if __name__ == "__main__":
    main()
