"""ORACLE -- TEST INFRASTRUCTURE ONLY. Never imported by the product package.

Plain-torch restatement of the reference's per-step hot path (Bimo99B9/BimodalAttack,
``bimodalattack/bimodal_attack.py``; every function cites the lines it follows). It keeps the
reference's *dataflow* on purpose -- one-hot @ W for the suffix gradient, sw-fold repeat of the
whole prompt, full-sequence forward, full-vocabulary logits, loss in model dtype -- because it is
the checker for the B200 path, not an implementation of it. Device-agnostic: run on CPU it is the
CPU baseline, run on CUDA tensors it reproduces the reference's CUDA RNG stream bit for bit
(torch.rand / torch.randint / torch.topk are the very calls the reference makes).

Parity pin: the reference has no golden vectors or tests of its own, so this
restatement is pinned against outputs of the reference itself, imported unmodified from
/root/reference in the build container by ``oracle/make_golden.py``; the resulting fixtures live
in ``tests/golden/`` and ``tests/test_oracle_golden.py`` replays them without the reference.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import this module.
"""
from __future__ import annotations

import copy
import time
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F


# initial characters for the extra buffer entries (reference bimodalattack/utils.py:8-12)
INIT_CHARS = [".", ",", "!", "?", ";", ":", "(", ")", "[", "]", "{", "}", "@", "#", "$", "%", "&", "*", "w", "x", "y", "z"]


# ------------------------------------------------------------------------------------------------
# (4) PGD step -- bimodal_attack.py:1030-1037
def pgd_step(image: torch.Tensor, eps: float, alpha: float, image_grad: torch.Tensor,
             image_original: torch.Tensor) -> torch.Tensor:
    moved = image - alpha * eps * torch.sign(image_grad)          # step size is alpha*eps (:1033)
    moved = moved.detach().requires_grad_()
    boxed = torch.clamp(moved, image_original - eps, image_original + eps)   # :1034
    return torch.clamp(boxed, 0, 1)                                          # :1035


# ------------------------------------------------------------------------------------------------
# (2) candidate sampling -- bimodal_attack.py:130-163
def sample_ids_from_grad(ids: torch.Tensor, grad: torch.Tensor, search_width: int, topk: int = 256,
                         n_replace: int = 1, not_allowed_ids: Optional[torch.Tensor] = None):
    """ids (L,), grad (L, V). NOTE: like the reference (:145) this writes +inf into ``grad``."""
    n_tok = len(ids)
    tiled = ids.repeat(search_width, 1)                                        # :142
    if not_allowed_ids is not None:
        grad[:, not_allowed_ids.to(grad.device)] = float("inf")               # :144-145
    best_tokens = (-grad).topk(topk, dim=1).indices                           # :147
    # first RNG draw: sw*L uniforms -> per-row permutation -> first n_replace positions (:150-154)
    order = torch.argsort(torch.rand((search_width, n_tok), device=grad.device))
    positions = order[..., :n_replace]
    # second RNG draw: sw*n_replace ints in [0, topk) (:156-160)
    picks = torch.randint(0, topk, (search_width, n_replace, 1), device=grad.device)
    values = torch.gather(best_tokens[positions], 2, picks).squeeze(2)
    return tiled.scatter_(1, positions, values)                               # :162


# bimodal_attack.py:166-186
def filter_ids(ids: torch.Tensor, tokenizer) -> torch.Tensor:
    texts = tokenizer.batch_decode(ids)
    kept = []
    for row, text in zip(ids, texts):
        again = tokenizer(text, return_tensors="pt", add_special_tokens=False).to(ids.device)["input_ids"][0]
        if torch.equal(row, again):
            kept.append(row)
    if not kept:
        raise RuntimeError(
            "No token sequences are the same after decoding and re-encoding. "
            "Consider setting filter_ids=False or trying a different optim_str_init")
    return torch.stack(kept)


# bimodalattack/utils.py:14-33
def nonascii_token_ids(tokenizer, device="cpu") -> torch.Tensor:
    bad = [i for i in range(tokenizer.vocab_size)
           if not (tokenizer.decode([i]).isascii() and tokenizer.decode([i]).isprintable())]
    for special in (tokenizer.bos_token_id, tokenizer.eos_token_id, tokenizer.pad_token_id,
                    tokenizer.unk_token_id):
        if special is not None:
            bad.append(special)
    return torch.tensor(bad, device=device)


# ------------------------------------------------------------------------------------------------
# segment orders -- bimodal_attack.py:1150-1215
def segment_order(mode: str, model_type: str, single: bool = False, no_joint_eval: bool = False,
                  no_target: bool = False) -> List[str]:
    gemma = model_type == "gemma3"
    with_image = (["before_img", "optim", "before_suffix", "image", "after"] if gemma
                  else ["before_img", "image", "before_suffix", "optim", "after"])
    if mode == "pgd":
        assert single, "PGD mode only supports single=True"                    # :1151
        return with_image + ["target"]
    if mode == "gcg":
        if single:
            return (["before_img", "optim", "before_suffix", "after", "target"] if gemma
                    else ["before_img", "before_suffix", "optim", "after", "target"])
        if no_joint_eval:
            return ["before", "optim", "after", "target"]
        if no_target:
            return ["before", "optim", "after"]
        raise ValueError("Invalid flags for BimodalAttack mode")
    if mode == "gcg_pgd":
        if not single and no_target:
            return with_image
        return with_image + ["target"]
    raise ValueError(f"Unknown mode '{mode}'")


# bimodal_attack.py:1217-1225
def assemble(parts: Dict[str, torch.Tensor], order: Sequence[str], search_width: Optional[int]):
    chunks = []
    for name in order:
        t = parts[name]
        if search_width is not None and t.shape[0] == 1:
            t = t.repeat(search_width, 1, 1)
        chunks.append(t)
    return torch.cat(chunks, dim=1)


# ------------------------------------------------------------------------------------------------
# (3) candidate loss -- bimodal_attack.py:1278-1310
def candidates_loss(model, input_embeds: torch.Tensor, target_ids: torch.Tensor, chunk: int,
                    early_stop: bool = False) -> Tuple[torch.Tensor, bool]:
    losses, matched = [], False
    n_target = target_ids.shape[1]
    for lo in range(0, input_embeds.shape[0], chunk):
        with torch.no_grad():
            batch = input_embeds[lo:lo + chunk]
            logits = model(inputs_embeds=batch).logits
            cut = input_embeds.shape[1] - n_target
            pred = logits[..., cut - 1:-1, :].contiguous()                     # :1289-1290
            labels = target_ids.repeat(batch.shape[0], 1)
            tok_loss = F.cross_entropy(pred.view(-1, pred.size(-1)), labels.view(-1), reduction="none")
            losses.append(tok_loss.view(batch.shape[0], -1).mean(dim=-1))     # :1298
            if early_stop and torch.any(torch.all(torch.argmax(pred, dim=-1) == labels, dim=-1)).item():
                matched = True                                                 # :1300-1306
    return torch.cat(losses, dim=0), matched


def image_features(model, pixel_values: torch.Tensor, gemma: bool) -> torch.Tensor:
    """model.get_image_features as the reference calls it (:528-536), returning the 4.50-style
    tensor on any transformers version (5.x wraps it in BaseModelOutputWithPooling)."""
    if gemma:
        out = model.get_image_features(pixel_values=pixel_values)
    else:
        out = model.get_image_features(pixel_values=pixel_values, vision_feature_layer=-2,
                                       vision_feature_select_strategy="default")
    if isinstance(out, torch.Tensor):
        return out
    pooled = out.pooler_output
    return torch.stack(list(pooled), 0) if isinstance(pooled, (list, tuple)) else pooled


# ------------------------------------------------------------------------------------------------
@dataclass
class Prepared:
    """Token-id segments and their embeddings (bimodal_attack.py:313-393)."""
    ids: Dict[str, torch.Tensor]
    embeds: Dict[str, torch.Tensor]
    target_ids: torch.Tensor
    gemma_processor: bool
    model_type: str


@dataclass
class StepLog:
    losses: List[float] = field(default_factory=list)
    suffix_ids: List[List[int]] = field(default_factory=list)
    winner_index: List[int] = field(default_factory=list)
    candidate_losses: List[torch.Tensor] = field(default_factory=list)
    sampled: List[torch.Tensor] = field(default_factory=list)
    images: List[torch.Tensor] = field(default_factory=list)
    grad_s: List[float] = field(default_factory=list)
    sample_s: List[float] = field(default_factory=list)
    pgd_s: List[float] = field(default_factory=list)
    loss_s: List[float] = field(default_factory=list)


def split_prompt(prompt: str, tokenizer, pgd_attack: bool, gemma_processor: bool, device) -> Dict[str, torch.Tensor]:
    """bimodal_attack.py:309-371 (BOS strip, split around the image token and {optim_str})."""
    if tokenizer.bos_token and prompt.startswith(tokenizer.bos_token):
        prompt = prompt.replace(tokenizer.bos_token, "")

    def enc(text, special=True):
        kw = {"padding": False} if special else {"add_special_tokens": False}
        return tokenizer(text, return_tensors="pt", **kw)["input_ids"].to(device, torch.int64)

    if not pgd_attack:
        before, after = prompt.split("{optim_str}")
        return {"before": enc(before), "after": enc(after, special=False)}
    if gemma_processor:
        before, rest = prompt.split("{optim_str}", 1)
        before_img = before.strip()
        if "<start_of_image>" not in rest:
            raise ValueError("Expected <start_of_image> token in Gemma PGD prompt.")
        head, sep, tail = rest.partition("<start_of_image>")
        before_suffix, after = (head + sep).strip(), tail.strip()
    else:
        if "<start_of_image>" in prompt:
            before_img, rest = prompt.split("<start_of_image>", 1)
        elif "<image>" in prompt:
            before_img, rest = prompt.split("<image>", 1)
        else:
            raise ValueError("No image token found in prompt for PGD attack")
        before_suffix, after = rest.split("{optim_str}", 1)
    return {"before_img": enc(before_img), "before_suffix": enc(before_suffix),
            "after": enc(after, special=False)}


def build_messages(messages, pgd_attack: bool):
    """bimodal_attack.py:271-302."""
    msgs = [{"role": "user", "content": messages}] if isinstance(messages, str) else copy.deepcopy(messages)
    last = msgs[-1]
    if isinstance(last["content"], str) and "{optim_str}" not in last["content"]:
        last["content"] = last["content"] + " {optim_str}"
    if pgd_attack:
        if isinstance(last["content"], str):
            last["content"] = [{"type": "text", "text": last["content"]}, {"type": "image"}]
        elif isinstance(last["content"], list) and not any(c.get("type") == "image" for c in last["content"]):
            last["content"].append({"type": "image"})
    return msgs


class OracleAttack:
    """Step loop of the reference (bimodal_attack.py:251-824) for all mode combinations."""

    def __init__(self, model, tokenizer, processor, config, normalize=None):
        self.model, self.tok, self.proc, self.cfg, self.normalize = model, tokenizer, processor, config, normalize
        self.emb = model.get_input_embeddings()
        self.not_allowed = None if config.allow_non_ascii else nonascii_token_ids(tokenizer, model.device)
        self.gemma_proc = processor.__class__.__name__ == "Gemma3Processor"
        self.model_type = getattr(model.config, "model_type", "")
        self.stop = False
        if not getattr(tokenizer, "chat_template", None):                     # :233-249
            tpl = ("USER: <image>\n{{ messages[0]['content'][0]['text'] }} \nASSISTANT: " if config.pgd_attack
                   else "{% for message in messages %}{{ message['content'] }}{% endfor %}")
            tokenizer.chat_template = tpl
            processor.chat_template = tpl

    # -- helpers ------------------------------------------------------------------------------
    def _features(self, image):
        return image_features(self.model, self.normalize(image), self.gemma_proc)

    def _parts(self, sampled_ids, feats=None):
        parts = dict(self.seg_embeds)
        parts["optim"] = self.emb(sampled_ids)                                  # :1142
        if feats is not None:
            parts["image"] = feats
        return parts

    def _loss(self, embeds, chunk):
        loss, matched = candidates_loss(self.model, embeds, self.target_ids, chunk, self.cfg.early_stop)
        self.stop = self.stop or matched
        return loss

    # bimodal_attack.py:953-1028
    def gradient(self, optim_ids, image=None):
        cfg, model = self.cfg, self.model
        onehot = F.one_hot(optim_ids, num_classes=self.emb.num_embeddings).to(model.device, model.dtype)
        if cfg.gcg_attack:
            onehot.requires_grad_()
        optim_embeds = onehot @ self.emb.weight                                 # :968 (no Gemma embed scale)
        e = self.seg_embeds
        if cfg.pgd_attack:
            feats = self._features(image)
            seq = [e["before_img"], feats, e["before_suffix"], optim_embeds, e["after"], e["target"]]   # :981-991
        else:
            seq = [e["before"], optim_embeds, e["after"], e["target"]]
        x = torch.cat(seq, dim=1)
        logits = model(inputs_embeds=x).logits
        shift = x.shape[1] - self.target_ids.shape[1]
        pred = logits[..., shift - 1:-1, :].contiguous()
        loss = F.cross_entropy(pred.view(-1, pred.size(-1)), self.target_ids.view(-1))
        if cfg.pgd_attack and cfg.gcg_attack:
            g_tok, g_img = torch.autograd.grad(loss, (onehot, image))
            return g_tok, g_img
        if cfg.pgd_attack:
            return None, torch.autograd.grad(loss, image)[0]
        if cfg.gcg_attack:
            return torch.autograd.grad(loss, onehot)[0], None
        return None, None

    # bimodal_attack.py:908-951
    def sample(self, step, optim_ids, grad):
        cfg = self.cfg
        width = cfg.search_width
        if cfg.dynamic_search:
            width = max(cfg.min_search_width, int(cfg.search_width * (1 - step / cfg.num_steps)))
        if not cfg.gcg_attack:
            return optim_ids, 1
        ids = sample_ids_from_grad(optim_ids.squeeze(0), grad.squeeze(0), width, cfg.topk, cfg.n_replace,
                                   not_allowed_ids=self.not_allowed)
        if cfg.filter_ids:
            ids = filter_ids(ids, self.tok)
        return ids, ids.shape[0]

    # bimodal_attack.py:251-824
    def run(self, messages, target: str, image: Optional[torch.Tensor] = None, record: bool = True) -> StepLog:
        cfg, model, tok = self.cfg, self.model, self.tok
        if cfg.seed is not None:
            from transformers import set_seed
            set_seed(cfg.seed)
            torch.use_deterministic_algorithms(True, warn_only=True)
        msgs = build_messages(messages, cfg.pgd_attack)
        prompt = self.proc.apply_chat_template(msgs, add_generation_prompt=True)
        seg_ids = split_prompt(prompt, tok, cfg.pgd_attack, self.gemma_proc, model.device)
        self.target_ids = tok(target, add_special_tokens=False, return_tensors="pt")["input_ids"].to(model.device, torch.int64)
        seg_ids["target"] = self.target_ids
        self.seg_ids = seg_ids
        self.seg_embeds = {k: self.emb(v) for k, v in seg_ids.items()}

        # init buffer (:826-906): the initial suffix plus buffer_size - 1 random rows of INIT_CHARS tokens (one
        # torch.randint on the default generator -- it moves the RNG stream of every later draw), all scored in the
        # gcg_pgd / gcg order; the run starts from the entry with the lowest loss (:396-397). Inside the loop the
        # buffer only collects (loss, suffix) pairs for logging (:618-619, :643-644), it never feeds back.
        optim_ids = tok(cfg.optim_str_init, add_special_tokens=False, return_tensors="pt")["input_ids"].to(model.device)
        if isinstance(cfg.optim_str_init, str) and cfg.buffer_size > 1:
            pool = tok(INIT_CHARS, add_special_tokens=False, return_tensors="pt")["input_ids"].squeeze().to(model.device)
            pick = torch.randint(0, pool.shape[0], (cfg.buffer_size - 1, optim_ids.shape[1]))
            optim_ids = torch.cat([optim_ids, pool[pick]], dim=0)
        n_init = max(1, cfg.buffer_size)
        if cfg.pgd_attack:
            feats = self._features(image)
            order = segment_order("gcg_pgd", self.model_type, single=True)
            init = assemble(self._parts(optim_ids, feats), order, n_init)
        else:
            init = assemble(self._parts(optim_ids), segment_order("gcg", self.model_type, no_joint_eval=True), n_init)
        init_losses = self._loss(init, n_init)
        entries: list = []                                                      # AttackBuffer.add (:96-106)
        for i in range(n_init):
            item = (init_losses[i], optim_ids[[i]])
            if cfg.buffer_size == 0:
                entries = [item]
                continue
            if len(entries) < cfg.buffer_size:
                entries.append(item)
            else:
                entries[-1] = item
            entries.sort(key=lambda x: x[0])
        self.init_loss = entries[0][0]
        self.init_buffer = entries
        optim_ids = entries[0][1]

        log = StepLog()
        if cfg.pgd_attack:
            image.requires_grad = True                                          # :425
            original = image.clone()
        for step in range(cfg.num_steps):
            t0 = time.perf_counter()
            g_tok, g_img = self.gradient(optim_ids, image if cfg.pgd_attack else None)
            log.grad_s.append(time.perf_counter() - t0)
            pgd_t = 0.0
            if cfg.pgd_attack and not cfg.pgd_after_gcg:                        # phase B (:462-478)
                t0 = time.perf_counter()
                image = pgd_step(image, cfg.eps, cfg.alpha, g_img, original)
                pgd_t = time.perf_counter() - t0
                if cfg.gcg_attack and not cfg.joint_eval:                       # phase C (:481-491)
                    g_tok, g_img = self.gradient(optim_ids, image)
            t0 = time.perf_counter()
            sampled, width = self.sample(step, optim_ids, g_tok)                # phase D (:504-514)
            log.sample_s.append(time.perf_counter() - t0)
            t0 = time.perf_counter()
            cand_loss = None
            with torch.no_grad():
                chunk = width if cfg.batch_size is None else cfg.batch_size
                if cfg.pgd_attack:
                    feats = self._features(image)                               # post-PGD image (:526-536)
                    best = 0
                    if cfg.joint_eval:
                        mode = "gcg_pgd" if cfg.pgd_after_gcg else "pgd"
                        emb = assemble(self._parts(sampled, feats), segment_order(mode, self.model_type, single=True), width)
                        cand_loss = self._loss(emb, chunk)
                        best = cand_loss.argmin()
                    elif cfg.gcg_attack:
                        emb = assemble(self._parts(sampled), segment_order("gcg", self.model_type, single=True), width)
                        cand_loss = self._loss(emb, chunk)
                        best = cand_loss.argmin()
                    chosen = sampled[best].unsqueeze(0)
                    if not cfg.pgd_after_gcg:                                   # re-score winner with image (:605-612)
                        full = assemble(self._parts(chosen, feats), segment_order("gcg_pgd", self.model_type), None)
                        current = self._loss(full, 1).item()
                        optim_ids = chosen
                else:                                                           # GCG only (:629-652)
                    emb = assemble(self._parts(sampled), segment_order("gcg", self.model_type, no_joint_eval=True), width)
                    cand_loss = self._loss(emb, chunk)
                    current = cand_loss.min().item()
                    best = cand_loss.argmin()
                    optim_ids = sampled[best].unsqueeze(0)
            log.loss_s.append(time.perf_counter() - t0)
            if cfg.pgd_after_gcg and cfg.pgd_attack:                            # phases E, F (:668-740)
                _, g_img = self.gradient(optim_ids, image)                      # still the OLD suffix (:673)
                t0 = time.perf_counter()
                image = pgd_step(image, cfg.eps, cfg.alpha, g_img, original)
                pgd_t = time.perf_counter() - t0
                with torch.no_grad():
                    feats = self._features(image)
                    full = assemble(self._parts(chosen, feats), segment_order("gcg_pgd", self.model_type), None)
                    current = self._loss(full, 1).item()
                    optim_ids = chosen
            log.pgd_s.append(pgd_t)
            log.losses.append(current)
            log.suffix_ids.append(optim_ids.squeeze(0).tolist())
            if record:
                log.winner_index.append(int(best))
                log.candidate_losses.append(None if cand_loss is None else cand_loss.detach().float().cpu())
                log.sampled.append(sampled.detach().cpu())
                if cfg.pgd_attack:
                    log.images.append(image.detach().cpu().clone())
            if self.stop:
                break
        self.final_image = image
        return log
