"""B200-native step loop behind the reference's API (``run`` / ``BimodalAttack``).

One optimisation step (reference bimodal_attack.py:441-794):

  A  gradient        HF autograd at batch 1 with a *leaf* suffix embedding; the token gradient
                     G_embeds . W_embed^T is the hand-written tcgen05 GEMM (``native.token_grad``)
                     instead of the backward of a one-hot matmul (:959-968, :1016-1025)
  B  PGD             one fused kernel (``native.pgd_step_``)                       (:1030-1037)
  D  sampling        device top-k + Philox replacement, torch-RNG exact            (:130-163)
     scoring         shared-prefix KV + tail-only decoder pass + fused LM-head CE + argmin
                     (``ScoringEngine``), sharded over ranks with one min all-reduce
                     (:1112-1310, :550-551)

The method seams of the reference class (compute_gradient, perform_pgd_step, candidate_sampling,
_build_input_embeds*, _compute_candidates_loss_original) are kept so A/B tests against the oracle
read like the reference. Everything numeric that this package owns runs in ``libbimodal_b200.so``;
the model's own forward/backward for phase A stays on PyTorch.
"""
from __future__ import annotations

import copy
import logging
import os
import time
import types
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple, Union

import torch
import torch.nn.functional as F

from . import dist as bdist
from . import native
from .config import BimodalAttackConfig, BimodalAttackResult
from .engine import ScoringEngine, unpack_key
from .segments import TailPlan, plan_tail, segment_order
from .utils import INIT_CHARS, AttackBuffer, get_nonascii_toks

logger = logging.getLogger("gcg")          # same logger name and format as the reference (:26-36)
if not logger.hasHandlers():
    _h = logging.StreamHandler()
    _h.setFormatter(logging.Formatter("%(asctime)s [%(filename)s:%(lineno)d] %(message)s",
                                      datefmt="%Y-%m-%d %H:%M:%S"))
    logger.addHandler(_h)
    logger.setLevel(logging.INFO)
logger.propagate = False


def image_features(model, pixel_values: torch.Tensor, gemma_processor: bool) -> torch.Tensor:
    """``model.get_image_features`` as the reference calls it (:528-536), normalised to a tensor
    (transformers 5.x returns BaseModelOutputWithPooling, 4.50 a tensor)."""
    if gemma_processor:
        out = model.get_image_features(pixel_values=pixel_values)
    else:
        out = model.get_image_features(pixel_values=pixel_values, vision_feature_layer=-2,
                                       vision_feature_select_strategy="default")
    if isinstance(out, torch.Tensor):
        return out
    pooled = out.pooler_output
    return torch.stack(list(pooled), 0) if isinstance(pooled, (list, tuple)) else pooled


def find_vision_parts(model):
    """(vision tower, multimodal projector) of a LLaVA-style HF object (transformers 4.50 and 5.x layouts),
    or None."""
    inner = getattr(model, "model", model)
    tower = getattr(inner, "vision_tower", None) or getattr(model, "vision_tower", None)
    proj = getattr(inner, "multi_modal_projector", None) or getattr(model, "multi_modal_projector", None)
    if tower is None or proj is None or not hasattr(proj, "linear_1"):
        return None
    return tower, proj


def _patch_conv_forward(conv: torch.nn.Conv2d, x: torch.Tensor) -> torch.Tensor:
    """forward() for a patch-embedding Conv2d whose stride equals its kernel (CLIP / SigLIP): the
    non-overlapping patches are unfolded by pure reshapes and multiplied by the [out, C*p*p] view
    of the same weight. Same arithmetic as the convolution, but its input gradient -- which the PGD
    step needs -- is a plain GEMM plus the inverse reshapes: deterministic and fast, whereas cuDNN's
    deterministic dgrad for this shape takes 64 ms on a B200 (reference sets
    torch.use_deterministic_algorithms, :267-269). Bound to the module as a method, so copies of the
    model (``copy.deepcopy``, ``.to()``) keep working on their own weights."""
    ph, pw = conv.kernel_size
    B, C, H, W = x.shape
    gh, gw = H // ph, W // pw
    x = x[..., :gh * ph, :gw * pw]
    patches = x.reshape(B, C, gh, ph, gw, pw).permute(0, 2, 4, 1, 3, 5).reshape(B, gh * gw, C * ph * pw)
    out = F.linear(patches, conv.weight.reshape(conv.out_channels, -1), conv.bias)
    return out.transpose(1, 2).reshape(B, conv.out_channels, gh, gw)


def install_patch_embed_gemm(model) -> int:
    """Swap the forward of every eligible patch-embedding convolution of ``model`` (stride == kernel,
    no padding, dilation 1, groups 1). Returns how many were swapped; ``conv._bma_orig_forward`` keeps
    the original and ``restore_patch_embed`` puts it back."""
    n = 0
    for mod in model.modules():
        if (isinstance(mod, torch.nn.Conv2d) and tuple(mod.stride) == tuple(mod.kernel_size)
                and mod.padding in ((0, 0), "valid") and tuple(mod.dilation) == (1, 1) and mod.groups == 1
                and not hasattr(mod, "_bma_orig_forward")):
            mod._bma_orig_forward = mod.forward
            mod.forward = types.MethodType(_patch_conv_forward, mod)
            n += 1
    return n


def restore_patch_embed(model) -> None:
    for mod in model.modules():
        if hasattr(mod, "_bma_orig_forward"):
            mod.forward = mod._bma_orig_forward
            del mod._bma_orig_forward


def sample_ids_from_grad(ids: torch.Tensor, grad: torch.Tensor, search_width: int, topk: int = 256,
                         n_replace: int = 1, not_allowed_ids: Optional[torch.Tensor] = None,
                         mask_bits: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Device sampler with the reference's signature (:130-163). Consumes torch's CUDA generator
    exactly like the reference's two draws and leaves it at the same offset. Unlike the reference
    it does not write +inf into ``grad``."""
    native.require_cuda(grad, "grad")
    dev = grad.device
    if mask_bits is None and not_allowed_ids is not None:
        mask_bits = native.pack_mask_bits(not_allowed_ids, grad.shape[1], dev)
    gen = torch.cuda.default_generators[dev.index if dev.index is not None else torch.cuda.current_device()]
    seed, offset = gen.initial_seed(), gen.get_offset()
    new_ids, advance, _ = native.sample_ids(ids, grad, search_width, topk, n_replace, mask_bits, seed, offset)
    gen.set_offset(offset + advance)
    return new_ids


def filter_ids(ids: torch.Tensor, tokenizer) -> torch.Tensor:
    """Keep rows that survive decode -> encode (:166-186). Same result as the reference's
    per-row loop, but one batched tokenizer call and one device round trip."""
    host = ids.cpu()
    texts = tokenizer.batch_decode(host)
    again = tokenizer(texts, add_special_tokens=False)["input_ids"]
    rows = host.tolist()
    keep = [i for i, (a, b) in enumerate(zip(rows, again)) if a == b]
    if not keep:
        raise RuntimeError(
            "No token sequences are the same after decoding and re-encoding. "
            "Consider setting filter_ids=False or trying a different optim_str_init")
    if len(keep) == len(rows):
        return ids
    return ids[torch.tensor(keep, device=ids.device)]


class _PngWriter:
    """Background PNG encoder: (pinned uint8 HWC tensor, CUDA event, path) items from a queue."""

    def __init__(self):
        import queue
        import threading
        self.q = queue.Queue()
        self.error = None
        self.thread = threading.Thread(target=self._work, name="bma-png-writer", daemon=True)
        self.thread.start()

    def _work(self):
        from PIL import Image
        while True:
            item = self.q.get()
            try:
                if item is None:
                    return
                host, done, path = item
                if done is not None:
                    done.synchronize()
                Image.fromarray(host.numpy()).save(path)
            except Exception as exc:                   # noqa: BLE001 - surfaced by flush()
                self.error = exc
            finally:
                self.q.task_done()

    def submit(self, host, done, path):
        self.q.put((host, done, path))

    def flush(self):
        self.q.join()
        if self.error is not None:
            err, self.error = self.error, None
            raise err


class _PhaseClock:
    """CUDA-event stopwatch for the phases of one step: ``mark(name)`` closes the interval since the
    previous mark and files it under ``name`` (None = not a phase); ``read()`` returns
    {name: [seconds, ...]} once the events have completed. Falls back to perf_counter off-GPU."""

    def __init__(self, device):
        self.cuda = torch.device(device).type == "cuda"
        self.device = device
        self.marks = []

    def _now(self):
        if not self.cuda:
            return time.perf_counter()
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(torch.cuda.current_stream(self.device))
        return ev

    def start(self):
        self.marks = [(None, self._now())]

    def mark(self, name):
        self.marks.append((name, self._now()))

    def read(self) -> Dict[str, List[float]]:
        out: Dict[str, List[float]] = {}
        if self.cuda and self.marks:
            self.marks[-1][1].synchronize()
        for (_, a), (name, b) in zip(self.marks, self.marks[1:]):
            if name is not None:
                out.setdefault(name, []).append(a.elapsed_time(b) * 1e-3 if self.cuda else b - a)
        return out


@dataclass
class CandidateBatch:
    """What ``_build_input_embeds*`` returns here instead of a materialised (sw, L, d) tensor: the
    segment order plus the per-candidate ids; the scorer builds only the tail rows on device."""
    order: List[str]
    sampled_ids: torch.Tensor
    image: Optional[torch.Tensor]
    search_width: Optional[int]

    @property
    def shape(self):
        return (self.sampled_ids.shape[0],)


class BimodalAttack:
    def __init__(self, model, tokenizer, processor, config: BimodalAttackConfig, normalize=None,
                 engine: Optional[ScoringEngine] = None):
        self.model, self.tokenizer, self.processor = model, tokenizer, processor
        self.config, self.normalize = config, normalize
        self.embedding_layer = model.get_input_embeddings()
        self.device = self.embedding_layer.weight.device
        native.load()                                   # fail loudly when the CUDA library is missing
        native.require_cuda(self.embedding_layer.weight, "model")
        self.not_allowed_ids = None if config.allow_non_ascii else get_nonascii_toks(tokenizer, device=self.device)
        self._mask_bits = native.pack_mask_bits(self.not_allowed_ids, self.embedding_layer.num_embeddings, self.device)
        self.stop_flag = False
        self.model_type = getattr(model.config, "model_type", "")
        self.gemma_processor = processor.__class__.__name__ == "Gemma3Processor"
        if not getattr(tokenizer, "chat_template", None):          # :233-249
            tpl = ("USER: <image>\n{{ messages[0]['content'][0]['text'] }} \nASSISTANT: " if config.pgd_attack
                   else "{% for message in messages %}{{ message['content'] }}{% endfor %}")
            logger.warning("Tokenizer does not have a chat template. Using custom chat template.")
            tokenizer.chat_template = tpl
            self.processor.chat_template = tpl
        # the engine holds the packed decoder weights and the scorer workspace; it can be shared
        # between attacks on the same model (one per (goal, target) pair in the reference's driver)
        self.engine = engine if engine is not None else ScoringEngine(model)
        self._plans: Dict[Tuple[str, ...], TailPlan] = {}
        self.world = bdist.world()
        if self.world.size > 1:
            # the library's own NCCL communicator: split prefill + the winner all-reduce on the scorer's stream
            self.engine.attach_comm(bdist.native_comm(self.device, self.world))
        self.gather_all_losses = False    # multi-rank: also all-reduce the per-candidate loss vector
        self.balance_shards = os.environ.get("BMA_BALANCE_SHARDS", "1") != "0"   # multi-rank: shards of equal ROWS (shard())
        # ... weighted by the ranks' measured speeds (rebalance()): GPUs of one box run at different clocks under
        # the power cap, and every layer's prefix exchange makes all ranks wait for the slowest
        # (opt-in, BMA_ADAPT_SHARES=1: at N = 2 the persistent skew is ~1 % of a step, inside the run-to-run noise)
        self.adapt_shares = os.environ.get("BMA_ADAPT_SHARES", "0") == "1"
        self.rank_share: Optional[List[float]] = None
        self.step_hook = None             # optional callable(step_index) run at the end of every step
        self.graph_gradient = os.environ.get("BMA_GRAPH_GRADIENT", "1") != "0"
        # patch-embedding convolutions run as unfold + GEMM while run() is active (see _patch_conv_forward);
        # run() puts the model's own forward back before it returns
        self.patch_embed_gemm = config.pgd_attack and os.environ.get("BMA_PATCH_EMBED_GEMM", "1") != "0"
        self._grad_graph = None
        self._feat_graph = None
        self._png = None
        import inspect
        self._keep_logits = "logits_to_keep" in inspect.signature(model.forward).parameters

    # ---------------------------------------------------------------------------------------------
    # prompt -> segments (:271-393)
    def _prepare(self, messages, target: str):
        cfg, tok, dev = self.config, self.tokenizer, self.device
        msgs = [{"role": "user", "content": messages}] if isinstance(messages, str) else copy.deepcopy(messages)
        last = msgs[-1]
        if isinstance(last["content"], str) and "{optim_str}" not in last["content"]:
            last["content"] = last["content"] + " {optim_str}"
        if cfg.pgd_attack:
            if isinstance(last["content"], str):
                last["content"] = [{"type": "text", "text": last["content"]}, {"type": "image"}]
            elif isinstance(last["content"], list) and not any(c.get("type") == "image" for c in last["content"]):
                last["content"].append({"type": "image"})
        prompt = self.processor.apply_chat_template(msgs, add_generation_prompt=True)
        if tok.bos_token and prompt.startswith(tok.bos_token):
            prompt = prompt.replace(tok.bos_token, "")

        def enc(text, special):
            kw = {"padding": False} if special else {"add_special_tokens": False}
            return tok(text, return_tensors="pt", **kw)["input_ids"].to(dev, torch.int64)

        ids: Dict[str, torch.Tensor] = {}
        if cfg.pgd_attack:
            if self.gemma_processor:
                before, rest = prompt.split("{optim_str}", 1)
                if "<start_of_image>" not in rest:
                    raise ValueError("Expected <start_of_image> token in Gemma PGD prompt.")
                head, sep, tail = rest.partition("<start_of_image>")
                b_img, b_suf, after = before.strip(), (head + sep).strip(), tail.strip()
            else:
                if "<start_of_image>" in prompt:
                    b_img, rest = prompt.split("<start_of_image>", 1)
                elif "<image>" in prompt:
                    b_img, rest = prompt.split("<image>", 1)
                else:
                    raise ValueError("No image token found in prompt for PGD attack")
                b_suf, after = rest.split("{optim_str}", 1)
            ids["before_img"], ids["before_suffix"] = enc(b_img, True), enc(b_suf, True)
            ids["after"] = enc(after, False)
        else:
            before, after = prompt.split("{optim_str}")
            ids["before"], ids["after"] = enc(before, True), enc(after, False)
        ids["target"] = enc(target, False)
        self.segment_ids = ids
        self.target_ids = ids["target"]
        vocab = self.embedding_layer.num_embeddings
        if self.target_ids.numel() == 0 or int(self.target_ids.min()) < 0 or int(self.target_ids.max()) >= vocab:
            raise ValueError("target tokenises to ids outside the vocabulary (or to nothing)")
        with torch.no_grad():
            self.segment_embeds = {k: self.embedding_layer(v) for k, v in ids.items()}
        # reference attribute names (:378-393)
        for k, v in ids.items():
            setattr(self, f"{k}_ids", v)
        for k, v in self.segment_embeds.items():
            setattr(self, f"{k}_embeds", v)
        self._plans.clear()
        self.engine.prefill_tag = None
        self._grad_graph = None          # the captured gradient step references the old segment embeddings

    # ---------------------------------------------------------------------------------------------
    # phase A (:953-1028)
    def compute_gradient(self, optim_ids: torch.Tensor, image: Optional[torch.Tensor] = None):
        """(one-hot gradient [1, L_s, V], image gradient) like the reference. The HF forward+backward
        at batch 1 is ~1500 small kernels and CPU-launch bound, so after two eager calls it is replayed
        from a CUDA graph (same kernels, same results); ``graph_gradient = False`` keeps it eager."""
        if self.graph_gradient and self.device.type == "cuda":
            out = self._gradient_graphed(optim_ids, image)
            if out is not None:
                return out
        return self._gradient_eager(optim_ids, image)

    def _native_gradient(self):
        """The native forward/backward (gradient.py) when this model is covered, else None (HF autograd)."""
        if not hasattr(self, "_native_grad"):
            self._native_grad = None
            self._native_grad_why = "disabled (BMA_NATIVE_GRAD=0)"
            if os.environ.get("BMA_NATIVE_GRAD", "1") != "0" and self.device.type == "cuda":
                from .gradient import NativeGradient
                self._native_grad_why = NativeGradient.supported(self.engine)
                if self._native_grad_why is None and self.target_ids.numel() <= 128:
                    self._native_grad = NativeGradient(self.engine)
        return self._native_grad

    def _native_tower(self):
        """The native CLIP tower + projector (gradient.NativeVisionTower) when covered, else None."""
        if not hasattr(self, "_native_vt"):
            self._native_vt = None
            self._native_vt_why = "disabled (BMA_NATIVE_TOWER=0)"
            if (self.config.pgd_attack and not self.gemma_processor and self.device.type == "cuda"
                    and os.environ.get("BMA_NATIVE_TOWER", "1") != "0"):
                from .gradient import NativeVisionTower
                self._native_vt_why = NativeVisionTower.supported(self.model, self.normalize)
                if self._native_vt_why is None:
                    self._native_vt = NativeVisionTower(self.model, self.normalize)
        return self._native_vt

    def _gradient_native(self, ng, optim_ids: torch.Tensor, image: Optional[torch.Tensor]):
        """compute_gradient on this package's kernels: decoder forward + input-gradient backward native
        (gradient.py); the token gradient is the tcgen05 GEMM; the image gradient continues from the
        image-feature gradient through the vision tower."""
        cfg, model, W = self.config, self.model, self.embedding_layer.weight.detach()
        e = self.segment_embeds
        optim_embeds = W[optim_ids.reshape(-1)].unsqueeze(0)           # the one-hot product of :959-968 is a row gather
        feats = None
        tower = self._native_tower() if cfg.pgd_attack else None
        if cfg.pgd_attack:
            if tower is not None:
                with torch.no_grad():
                    feats = tower.forward(image)             # raw pixels in: Normalize is folded into the im2col
            else:
                with torch.enable_grad():
                    feats = image_features(model, self._normalize_pixels(image), self.gemma_processor)
            seq = [e["before_img"], feats.detach().to(W.dtype), e["before_suffix"], optim_embeds, e["after"], e["target"]]
        else:
            seq = [e["before"], optim_embeds, e["after"], e["target"]]
        offs, at = {}, 0
        for name, t in zip(("a", "img", "b", "optim") if cfg.pgd_attack else ("a", "optim"), seq):
            offs[name] = (at, at + t.shape[1])
            at += t.shape[1]
        with torch.no_grad():
            dx0 = ng.forward_backward(torch.cat(seq, dim=1)[0], self.target_ids)
        onehot_grad = g_image = None
        if cfg.gcg_attack:
            lo, hi = offs["optim"]
            onehot_grad = native.token_grad(dx0[lo:hi], W).unsqueeze(0)
        if cfg.pgd_attack:
            lo, hi = offs["img"]
            if tower is not None:
                with torch.no_grad():
                    g_image = tower.backward(dx0[lo:hi]).view_as(image)
            else:
                g_image = torch.autograd.grad(feats, image, grad_outputs=dx0[lo:hi].reshape(feats.shape).to(feats.dtype))[0]
        return onehot_grad, g_image

    def _gradient_eager(self, optim_ids: torch.Tensor, image: Optional[torch.Tensor]):
        ng = self._native_gradient()
        if ng is not None:
            return self._gradient_native(ng, optim_ids, image)
        return self._gradient_hf(optim_ids, image)

    def _gradient_hf(self, optim_ids: torch.Tensor, image: Optional[torch.Tensor]):
        """HF autograd at batch 1 (models the native step does not cover, and the A/B reference for it)."""
        cfg, model, emb = self.config, self.model, self.embedding_layer
        W = emb.weight
        # the reference multiplies a one-hot by W (:959-968): a row gather, and -- also for Gemma --
        # WITHOUT the embedding scale. Here the gathered rows are the autograd leaf.
        optim_embeds = W.detach()[optim_ids].clone()
        if cfg.gcg_attack:
            optim_embeds.requires_grad_()
        e = self.segment_embeds
        if cfg.pgd_attack:
            feats = image_features(model, self._normalize_pixels(image), self.gemma_processor)
            seq = [e["before_img"], feats, e["before_suffix"], optim_embeds, e["after"], e["target"]]  # :981-991
        else:
            seq = [e["before"], optim_embeds, e["after"], e["target"]]
        x = torch.cat(seq, dim=1)
        T = self.target_ids.shape[1]
        if self._keep_logits:
            # only the T rows that predict the target are needed (the reference computes all L rows and
            # slices, :1003-1008); row values do not depend on which other rows go through the head
            pred = model(inputs_embeds=x, logits_to_keep=T + 1).logits[..., :-1, :].contiguous()
        else:
            logits = model(inputs_embeds=x).logits
            shift = x.shape[1] - T
            pred = logits[..., shift - 1:-1, :].contiguous()
        loss = F.cross_entropy(pred.view(-1, pred.size(-1)), self.target_ids.view(-1))
        g_embeds = g_image = None
        if cfg.pgd_attack and cfg.gcg_attack:
            g_embeds, g_image = torch.autograd.grad(loss, (optim_embeds, image))
        elif cfg.pgd_attack:
            g_image = torch.autograd.grad(loss, image)[0]
        elif cfg.gcg_attack:
            g_embeds = torch.autograd.grad(loss, optim_embeds)[0]
        onehot_grad = None
        if g_embeds is not None:
            # (1) token gradient: G_embeds . W^T on the tcgen05 GEMM, W streamed once
            # (an fp16 model: bf16 copies of both operands -- wider exponent range, 3 mantissa bits fewer)
            onehot_grad = native.token_grad(g_embeds[0].to(torch.bfloat16), self.engine._w(emb.weight)).unsqueeze(0)
        return onehot_grad, g_image

    def _gradient_graphed(self, optim_ids: torch.Tensor, image: Optional[torch.Tensor]):
        """CUDA-graph replay of ``_gradient_eager`` on static input/output buffers. Returns None
        while warming up (first two calls) or when capture is not possible (then stays eager)."""
        key = (tuple(optim_ids.shape), None if image is None else tuple(image.shape), id(self.segment_embeds))
        st = self._grad_graph
        if st is None or st.get("key") != key:
            st = self._grad_graph = {"key": key, "calls": 0, "graph": None, "failed": False}
        if st["failed"]:
            return None
        st["calls"] += 1
        if st["graph"] is None:
            if st["calls"] <= 2:
                return None                            # eager warm-up (cuBLAS/cuDNN handles, TMA maps, allocator)
            try:
                ids_s = optim_ids.detach().clone()
                img_s = None if image is None else image.detach().clone().requires_grad_()
                side = torch.cuda.Stream(device=self.device)
                side.wait_stream(torch.cuda.current_stream(self.device))
                with torch.cuda.stream(side):
                    self._gradient_eager(ids_s, img_s)     # once more on the capture stream's allocator pool
                torch.cuda.current_stream(self.device).wait_stream(side)
                torch.cuda.synchronize(self.device)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    out_tok, out_img = self._gradient_eager(ids_s, img_s)
                st.update(graph=graph, ids=ids_s, img=img_s, out_tok=out_tok, out_img=out_img)
            except Exception as exc:                   # noqa: BLE001 - any capture problem: stay eager
                logger.warning(f"CUDA-graph capture of the gradient step failed ({type(exc).__name__}: {exc}); "
                               "staying on the eager path")
                if os.environ.get("BMA_GRAPH_DEBUG"):
                    import traceback
                    traceback.print_exc()
                st["failed"] = True
                torch.cuda.synchronize(self.device)
                return None
        with torch.no_grad():
            st["ids"].copy_(optim_ids)
            if image is not None:
                st["img"].copy_(image.detach())
        st["graph"].replay()
        out_tok = None if st["out_tok"] is None else st["out_tok"].clone()
        out_img = None if st["out_img"] is None else st["out_img"].clone()
        return out_tok, out_img

    # phase B (:1030-1037)
    def perform_pgd_step(self, image: torch.Tensor, eps: float, alpha: float, image_grad: torch.Tensor,
                         image_original: torch.Tensor) -> torch.Tensor:
        new = image.detach().clone()
        # a torchvision Normalize exposes mean/std: the kernel then also emits the normalised pixels
        mean, std = getattr(self.normalize, "mean", None), getattr(self.normalize, "std", None)
        self._pixel_values = native.pgd_step_(new, image_grad.contiguous(), image_original, alpha * eps, eps,
                                              mean, std)
        self._pixel_values_of = new
        return new.requires_grad_()

    def _normalize_pixels(self, image: torch.Tensor) -> torch.Tensor:
        """``self.normalize(image)``. A torchvision ``Normalize`` builds its mean/std tensors from Python
        lists on every call (a host-to-device copy, which also cannot be captured in a CUDA graph); here
        they are cached on the device and applied with the same two ops (sub, then div)."""
        mean, std = getattr(self.normalize, "mean", None), getattr(self.normalize, "std", None)
        if mean is None or std is None or not image.is_floating_point():
            return self.normalize(image)
        cache = getattr(self, "_norm_consts", None)
        if cache is None or cache[0].device != image.device or cache[0].dtype != image.dtype:
            m = torch.as_tensor(mean, dtype=image.dtype, device=image.device).view(-1, 1, 1)
            d = torch.as_tensor(std, dtype=image.dtype, device=image.device).view(-1, 1, 1)
            cache = self._norm_consts = (m, d)
        return (image - cache[0]) / cache[1]

    def scoring_features(self, image: torch.Tensor) -> torch.Tensor:
        """Image features for the scoring phase (no gradient): the native CLIP tower + projector on the raw
        pixels when the model is covered, else normalise (reusing the PGD kernel's fused output when it is
        current) + the HF tower. Either way ~100-400 small launches at batch 1, so after two eager calls the
        pass is replayed from a CUDA graph on a static input buffer."""
        tower = self._native_tower()
        if tower is not None:
            src = image.detach()
            fn = lambda t: tower.forward(t, keep=False).clone()                     # noqa: E731
        else:
            src = self._normalized(image)
            fn = lambda t: image_features(self.model, t, self.gemma_processor)   # noqa: E731
        st = self._feat_graph
        key = (tuple(src.shape), src.dtype, tower is not None)
        if not self.graph_gradient or st is None or st.get("key") != key:
            st = self._feat_graph = {"key": key, "calls": 0, "graph": None, "failed": not self.graph_gradient}
        st["calls"] += 1
        if st["failed"] or st["calls"] <= 2:
            with torch.no_grad():
                return fn(src)
        if st["graph"] is None:
            try:
                with torch.no_grad():
                    src_s = src.detach().clone()
                    side = torch.cuda.Stream(device=self.device)
                    side.wait_stream(torch.cuda.current_stream(self.device))
                    with torch.cuda.stream(side):
                        fn(src_s)
                    torch.cuda.current_stream(self.device).wait_stream(side)
                    torch.cuda.synchronize(self.device)
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph):
                        out = fn(src_s)
                st.update(graph=graph, pix=src_s, out=out)
            except Exception as exc:                   # noqa: BLE001
                logger.warning(f"CUDA-graph capture of the vision tower failed ({type(exc).__name__}: {exc}); "
                               "staying on the eager path")
                st["failed"] = True
                torch.cuda.synchronize(self.device)
                with torch.no_grad():
                    return fn(src)
        with torch.no_grad():
            st["pix"].copy_(src)
            st["graph"].replay()
            return st["out"].clone()

    def _normalized(self, image: torch.Tensor) -> torch.Tensor:
        """Normalised pixels for the scoring phase; reuses the PGD kernel's fused output."""
        if getattr(self, "_pixel_values_of", None) is image and self._pixel_values is not None:
            return self._pixel_values
        return self._normalize_pixels(image.detach())

    # phase D sampling (:908-951)
    def candidate_sampling(self, config, i, optim_ids, optim_ids_onehot_grad, tokenizer, sampling_times,
                           total_sampling_time):
        width = config.search_width
        if config.dynamic_search:
            width = max(config.min_search_width, int(config.search_width * (1 - i / config.num_steps)))
        if not config.gcg_attack:
            return optim_ids, 1, 0.0, total_sampling_time
        t0 = time.perf_counter()
        sampled = sample_ids_from_grad(optim_ids.squeeze(0), optim_ids_onehot_grad.squeeze(0), width,
                                       config.topk, config.n_replace, mask_bits=self._mask_bits)
        if config.filter_ids:
            sampled = filter_ids(sampled, tokenizer)
        # the scorer computes the rows a candidate shares with the suffix it was sampled from once
        self._base = (sampled, optim_ids.reshape(-1).to(torch.int64).contiguous())
        dt = time.perf_counter() - t0
        sampling_times.append(dt)
        return sampled, sampled.shape[0], dt, total_sampling_time + dt

    # ---------------------------------------------------------------------------------------------
    # candidate assembly (:1112-1276): returns a descriptor, nothing is repeated sw times
    def _build_input_embeds(self, sampled_ids, image=None, search_width=None, mode="gcg", single=False,
                            no_joint_eval=False, no_target=False) -> CandidateBatch:
        order = segment_order(mode, self.model_type, single, no_joint_eval, no_target)
        return CandidateBatch(order, sampled_ids, image, search_width)

    def _build_input_embeds_pgd(self, sampled_ids, image, search_width, single=False):
        return self._build_input_embeds(sampled_ids, image=image, search_width=search_width, mode="pgd", single=single)

    def _build_input_embeds_gcg(self, sampled_ids, search_width, single=False, no_joint_eval=False, no_target=False):
        return self._build_input_embeds(sampled_ids, search_width=search_width, mode="gcg", single=single,
                                        no_joint_eval=no_joint_eval, no_target=no_target)

    def _build_input_embeds_gcg_pgd(self, sampled_ids, image, search_width=None, single=False, no_target=False):
        return self._build_input_embeds(sampled_ids, image=image, search_width=search_width, mode="gcg_pgd",
                                        single=single, no_target=no_target)

    def materialize(self, batch: CandidateBatch) -> torch.Tensor:
        """Dense (B, L, d) embeddings of a CandidateBatch -- only for HF ``generate`` in
        ``debug_output`` (:745-777); never on the scoring path."""
        parts = []
        n = batch.sampled_ids.shape[0]
        for name in batch.order:
            if name == "optim":
                t = self.embedding_layer(batch.sampled_ids)
            elif name == "image":
                t = batch.image
            else:
                t = self.segment_embeds[name]
            parts.append(t.expand(n, -1, -1) if t.shape[0] == 1 else t)
        return torch.cat(parts, dim=1)

    def _plan(self, batch: CandidateBatch) -> TailPlan:
        if "target" not in batch.order:
            raise ValueError("scoring needs a target segment")
        parts = dict(self.segment_embeds)
        if batch.image is not None:
            parts["image"] = batch.image
        key = tuple(batch.order)
        has_image = "image" in batch.order
        plan = None if has_image else self._plans.get(key)
        if plan is None:
            fixed = {n: parts[n] for n in batch.order if n != "optim"}
            plan = plan_tail(batch.order, fixed, batch.sampled_ids.shape[1])
            if not has_image:
                self._plans[key] = plan
        return plan

    # scoring (:1278-1310). search_batch_size is accepted for interface compatibility only: the native
    # workspace is sized up front and the scorer chunks internally, so the reference's OOM-driven
    # halving wrapper (utils.py:57-115, find_executable_batch_size) has nothing to do here and is not ported.
    def _compute_candidates_loss_original(self, search_batch_size: int, batch: CandidateBatch,
                                          gather: bool = False) -> torch.Tensor:
        plan = self._plan(batch)
        eng = self.engine
        tag = (id(self), tuple(batch.order),
               None if batch.image is None else (batch.image.data_ptr(), batch.image._version), plan.n_prefix)
        if tag != eng.prefill_tag or "image" in batch.order:
            eng.prefill(plan.prefix)      # deferred: rides in the first chunk of the score call below
            eng.prefill_tag = tag
        ids = batch.sampled_ids
        eng.want_all_match = bool(self.config.early_stop)
        eng.reset_key()
        loss = torch.full((ids.shape[0],), float("inf"), dtype=torch.float32, device=self.device)
        base = getattr(self, "_base", None)
        base_ids = base[1] if base is not None and base[0] is ids else None
        lo, hi, hint = self.shard(ids, base_ids, plan)       # hint = the largest shard: every rank sizes its workspace alike
        # with a split prefill every rank takes part in the per-layer exchange, candidates or not
        if hi > lo or (self.world.size > 1 and base_ids is not None):
            loss[lo:hi] = eng.score(ids[lo:hi], plan, self.target_ids, idx_base=lo, reset=False, base_ids=base_ids,
                                    batch_hint=hint)
        bdist.reduce_winner(eng.key, eng.flag if self.config.early_stop else None, self.world)
        if self.world.size > 1 and (gather or self.gather_all_losses):
            bdist.gather_losses(loss, lo, hi, self.world)
        if self.config.early_stop and int(eng.flag.item()) != 0:
            self.stop_flag = True
        return loss

    def shard(self, ids: torch.Tensor, base_ids: Optional[torch.Tensor], plan: TailPlan) -> Tuple[int, int, int]:
        """(lo, hi, largest shard) of this rank's contiguous slice of the candidates. With a base suffix the
        shards are balanced by decoder ROWS, not by count: a candidate first replaced at position j owns the
        tail rows from j on (suffix-head sharing), so equal counts would leave the ranks with unequal work.
        Every rank derives the same cuts from the (broadcast) ids."""
        w, n = self.world, int(ids.shape[0])
        if w.size == 1:
            return 0, n, n
        even = -(-n // w.size)
        if base_ids is None or not self.balance_shards or n < 2 * w.size or not self.engine.share_suffix_head:
            lo, hi = bdist.shard_bounds(n, w)
            return lo, hi, even
        diff = ids != base_ids.view(1, -1)
        first = torch.where(diff.any(1), diff.int().argmax(1), ids.shape[1] + plan.n_tail)   # == base: no rows at all
        cuts = bdist.balanced_cuts((plan.n_tail - first).clamp_(min=0) + 1, w.size, self.rank_share)
        widest = max(b - a for a, b in zip(cuts, cuts[1:]))
        # a stable capacity hint (the workspace is re-created, collectively, only when it grows)
        return cuts[w.rank], cuts[w.rank + 1], max(widest, even + even // 4 + 1)

    def rebalance(self, step_ms: float) -> Optional[List[float]]:
        """Collective, after a scoring step (every rank, same call order): all-gathers (this rank's scoring time,
        the time it spent waiting for other ranks in the prefix exchange) and moves the candidate shares towards
        the ranks' measured speeds (``dist.updated_shares``); the next ``shard()`` uses them. Results do not depend
        on the cuts (rows are computed independently). No-op on one rank, when disabled, or when nothing was
        exchanged (no split prefill)."""
        w = self.world
        if w.size == 1 or not self.adapt_shares or not self.balance_shards:
            return None
        self._rebalance_calls = getattr(self, "_rebalance_calls", 0) + 1
        if self._rebalance_calls <= 2:       # the first steps carry one-time costs (workspace, graph capture) that differ per rank
            self.engine.exchange_wait_ms(reset=True)
            return self.rank_share
        mine = torch.tensor([float(step_ms), self.engine.exchange_wait_ms(reset=True)], dtype=torch.float32, device=self.device)
        allr = torch.empty(2 * w.size, dtype=torch.float32, device=self.device)
        torch.distributed.all_gather_into_tensor(allr, mine)
        vals = allr.view(w.size, 2).tolist()
        steps, waits = [v[0] for v in vals], [v[1] for v in vals]
        if max(waits) <= 0.0 or min(steps) <= 0.0:
            return self.rank_share
        share = self.rank_share if self.rank_share is not None else [1.0 / w.size] * w.size
        self.rank_share = bdist.updated_shares(share, steps, waits)
        return self.rank_share

    def _winner(self) -> Tuple[float, int]:
        """(min loss, argmin) from the packed key the scorer (and the all-reduce) left on device."""
        return unpack_key(int(self.engine.key.item()))

    def _save_image(self, image, path):
        """Per-step PNG of the adversarial image (:743-744, :1312-1317), off the critical path: the
        uint8 conversion (same truncation as the reference's ``(arr * 255).astype(uint8)``) runs on
        the device, the copy into pinned memory is asynchronous and a writer thread encodes the file."""
        if self._png is None:
            self._png = _PngWriter()
        img = image.detach().squeeze(0)
        u8 = (img * 255).to(torch.uint8).permute(1, 2, 0).contiguous()
        host = torch.empty(u8.shape, dtype=torch.uint8, pin_memory=u8.is_cuda)
        host.copy_(u8, non_blocking=True)
        done = None
        if u8.is_cuda:
            done = torch.cuda.Event()
            done.record(torch.cuda.current_stream(u8.device))
        self._png.submit(host, done, path)

    def flush_images(self):
        """Block until every image handed to ``_save_image`` is on disk (called before run() returns)."""
        if self._png is not None:
            self._png.flush()

    # ---------------------------------------------------------------------------------------------
    def init_buffer(self, image) -> AttackBuffer:
        """Initial suffix and its loss (:826-906)."""
        cfg, tok, dev = self.config, self.tokenizer, self.device
        buf = AttackBuffer(cfg.buffer_size)
        if isinstance(cfg.optim_str_init, str):
            init = tok(cfg.optim_str_init, add_special_tokens=False, return_tensors="pt")["input_ids"].to(dev)
            if cfg.buffer_size > 1:
                pool = tok(INIT_CHARS, add_special_tokens=False, return_tensors="pt")["input_ids"].squeeze().to(dev)
                pick = torch.randint(0, pool.shape[0], (cfg.buffer_size - 1, init.shape[1]))
                init = torch.cat([init, pool[pick]], dim=0)
        else:
            if len(cfg.optim_str_init) != cfg.buffer_size:
                logger.warning(f"Using {len(cfg.optim_str_init)} initializations but buffer size is set to {cfg.buffer_size}")
            init = tok(cfg.optim_str_init, add_special_tokens=False, return_tensors="pt")["input_ids"].to(dev)
        n = max(1, cfg.buffer_size)
        with torch.no_grad():
            if cfg.pgd_attack:
                feats = self.scoring_features(image)
                batch = self._build_input_embeds_gcg_pgd(init, feats, search_width=n, single=True)
            else:
                batch = self._build_input_embeds_gcg(init, search_width=n, no_joint_eval=True)
            # every rank needs every entry's loss: the buffer (and the first suffix) must agree across ranks
            losses = self._compute_candidates_loss_original(n, batch, gather=True)
        for j, value in enumerate(losses.tolist()):
            buf.add(value, init[[j]])
        buf.log_buffer(tok, logger)
        return buf

    # ---------------------------------------------------------------------------------------------
    def run(self, messages: Union[str, List[dict]], goal: str, target: str,
            image: Optional[torch.Tensor] = None) -> BimodalAttackResult:
        """The step loop (:441-794). Process- and model-wide switches this path flips for speed are put
        back before returning: torch's fill_uninitialized_memory flag and the patch-embedding forward."""
        fill_flag = torch.utils.deterministic.fill_uninitialized_memory
        patched = install_patch_embed_gemm(self.model) if self.patch_embed_gemm else 0
        try:
            return self._run(messages, goal, target, image)
        finally:
            torch.utils.deterministic.fill_uninitialized_memory = fill_flag
            if patched:
                restore_patch_embed(self.model)

    def _run(self, messages, goal: str, target: str, image: Optional[torch.Tensor]) -> BimodalAttackResult:
        cfg, model, tok = self.config, self.model, self.tokenizer
        self.initial_prompt = goal
        if cfg.pgd_attack and cfg.images_folder:
            os.makedirs(cfg.images_folder, exist_ok=True)
        if cfg.seed is not None:
            from transformers import set_seed
            set_seed(cfg.seed)
            torch.use_deterministic_algorithms(True, warn_only=True)     # as the reference does (:267-269)
            # ... but without torch's debug aid that memsets every freshly allocated tensor in this mode:
            # it changes no result and doubles the kernel count of the batch-1 gradient step
            torch.utils.deterministic.fill_uninitialized_memory = False
        if cfg.pgd_attack:
            if image is None:
                raise ValueError("pgd_attack=True needs an image")
            native.require_cuda(image, "image")
        self._prepare(messages, target)
        buffer = self.init_buffer(image)
        optim_ids = buffer.get_best_ids()

        losses, strings, suffixes, outputs = [], [], [], []
        g_times, s_times, l_times, p_times, totals = [], [], [], [], []
        tot_s = 0.0
        best_loss = float("inf")
        self.best_optim_ids = None
        self.best_image = None
        current = None
        if cfg.pgd_attack:
            logger.warning(f"Using alpha: {cfg.alpha}, eps: {cfg.eps}")
            image.requires_grad = True                       # the reference flags the caller's tensor (:425)
            original = image.detach().clone()

        def record(loss_value, ids, img):
            nonlocal best_loss
            losses.append(loss_value)
            strings.append(tok.batch_decode(ids)[0])
            if buffer.size == 0 or loss_value < buffer.get_highest_loss():
                buffer.add(loss_value, ids)
            if loss_value < best_loss:
                best_loss = loss_value
                self.best_optim_ids = ids.clone()
                if img is not None:
                    self.best_image = img.detach().clone()

        lead = self.world.rank == 0         # gradient, PGD and sampling run on rank 0 only
        n_tok = optim_ids.shape[1]
        # Phase times are DEVICE times: CUDA events on the launching stream bracket every phase and are
        # read once per step, after the step's last host sync (the reference's perf_counter deltas measure
        # launch time only once work is asynchronous or replayed from a graph).
        clock = _PhaseClock(self.device)
        i = -1
        for i in range(cfg.num_steps):
            grad_time = pgd_time = sampling_time = 0.0
            sampled, width = None, 0
            clock.start()
            if lead:
                onehot_grad, image_grad = self.compute_gradient(optim_ids, image if cfg.pgd_attack else None)
                clock.mark("gradient")
                if cfg.pgd_attack and not cfg.pgd_after_gcg:
                    image = self.perform_pgd_step(image, cfg.eps, cfg.alpha, image_grad, original)
                    clock.mark("pgd")
                    if cfg.gcg_attack and not cfg.joint_eval:        # phase C (:481-491)
                        onehot_grad, image_grad = self.compute_gradient(optim_ids, image)
                        clock.mark("gradient")
                sampled, width, _, _ = self.candidate_sampling(cfg, i, optim_ids, onehot_grad, tok, [], 0.0)
                clock.mark("sampling")
            if self.world.size > 1:
                sampled = bdist.broadcast_candidates(sampled, n_tok, self.device, self.world)
                width = sampled.shape[0]
                if cfg.pgd_attack and not cfg.pgd_after_gcg:
                    # rank 0 sends its own storage; the others receive into a copy, never into the caller's tensor
                    shared = bdist.broadcast_tensor(image.detach() if lead else image.detach().clone(), self.world)
                    image = image if lead else shared
                if cfg.gcg_attack:      # every rank knows the current suffix: the scorer's sharing base
                    self._base = (sampled, optim_ids.reshape(-1).to(torch.int64).contiguous())

            clock.mark(None)                # broadcasts belong to no phase
            with torch.no_grad():
                chunk = width if cfg.batch_size is None else cfg.batch_size
                if cfg.pgd_attack:
                    feats = self.scoring_features(image)
                    best_idx = 0
                    scored_with_image = False
                    if cfg.joint_eval:
                        build = self._build_input_embeds_gcg_pgd if cfg.pgd_after_gcg else self._build_input_embeds_pgd
                        batch = build(sampled, feats, search_width=width, single=True)
                        self._compute_candidates_loss_original(chunk, batch)
                        cand_min, best_idx = self._winner()
                        scored_with_image = True
                    elif cfg.gcg_attack:
                        batch = self._build_input_embeds_gcg(sampled, search_width=width, single=True)
                        self._compute_candidates_loss_original(chunk, batch)
                        cand_min, best_idx = self._winner()
                    chosen = sampled[best_idx].unsqueeze(0)
                    if not cfg.pgd_after_gcg:
                        if scored_with_image:
                            # the reference re-scores the winner alone (:605-612); the same sequence was
                            # just scored inside the batch, so its fp32 loss is reused
                            current = cand_min
                        else:
                            full = self._build_input_embeds_gcg_pgd(chosen, feats)
                            self._compute_candidates_loss_original(1, full)
                            current, _ = self._winner()
                        optim_ids = chosen
                        record(current, optim_ids, image)
                else:
                    batch = self._build_input_embeds_gcg(sampled, search_width=width, no_joint_eval=True)
                    self._compute_candidates_loss_original(chunk, batch)
                    current, best_idx = self._winner()
                    optim_ids = sampled[best_idx].unsqueeze(0)
                    record(current, optim_ids, None)
                clock.mark("loss")

            if cfg.pgd_after_gcg and cfg.pgd_attack:             # phases E, F (:668-740)
                if lead:
                    _, image_grad = self.compute_gradient(optim_ids, image)   # still the previous suffix (:673)
                    clock.mark("gradient")
                    image = self.perform_pgd_step(image, cfg.eps, cfg.alpha, image_grad, original)
                    clock.mark("pgd")
                shared = bdist.broadcast_tensor(image.detach() if lead else image.detach().clone(), self.world)
                image = image if lead else shared
                clock.mark(None)
                with torch.no_grad():
                    feats = self.scoring_features(image)
                    full = self._build_input_embeds_gcg_pgd(chosen, feats)
                    self._compute_candidates_loss_original(1, full)
                    current, _ = self._winner()
                    optim_ids = chosen
                    record(current, optim_ids, image)
                clock.mark("loss")

            # one read per step: every phase's events are complete after the winner came back to the host
            phases = clock.read()
            for name, dst in (("gradient", g_times), ("pgd", p_times), ("sampling", s_times), ("loss", l_times)):
                dst.extend(phases.get(name, []))
            if self.world.size > 1 and phases.get("loss"):
                self.rebalance(1e3 * sum(phases["loss"]))
            grad_time = phases["gradient"][-1] if phases.get("gradient") else 0.0
            pgd_time = phases["pgd"][-1] if phases.get("pgd") else 0.0
            sampling_time = sum(phases.get("sampling", []))
            loss_time = sum(phases.get("loss", []))

            if cfg.pgd_attack and cfg.images_folder:
                self._save_image(image, os.path.join(cfg.images_folder, f"{i}.png"))
            gen_output = ""
            if cfg.debug_output and i % 10 == 0:
                with torch.no_grad():
                    if cfg.pgd_attack:
                        feats = self.scoring_features(image)
                        dbg = self._build_input_embeds_gcg_pgd(sampled, feats, search_width=width, no_target=True)
                    else:
                        dbg = self._build_input_embeds_gcg(sampled, search_width=width, no_target=True)
                    gen = model.generate(inputs_embeds=self.materialize(dbg), max_new_tokens=120)
                    gen_output = tok.decode(gen[0], skip_special_tokens=True)
            outputs.append(gen_output)
            suffixes.append(tok.batch_decode(optim_ids)[0])
            buffer.log_buffer(tok, logger)
            if self.stop_flag:
                logger.info("Early stopping due to finding a perfect match.")
                break
            if self.step_hook is not None:
                self.step_hook(i)
            totals.append(grad_time + sampling_time + pgd_time + loss_time)
            logger.info(f"[Iteration {i}] loss {current:.4f} | grad {grad_time:.4f}s sample {sampling_time:.4f}s "
                        f"pgd {pgd_time:.4f}s loss {loss_time:.4f}s")

        self.flush_images()
        self.final_image = image
        self.final_optim_ids = optim_ids
        best = losses.index(min(losses))
        return BimodalAttackResult(best_loss=losses[best], best_string=strings[best], losses=losses,
                                   strings=strings, adversarial_suffixes=suffixes, model_outputs=outputs,
                                   gradient_times=g_times, sampling_times=s_times, loss_times=l_times,
                                   pgd_times=p_times, total_times=totals)


def run(model, tokenizer, processor, messages: Union[str, List[dict]], goal: str, target: str,
        image: Optional[torch.Tensor] = None, config: Optional[BimodalAttackConfig] = None,
        normalize=None) -> BimodalAttackResult:
    """Drop-in for the reference's ``bimodalattack.run`` (:1323-1338)."""
    if config is None:
        config = BimodalAttackConfig()
    logger.setLevel(getattr(logging, config.verbosity))
    return BimodalAttack(model, tokenizer, processor, config, normalize).run(messages, goal, target, image)
