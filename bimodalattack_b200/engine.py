"""Python side of the native candidate scorer (C ABI: ``bma_scorer_*`` in include/bimodal_b200.h).

``ScoringEngine`` takes a Hugging Face LLaVA-style model object, packs the decoder weights into the
layouts the tcgen05 GEMM epilogues expect (fused QKV rows; gate/up rows interleaved per 128 so the
SwiGLU epilogue sees matching columns in one tile), builds the rotary tables with the model's own
rotary module, and drives the scorer. All arithmetic happens in the shared library; there is no
torch fallback for any of it.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Tuple

import torch

from . import native
from .segments import TailPlan

KEY_IDX_BITS = 31
KEY_INIT = (1 << 63) - 1


def unpack_key(key: int) -> Tuple[float, int]:
    """Inverse of the device-side packing: (loss, global candidate index)."""
    idx = key & ((1 << KEY_IDX_BITS) - 1)
    order = (key >> KEY_IDX_BITS) & 0xFFFFFFFF
    bits = order & 0x7FFFFFFF if order & 0x80000000 else (~order) & 0xFFFFFFFF
    import struct
    return struct.unpack("<f", struct.pack("<I", bits))[0], idx


def find_decoder(model):
    """(decoder stack, lm_head, text_config) for transformers 4.50 and 5.x module layouts."""
    inner = getattr(model, "model", None)
    lm = getattr(inner, "language_model", None) if inner is not None else None
    if lm is None:                       # 4.50: model.language_model is a *ForCausalLM
        causal = getattr(model, "language_model", None)
        if causal is None:
            raise RuntimeError("bimodal_b200: cannot locate the language model inside the HF object")
        lm = getattr(causal, "model", causal)
    head = model.get_output_embeddings()
    cfg = getattr(model.config, "text_config", model.config)
    return lm, head, cfg


class ScoringEngine:
    def __init__(self, model, max_rows: int = 0, max_prefix: int = 0, max_positions: int = 0):
        self.lib = native.load()
        self.model = model
        lm, head, cfg = find_decoder(model)
        self.lm, self.head, self.cfg = lm, head, cfg
        self.device = next(lm.parameters()).device
        native.require_cuda(next(lm.parameters()), "model")
        # The kernels compute in bf16 (north_star). An fp16 model -- the reference loads LLaVA / LLaVA-RC in
        # float16, utils/experiments_utils.py:91-100 -- is ACCEPTED: the scorer works on bf16 copies of the decoder
        # weights (``_w``), its inputs are cast on entry and the per-candidate losses then agree with the fp16
        # reference to bf16 tolerance, not to fp16's. The model object itself is not touched.
        self.model_dtype = next(lm.parameters()).dtype
        if self.model_dtype not in (torch.bfloat16, torch.float16):
            raise RuntimeError(f"bimodal_b200 scorer computes in bf16 and accepts bf16 / fp16 models; model is {self.model_dtype}")
        self.dtype = torch.bfloat16
        self._shadow = {}
        mt = getattr(cfg, "model_type", "")
        if mt not in ("llama", "gemma3_text"):
            raise RuntimeError(f"bimodal_b200 scorer: decoder type '{mt}' is not supported (llama, gemma3_text)")
        self.gemma = mt == "gemma3_text"
        self.d = cfg.hidden_size
        self.n_layers = cfg.num_hidden_layers
        self.n_heads = cfg.num_attention_heads
        self.n_kv = getattr(cfg, "num_key_value_heads", self.n_heads) or self.n_heads
        self.head_dim = getattr(cfg, "head_dim", None) or self.d // self.n_heads
        self.d_ff = cfg.intermediate_size
        self.vocab = head.weight.shape[0]
        if getattr(cfg, "attention_bias", False) or getattr(cfg, "mlp_bias", False):
            raise RuntimeError("bimodal_b200 scorer: projection biases are not supported")
        act = getattr(cfg, "hidden_activation", None) or getattr(cfg, "hidden_act", "silu")
        if act not in ("silu", "gelu_pytorch_tanh"):
            raise RuntimeError(f"bimodal_b200 scorer: activation '{act}' is not supported")
        self.act = 0 if act == "silu" else 1
        self.embed = model.get_input_embeddings()
        self.embed_scale, self.attn_scale = 1.0, float(self.head_dim) ** -0.5
        self.max_seq = 1 << 30
        if self.gemma:
            # Gemma-3 decoder (HF modeling_gemma3.py): (1+w) RMSNorm, q/k-norm, 4 norms per layer, GeGLU,
            # scaled embeddings, softmax scale query_pre_attn_scalar^-0.5, local/global rotary tables.
            if getattr(cfg, "attn_logit_softcapping", None) is not None:
                raise RuntimeError("bimodal_b200 scorer: attn_logit_softcapping is not supported")
            if getattr(cfg, "use_bidirectional_attention", False):
                raise RuntimeError("bimodal_b200 scorer: bidirectional attention is not supported")
            if type(model).__name__.endswith("ForCausalLM") and getattr(cfg, "final_logit_softcapping", None) is not None:
                raise RuntimeError("bimodal_b200 scorer: final_logit_softcapping is not supported")
            scale = getattr(self.embed, "embed_scale", None)
            if scale is not None:        # HF multiplies by the scale cast to the weight dtype
                self.embed_scale = float(torch.as_tensor(scale).to(self.dtype).float())
            self.attn_scale = float(cfg.query_pre_attn_scalar) ** -0.5
            if "sliding_attention" in getattr(cfg, "layer_types", []):
                # the local layers' window never binds while the whole sequence fits inside it
                self.max_seq = int(cfg.sliding_window)
        self._pack_weights()
        self._scorer = None
        self._cap_rows = 0
        self._cap_prefix = 0
        self._want_rows = max_rows
        self._want_prefix = max_prefix
        self._max_positions = max_positions or 4096
        self._build_rope(self._max_positions)
        self.key = torch.empty(1, dtype=torch.int64, device=self.device)
        self.flag = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.want_all_match = True       # track the early-stop flag (reference :1300-1306); off saves the argmax pass
        self.prefix_len = 0
        self.share_suffix_head = True    # compute tail rows shared with the base suffix once (see score())
        self.dedupe_candidates = True    # ... and the rows of identical candidates once
        self.fused_publish = None        # None: library default (on; BMA_FUSED_PUBLISH=0 turns it off). Set BEFORE the
                                         # first score call: prefix / base K/V leave through the QKV GEMM's epilogue
        self.early_signal = None         # None: library default (off; BMA_EARLY_SIGNAL=1). See set_early_signal()
        # fused prefill: the prefix positions ride in the first chunk of the next score() call as extra
        # rows (one [prefix | base | candidates] row space per layer) instead of a separate pass
        self.fused_prefill = True
        self.prefill_tag = None          # what the resident prefix K/V were computed from (set by the attack)
        self.comm = None                 # native NCCL communicator (dist.native_comm) for multi-rank runs
        # multi-rank split prefill through NVLink peer memory (cudaIpc-mapped prefix buffers + flags)
        # instead of an NCCL all-gather per layer; BMA_PEER_MEMORY=0 keeps the all-gather (A/B runs)
        import os
        self.peer_memory = os.environ.get("BMA_PEER_MEMORY", "1") != "0"
        self._last_prefix: Optional[torch.Tensor] = None
        self._prefix_stale = False

    # -------------------------------------------------------------------------------------------
    def _w(self, t: torch.Tensor) -> torch.Tensor:
        """The tensor the kernels read for parameter ``t``: ``t`` itself for a bf16 model (borrowed, never
        copied), a cached bf16 copy for an fp16 model."""
        if t.dtype == torch.bfloat16:
            return t
        c = self._shadow.get(id(t))
        if c is None:
            c = self._shadow[id(t)] = t.detach().to(torch.bfloat16).contiguous()
        return c

    def _pack_weights(self):
        self.w_qkv, self.w_gate_up = [], []
        with torch.no_grad():
            for layer in self.lm.layers:
                a, m = layer.self_attn, layer.mlp
                self.w_qkv.append(torch.cat([a.q_proj.weight, a.k_proj.weight, a.v_proj.weight], 0).to(self.dtype).contiguous())
                g, u = m.gate_proj.weight.to(self.dtype), m.up_proj.weight.to(self.dtype)
                if self.d_ff % 128 != 0:
                    raise RuntimeError("bimodal_b200 scorer: intermediate_size must be a multiple of 128")
                gu = torch.stack([g.view(self.d_ff // 128, 128, self.d), u.view(self.d_ff // 128, 128, self.d)], 1)
                self.w_gate_up.append(gu.reshape(2 * self.d_ff, self.d).contiguous())

    def _build_rope(self, max_pos: int):
        """Rotary tables from the model's own rotary module, one (cos, sin) pair per layer type."""
        with torch.no_grad():
            pos = torch.arange(max_pos, device=self.device).unsqueeze(0)
            dummy = torch.zeros(1, dtype=self.dtype, device=self.device)
            half = self.head_dim // 2
            self.rope = {}
            kinds = sorted(set(self.cfg.layer_types)) if self.gemma else [None]
            for kind in kinds:
                cos, sin = (self.lm.rotary_emb(dummy, pos) if kind is None
                            else self.lm.rotary_emb(dummy, pos, layer_type=kind))   # [1, max_pos, head_dim]
                self.rope[kind] = (cos[0, :, :half].float().contiguous(), sin[0, :, :half].float().contiguous())

    def _layer_rope(self, i: int):
        return self.rope[self.cfg.layer_types[i] if self.gemma else None]

    def _ensure(self, rows: int, prefix: int):
        if self._scorer is not None and rows <= self._cap_rows and prefix <= self._cap_prefix:
            return
        if self._scorer is not None:
            self.lib.bma_scorer_destroy(self._scorer)
            self._scorer = None
        cap_rows = max(rows, self._want_rows, 128)
        cap_prefix = max(prefix, self._want_prefix, 1)
        desc = native.ModelDesc(
            n_layers=self.n_layers, d_model=self.d, n_heads=self.n_heads, n_kv_heads=self.n_kv,
            head_dim=self.head_dim, d_ff=self.d_ff, vocab=self.vocab,
            rms_eps=float(getattr(self.cfg, "rms_norm_eps", 1e-6)), norm_mode=int(self.gemma), act=self.act,
            embed_scale=self.embed_scale, attn_scale=self.attn_scale,
            max_rows=cap_rows, max_prefix=cap_prefix, qk_norm=int(self.gemma), sandwich_norm=int(self.gemma))
        with torch.cuda.device(self.device):
            handle = self.lib.bma_scorer_create(C.byref(desc))
        if not handle:
            raise RuntimeError("bimodal_b200: scorer_create failed: " +
                               self.lib.bma_last_error().decode("utf-8", "replace"))
        self._scorer = handle
        native.check(self.lib.bma_scorer_set_dedupe(handle, int(self.dedupe_candidates)), "scorer_set_dedupe")
        if self.fused_publish is not None:
            native.check(self.lib.bma_scorer_set_fused_publish(handle, int(self.fused_publish)), "scorer_set_fused_publish")
        if self.early_signal is not None:
            native.check(self.lib.bma_scorer_set_early_signal(handle, int(self.early_signal)), "scorer_set_early_signal")
        if self.comm is not None:
            native.check(self.lib.bma_scorer_set_comm(handle, self.comm), "scorer_set_comm")
            if self.peer_memory:
                # collective: every rank re-creates its scorer at the same call (capacities depend only on
                # quantities that are equal on all ranks), so the handle exchange lines up
                from . import dist as bdist
                bdist.exchange_peer_handles(self.lib, handle, self.device)
        self._cap_rows, self._cap_prefix = cap_rows, cap_prefix
        self._prefix_stale = True        # a new scorer has no prefix K/V yet
        for i, layer in enumerate(self.lm.layers):
            cos, sin = self._layer_rope(i)
            w = native.LayerWeights(
                w_qkv=self.w_qkv[i].data_ptr(), w_o=self._w(layer.self_attn.o_proj.weight).data_ptr(),
                w_gate_up=self.w_gate_up[i].data_ptr(), w_down=self._w(layer.mlp.down_proj.weight).data_ptr(),
                norm_attn=self._w(layer.input_layernorm.weight).data_ptr(),
                rope_cos=cos.data_ptr(), rope_sin=sin.data_ptr())
            if self.gemma:       # HF names: post_attention_layernorm is the POST-attention norm here
                w.norm_mlp = self._w(layer.pre_feedforward_layernorm.weight).data_ptr()
                w.norm_post_attn = self._w(layer.post_attention_layernorm.weight).data_ptr()
                w.norm_post_mlp = self._w(layer.post_feedforward_layernorm.weight).data_ptr()
                w.q_norm = self._w(layer.self_attn.q_norm.weight).data_ptr()
                w.k_norm = self._w(layer.self_attn.k_norm.weight).data_ptr()
            else:                # Llama: post_attention_layernorm is the pre-MLP norm
                w.norm_mlp = self._w(layer.post_attention_layernorm.weight).data_ptr()
            if not (layer.self_attn.o_proj.weight.is_contiguous() and layer.mlp.down_proj.weight.is_contiguous()):
                raise RuntimeError("bimodal_b200 scorer: non-contiguous projection weight")
            native.check(self.lib.bma_scorer_set_layer(handle, i, C.byref(w)), "scorer_set_layer")
        ew = self._w(self.embed.weight)
        native.check(self.lib.bma_scorer_set_head(handle, self._w(self.lm.norm.weight).data_ptr(),
                                                  self._w(self.head.weight).data_ptr(), ew.data_ptr(), ew.stride(0)),
                     "scorer_set_head")

    def close(self):
        if self._scorer is not None:
            self.lib.bma_scorer_destroy(self._scorer)
            self._scorer = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    PROFILE_CLASSES = ("tail_embed", "rmsnorm", "gemm_qkv_rope", "attention", "gemm_o_proj",
                       "gemm_gate_up_glu", "gemm_down", "gemm_lm_head_ce", "ce_finalize")

    def profile(self, enable: bool, max_events: int = 16384):
        """Per-kernel-class CUDA-event timing inside the scoring pass (used by bench.py's roofline)."""
        self._ensure(max(self._cap_rows, 1), self._cap_prefix)
        native.check(self.lib.bma_scorer_profile(self._scorer, int(enable), max_events), "scorer_profile")

    def profile_read(self) -> dict:
        n = len(self.PROFILE_CLASSES)
        ms, fl, cnt = (C.c_double * n)(), (C.c_double * n)(), (C.c_longlong * n)()
        native.check(self.lib.bma_scorer_profile_read(self._scorer, ms, fl, cnt, n), "scorer_profile_read")
        return {name: {"ms": ms[i], "flops": fl[i], "launches": int(cnt[i])}
                for i, name in enumerate(self.PROFILE_CLASSES)}

    @property
    def workspace_bytes(self) -> int:
        return 0 if self._scorer is None else int(self.lib.bma_scorer_workspace_bytes(self._scorer))

    # -------------------------------------------------------------------------------------------
    def reserve(self, rows: int, prefix: int):
        self._ensure(rows, prefix)

    def attach_comm(self, comm):
        """Native communicator (``dist.native_comm``): the fused prefill is then split over its ranks.
        Collective when a scorer already exists (peer handles are exchanged)."""
        self.comm = comm
        if self._scorer is not None:
            native.check(self.lib.bma_scorer_set_comm(self._scorer, comm), "scorer_set_comm")
            if comm is None:
                native.check(self.lib.bma_scorer_clear_peers(self._scorer), "scorer_clear_peers")
            elif self.peer_memory:
                from . import dist as bdist
                bdist.exchange_peer_handles(self.lib, self._scorer, self.device)

    def set_early_signal(self, enable: bool) -> None:
        """Split prefill over peer memory: the QKV GEMM publishes first and raises this rank's flag itself (True)
        or one signal + wait launch follows it (False, default). Collective in effect: set it on every rank alike."""
        self.early_signal = bool(enable)
        if self._scorer is not None:
            native.check(self.lib.bma_scorer_set_early_signal(self._scorer, int(self.early_signal)), "scorer_set_early_signal")

    def peers_barrier(self, count: int = 1) -> None:
        """``count`` data-free prefix exchanges on the current stream (collective over the ranks; no-op without peers)."""
        if self._scorer is not None:
            native.check(self.lib.bma_scorer_peers_barrier(self._scorer, count, torch.cuda.current_stream().cuda_stream),
                         "peers_barrier")

    def exchange_wait_ms(self, reset: bool = True) -> float:
        """Milliseconds this rank's scorer stream spent waiting for the other ranks inside the per-layer prefix
        exchange since the last reset (0 without peers). Synchronises the device. The rank whose figure stays near
        zero is the one the others wait for (``BimodalAttack.rebalance``)."""
        if self._scorer is None:
            return 0.0
        import ctypes as C
        ns = C.c_ulonglong(0)
        native.check(self.lib.bma_scorer_exchange_wait_ns(self._scorer, C.byref(ns), 1 if reset else 0), "exchange_wait_ns")
        return ns.value * 1e-6

    def prefill(self, prefix_embeds: Optional[torch.Tensor], deferred: Optional[bool] = None):
        """Compute and keep the K/V of the shared prefix ([P, d] bf16; None/empty = no prefix). With
        ``fused_prefill`` (default) the work is deferred into the next ``score`` call that carries
        ``base_ids``; any other score call consumes it with a plain prefill pass first."""
        P = 0 if prefix_embeds is None else prefix_embeds.shape[0]
        deferred = self.fused_prefill if deferred is None else deferred
        # room for the base suffix's keys behind the prefix (suffix-head sharing, see score())
        self._ensure(max(self._cap_rows, P, 1), P + (64 if self.share_suffix_head else 0))
        if P > 0:
            if P + 1 > self._max_positions:
                raise RuntimeError("bimodal_b200 scorer: sequence longer than the rotary table")
            pe = prefix_embeds.to(self.dtype).contiguous()
        with torch.cuda.device(self.device):
            if deferred:
                native.check(self.lib.bma_scorer_prefill_deferred(self._scorer, None if P == 0 else pe.data_ptr(), P),
                             "scorer_prefill_deferred")
            else:
                native.check(self.lib.bma_scorer_prefill(self._scorer, None if P == 0 else pe.data_ptr(), P,
                                                         native.stream_ptr(self.device)), "scorer_prefill")
        self._pending_rows = P if deferred else 0
        self._last_deferred = deferred
        self.prefix_len = P
        self._last_prefix = None if P == 0 else pe
        self._prefix_stale = False

    def _fixed(self, plan: TailPlan) -> torch.Tensor:
        """``plan.fixed`` in the compute dtype (a cast copy for an fp16 model; alive until the call is enqueued)."""
        f = plan.fixed
        return f if f.dtype == self.dtype else f.to(self.dtype)

    def reset_key(self):
        with torch.cuda.device(self.device):
            native.check(self.lib.bma_scorer_reset_key(self._scorer, self.key.data_ptr(), self.flag.data_ptr(),
                                                       native.stream_ptr(self.device)), "scorer_reset_key")

    def score(self, ids: torch.Tensor, plan: TailPlan, target_ids: torch.Tensor, idx_base: int = 0,
              reset: bool = True, base_ids: Optional[torch.Tensor] = None, batch_hint: int = 0) -> torch.Tensor:
        """Per-candidate mean CE (fp32, [B]). ``self.key`` receives the packed (loss, index) minimum
        and ``self.flag`` the early-stop indicator. ``base_ids`` ([n_optim] int64): the suffix the
        candidates were sampled from; rows a candidate shares with it are computed once."""
        native.require_cuda(ids, "ids")
        B = ids.shape[0]
        T = target_ids.numel()
        S, row0 = self._tail_rows(plan, T)
        self._check_length(S)
        share = self._share_rows(ids, base_ids)
        # batch_hint: the largest shard of a multi-rank call, so every rank sizes (and re-creates) alike
        self._ensure(max(self._cap_rows, (min(max(B, batch_hint), 4096) + (1 if share else 0)) * S + self._fused_rows(share)),
                     max(self._cap_prefix, self.prefix_len + share))
        if self._prefix_stale:           # the workspace was re-created after the last prefill
            self.prefill(self._last_prefix, deferred=getattr(self, "_last_deferred", None))
        ids = ids.to(torch.int64).contiguous()
        tgt = target_ids.reshape(-1).to(torch.int64).contiguous()   # validated by the caller
        loss = torch.empty(B, dtype=torch.float32, device=self.device)
        fixed = self._fixed(plan)
        if reset:
            self.reset_key()
        with torch.cuda.device(self.device):
            native.check(self.lib.bma_scorer_score(
                self._scorer, ids.data_ptr(), ids.stride(0), B, plan.slot.data_ptr(), fixed.data_ptr(),
                S, tgt.data_ptr(), T, row0, loss.data_ptr(), self.key.data_ptr(),
                self.flag.data_ptr() if self.want_all_match else None,
                idx_base, None if not share else base_ids.data_ptr(), plan.n_id_rows,
                native.stream_ptr(self.device)), "scorer_score")
        self._pending_rows = 0
        return loss

    def _fused_rows(self, share: int) -> int:
        """Extra row capacity a deferred prefix needs inside a ragged chunk (reserved whenever the fused
        prefill is on, so a workspace re-created between prefill() and score() keeps its chunk size)."""
        return self.prefix_len if (share and self.fused_prefill) else 0

    def _share_rows(self, ids: torch.Tensor, base_ids: Optional[torch.Tensor]) -> int:
        """Extra prefix keys suffix-head sharing needs (0 = sharing off)."""
        if base_ids is None or not self.share_suffix_head:
            return 0
        if base_ids.dtype != torch.int64 or not base_ids.is_contiguous() or base_ids.numel() != ids.shape[1]:
            raise RuntimeError("bimodal_b200 scorer: base_ids must be a contiguous int64 vector of the suffix length")
        return int(ids.shape[1])

    @property
    def last_rows(self) -> int:
        """Tail rows the last score call ran through the decoder."""
        return 0 if self._scorer is None else int(self.lib.bma_scorer_last_rows(self._scorer))

    def _check_length(self, S: int):
        if self.prefix_len + S > self._max_positions:
            raise RuntimeError("bimodal_b200 scorer: sequence longer than the rotary table")
        if self.prefix_len + S + 1 > self.max_seq:
            raise RuntimeError("bimodal_b200 scorer: sequence longer than the sliding window of the local layers")

    @staticmethod
    def _tail_rows(plan: TailPlan, T: int):
        """(rows to compute, first CE row). The tail ends with the T target tokens; the last one's
        row predicts nothing that is scored and is attended to by no other row, so it is dropped."""
        if plan.n_tail < T + 1:
            raise RuntimeError("bimodal_b200 scorer: the tail must hold the target plus one preceding token")
        S = plan.n_tail - 1
        return S, S - T

    def score_host(self, ids_host: torch.Tensor, plan: TailPlan, target_ids: torch.Tensor,
                   loss_host: torch.Tensor, key_host: torch.Tensor, idx_base: int = 0,
                   base_ids_host: Optional[torch.Tensor] = None, batch_hint: int = 0):
        """End-to-end entry: pinned host ids in, pinned host losses + key out (synchronises)."""
        B, T = ids_host.shape[0], target_ids.numel()
        S, row0 = self._tail_rows(plan, T)
        self._check_length(S)
        share = self._share_rows(ids_host, base_ids_host)
        Bh = max(B, batch_hint)
        self._ensure(max(self._cap_rows, (min(Bh, 4096) + (1 if share else 0)) * S + self._fused_rows(share),
                         Bh * ids_host.shape[1]),
                     max(self._cap_prefix, self.prefix_len + share))
        if self._prefix_stale:
            self.prefill(self._last_prefix, deferred=getattr(self, "_last_deferred", None))
        fixed = self._fixed(plan)
        with torch.cuda.device(self.device):
            native.check(self.lib.bma_scorer_score_host(
                self._scorer, ids_host.data_ptr(), ids_host.stride(0), B, plan.slot.data_ptr(),
                fixed.data_ptr(), S, target_ids.data_ptr(), T, row0, loss_host.data_ptr(),
                key_host.data_ptr(), idx_base, None if not share else base_ids_host.data_ptr(), plan.n_id_rows,
                native.stream_ptr(self.device)), "scorer_score_host")
        self._pending_rows = 0


def group_score(engines, ids: torch.Tensor, plan: TailPlan, target_ids: torch.Tensor,
                base_ids: torch.Tensor, use_peers: bool = True) -> torch.Tensor:
    """Single-process form of the multi-rank scoring path (C ABI ``bma_scorer_group_score``): the
    engines act as the ranks of one box on one device -- contiguous candidate shards, the deferred
    prefix split over them, the per-layer K/V all-gather done as local copies. Every engine must have
    been given the same prefix with ``prefill(..., deferred=True)``. Returns the [B] losses; the winner
    key / early-stop flag land in ``engines[0].key`` / ``.flag``."""
    e0 = engines[0]
    B, T = ids.shape[0], target_ids.numel()
    S, row0 = e0._tail_rows(plan, T)
    share = e0._share_rows(ids, base_ids)
    for e in engines:
        e._check_length(S)
        e._ensure(max(e._cap_rows, (min(B, 4096) + 1) * S + e._fused_rows(share)),
                  max(e._cap_prefix, e.prefix_len + share))
        if e._prefix_stale:
            e.prefill(e._last_prefix, deferred=True)
    ids = ids.to(torch.int64).contiguous()
    tgt = target_ids.reshape(-1).to(torch.int64).contiguous()
    loss = torch.empty(B, dtype=torch.float32, device=e0.device)
    e0.reset_key()
    handles = (C.c_void_p * len(engines))(*[e._scorer for e in engines])
    fixed = e0._fixed(plan)
    with torch.cuda.device(e0.device):
        native.check(e0.lib.bma_scorer_group_score(
            handles, len(engines), ids.data_ptr(), ids.stride(0), B, plan.slot.data_ptr(), fixed.data_ptr(),
            S, tgt.data_ptr(), T, row0, loss.data_ptr(), e0.key.data_ptr(), e0.flag.data_ptr(),
            base_ids.data_ptr(), plan.n_id_rows, int(use_peers), native.stream_ptr(e0.device)), "scorer_group_score")
    for e in engines:
        e._pending_rows = 0
    return loss
