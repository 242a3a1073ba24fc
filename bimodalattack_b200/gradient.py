"""Native gradient step: the batch-1 forward + INPUT-gradient backward of
``BimodalAttack.compute_gradient`` (reference bimodal_attack.py:953-1028) on this package's kernels.

The reference runs the HF model under autograd and asks only for the gradients of the loss w.r.t.
the one-hot suffix and the raw image (:1016-1025) -- no weight gradients. That backward is a chain of
``dX = dY . W`` contractions (the same tcgen05 GEMM as the scorer, fed with transposed weight copies),
the per-head attention products (its batched mode), and row kernels for everything in between
(csrc/phase_a.cu), keeping the rounding points of the bf16 autograd graph. The host below only strings
C-ABI calls together on the current stream; ``BimodalAttack`` captures the whole step into one CUDA
graph after warm-up.

Covered: Llama-style decoders with multi-head attention (LLaVA-1.5 / LLaVA-RC) and CLIP vision towers
with the LLaVA projector. Anything else (Gemma-3 decoders, SigLIP towers, GQA) keeps the HF autograd
path in ``attack.py``.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import native
from .engine import ScoringEngine

N_SM = 148


def _pick_tile(M: int, N: int) -> Tuple[int, int]:
    """(block_n, ctas) with the fewest, fullest waves for an M x N output on 148 SMs."""
    best, best_cost = (128, 1), float("inf")
    for bn, ctas, eff in ((256, 2, 1.0), (256, 1, 0.95), (160, 1, 0.92), (128, 1, 0.88)):
        bm = 128 * ctas
        tiles = -(-M // bm) * -(-N // bn)
        waves = -(-tiles // (N_SM // ctas))
        cost = waves * bm * bn / ctas / eff          # time per SM ~ tile area it owns per wave
        if cost < best_cost:
            best, best_cost = (bn, ctas), cost
    return best


class _Lib:
    """Thin typed front of the C ABI used here (tensors in, raw pointers out)."""

    def __init__(self, device):
        self.lib = native.load()
        self.device = device

    @property
    def st(self):
        return native.stream_ptr(self.device)

    def gemm(self, A, B, out=None, out32=None, M=None, N=None, K=None, lda=None, ldb=None, ldc=None, R=None, ldr=0,
             bias=None, tile=None, batch=1, a_row=0, a_col=0, b_row=0, b_col=0, c_batch=0, transpose_out=False,
             a_off=0, b_off=0, c_off=0, causal=0):
        """C = A . B^T with raw strides; ``*_off`` are element offsets into the tensors."""
        g = native.GemmDesc()
        g.A = A.data_ptr() + 2 * a_off
        g.lda = lda if lda is not None else A.stride(-2)
        g.B = B.data_ptr() + 2 * b_off
        g.ldb = ldb if ldb is not None else B.stride(-2)
        if out32 is not None:
            g.C, g.C32 = None, out32.data_ptr() + 4 * c_off
            g.ldc = ldc if ldc is not None else out32.stride(-2)
        else:
            g.C, g.C32 = out.data_ptr() + 2 * c_off, None
            g.ldc = ldc if ldc is not None else out.stride(-2)
        g.R = None if R is None else R.data_ptr()
        g.ldr = ldr if R is not None else 0
        g.bias = None if bias is None else bias.data_ptr()
        g.M, g.N, g.K = M, N, K
        bn, ctas = tile if tile is not None else _pick_tile(M, N)
        g.block_n, g.ctas = bn, ctas
        g.batch, g.a_row, g.a_col, g.b_row, g.b_col, g.c_batch = batch, a_row, a_col, b_row, b_col, c_batch
        g.transpose_out = int(transpose_out)
        g.causal = causal
        native.check(self.lib.bma_gemm_ex(C.byref(g), self.st), "gemm_ex")


class _AttentionProducts:
    """Per-head attention of ONE sequence as batched tcgen05 GEMMs + row kernels, forward and backward.
    Expects on ``self``: k (_Lib), H, hd, q_dim, qkv_dim, Lp and the buffers S32, VT, ATT, DATT, DATT_T,
    QK_T, DS, DS_T, P_T, DQKV. QKV is [L, Q | K | V] row-major, P is [H, L, Lp]."""

    def _attn_forward(self, QKV, P, L, scale, causal):
        k, lib, H, hd, qd, Lp = self.k, self.k.lib, self.H, self.hd, self.q_dim, self.Lp
        # S_h = Q_h K_h^T (fp32), heads side by side along the columns of QKV
        # (causal sequences: tiles above the diagonal are never computed, K ranges stop / start at the diagonal)
        k.gemm(QKV, QKV, out32=self.S32, M=L, N=L, K=hd, lda=self.qkv_dim, ldb=self.qkv_dim, ldc=Lp, tile=(128, 1),
               batch=H, a_col=hd, b_col=hd, b_off=qd, c_batch=L * Lp, causal=1 if causal else 0)
        native.check(lib.bma_softmax_rows(self.S32.data_ptr(), Lp, P.data_ptr(), Lp, H, L, scale, causal, k.st), "softmax")
        # O_h = P_h V_h: P stacked along the rows, V^T stacked along the rows
        k.gemm(P, self.VT, out=self.ATT, M=L, N=hd, K=L, lda=Lp, ldb=Lp, ldc=qd, tile=(128, 1),
               batch=H, a_row=L, b_row=hd, c_batch=hd, causal=2 if causal else 0)

    def _attn_backward(self, QKV, P, L, scale, causal):
        """DATT (gradient of the attention output) -> DQKV (gradients of Q | K | V, before any rotary)."""
        k, lib, H, hd, qd, Lp = self.k, self.k.lib, self.H, self.hd, self.q_dim, self.Lp
        # dP_h = dO_h V_h^T (fp32)
        k.gemm(self.DATT, QKV, out32=self.S32, M=L, N=L, K=hd, lda=qd, ldb=self.qkv_dim, ldc=Lp, tile=(128, 1),
               batch=H, a_col=hd, b_col=hd, b_off=2 * qd, c_batch=L * Lp, causal=1 if causal else 0)
        native.check(lib.bma_softmax_bwd_rows(P.data_ptr(), Lp, self.S32.data_ptr(), Lp, self.DS.data_ptr(), Lp,
                                              H, L, scale, causal, k.st), "softmax_bwd")
        # K-major operands of the three products: Q^T | K^T, dO^T, dS^T, P^T
        self._transpose(QKV, self.qkv_dim, 0, self.QK_T, Lp, 0, L, 2 * qd, 1)
        self._transpose(self.DATT, qd, 0, self.DATT_T, Lp, 0, L, qd, 1)
        self._transpose(self.DS, Lp, L * Lp, self.DS_T, Lp, L * Lp, L, L, H)
        self._transpose(P, Lp, L * Lp, self.P_T, Lp, L * Lp, L, L, H)
        # dQ_h = dS_h K_h ; dK_h = dS_h^T Q_h ; dV_h = P_h^T dO_h  -> the Q | K | V columns of DQKV
        k.gemm(self.DS, self.QK_T, out=self.DQKV, M=L, N=hd, K=L, lda=Lp, ldb=Lp, ldc=self.qkv_dim, tile=(128, 1),
               batch=H, a_row=L, b_row=hd, b_off=qd * Lp, c_batch=hd, causal=2 if causal else 0)
        k.gemm(self.DS_T, self.QK_T, out=self.DQKV, M=L, N=hd, K=L, lda=Lp, ldb=Lp, ldc=self.qkv_dim, tile=(128, 1),
               batch=H, a_row=L, b_row=hd, c_batch=hd, c_off=qd, causal=3 if causal else 0)
        k.gemm(self.P_T, self.DATT_T, out=self.DQKV, M=L, N=hd, K=L, lda=Lp, ldb=Lp, ldc=self.qkv_dim, tile=(128, 1),
               batch=H, a_row=L, b_row=hd, c_batch=hd, c_off=2 * qd, causal=3 if causal else 0)

    def _transpose(self, src, lds, src_batch, dst, ldd, dst_batch, rows, cols, batch, src_off=0):
        native.check(self.k.lib.bma_transpose_bf16(src.data_ptr() + 2 * src_off, lds, src_batch, dst.data_ptr(), ldd,
                                                   dst_batch, rows, cols, batch, self.k.st), "transpose")

    def _attn_buffers(self, L, z):
        H, qd, Lp = self.H, self.q_dim, self.Lp
        self.VT = z(qd, Lp)
        self.S32 = z(H, L, Lp, dt=torch.float32)
        self.ATT = z(L, qd)
        self.DATT = z(L, qd)
        self.DATT_T = z(qd, Lp)
        self.QK_T = z(2 * qd, Lp)
        self.DS = z(H, L, Lp)
        self.DS_T = z(H, L, Lp)
        self.P_T = z(H, L, Lp)
        self.DQKV = z(L, self.qkv_dim)


class NativeGradient(_AttentionProducts):
    """Forward + input-gradient backward of one [prefix | image | suffix | after | target] sequence."""

    @staticmethod
    def supported(engine: ScoringEngine) -> Optional[str]:
        """None when the decoder is covered, else the reason it is not."""
        if getattr(engine, "model_dtype", torch.bfloat16) != torch.bfloat16:
            return f"{engine.model_dtype} model (the native gradient step reads the bf16 parameters in place)"
        if engine.gemma:
            return "Gemma-3 decoder (sandwich norms, q/k-norm)"
        if engine.n_kv != engine.n_heads:
            return "grouped-query attention"
        if engine.head_dim not in (64, 128):
            return f"head_dim {engine.head_dim}"
        return None

    def __init__(self, engine: ScoringEngine):
        why = self.supported(engine)
        if why is not None:
            raise RuntimeError(f"bimodal_b200 native gradient: {why} is not covered")
        self.eng = engine
        self.dev = engine.device
        self.k = _Lib(self.dev)
        e = engine
        self.d, self.H, self.hd, self.dff, self.V = e.d, e.n_heads, e.head_dim, e.d_ff, e.vocab
        self.q_dim = self.H * self.hd
        self.qkv_dim = 3 * self.q_dim
        self.eps = float(getattr(e.cfg, "rms_norm_eps", 1e-6))
        with torch.no_grad():
            # dX = dY . W is C = A . B^T with B = W^T: transposed copies of every projection
            self.wT_qkv = [w.t().contiguous() for w in e.w_qkv]
            self.wT_gu = [w.t().contiguous() for w in e.w_gate_up]
            self.wT_o = [l.self_attn.o_proj.weight.t().contiguous() for l in e.lm.layers]
            self.wT_down = [l.mlp.down_proj.weight.t().contiguous() for l in e.lm.layers]
            self.headT = e.head.weight.t().contiguous()          # [d, V]
        self.L = 0

    # -------------------------------------------------------------------------------------------
    def _alloc(self, L: int, T: int):
        if L == self.L and T == getattr(self, "T", -1):
            return
        dev, bf = self.dev, torch.bfloat16
        nl, d, H, hd = self.eng.n_layers, self.d, self.H, self.hd
        Lp = (L + 7) // 8 * 8
        z = lambda *s, dt=bf: torch.zeros(s, dtype=dt, device=dev)   # noqa: E731
        self.L, self.T, self.Lp = L, T, Lp
        # saved for the backward
        self.X = z(nl + 1, L, d)
        self.XMID = z(nl, L, d)
        self.QKV = z(nl, L, self.qkv_dim)
        self.P = z(nl, H, L, Lp)
        self.GU = z(nl, L, 2 * self.dff)
        # scratch
        self.XN = z(L, d)
        self._attn_buffers(L, z)
        self.ACT = z(L, self.dff)
        self.XF = z(T, d)
        self.LOGITS = z(T, self.V)
        self.DLOGITS = z(T, self.V)
        self.DXF = z(T, d)
        self.LOSS = z(T, dt=torch.float32)
        self.DX = z(L, d)
        self.DXM = z(L, d)
        self.DXN = z(L, d)
        self.DACT = z(L, self.dff)
        self.DGU = z(L, 2 * self.dff)

    # -------------------------------------------------------------------------------------------
    def _rmsnorm(self, x, w, y, rows):
        native.check(self.k.lib.bma_rmsnorm(x.data_ptr(), x.stride(0), w.data_ptr(), y.data_ptr(), y.stride(0), rows,
                                            self.d, self.eps, 0, self.k.st), "rmsnorm")

    def _rmsnorm_bwd(self, dy, x, w, dres, dx, rows):
        native.check(self.k.lib.bma_rmsnorm_bwd(
            dy.data_ptr(), dy.stride(0), x.data_ptr(), x.stride(0), w.data_ptr(),
            None if dres is None else dres.data_ptr(), 0 if dres is None else dres.stride(0),
            dx.data_ptr(), dx.stride(0), rows, self.d, self.eps, 0, self.k.st), "rmsnorm_bwd")

    # -------------------------------------------------------------------------------------------
    def forward_backward(self, x0: torch.Tensor, target_ids: torch.Tensor) -> torch.Tensor:
        """x0: [L, d] bf16 input embeddings of the whole sequence (target last); returns dLoss/dx0 [L, d]
        (a view of an internal buffer, valid until the next call). Loss = mean CE of the target
        tokens given everything before them (:1003-1012)."""
        e, k, lib = self.eng, self.k, self.k.lib
        L, T = x0.shape[0], target_ids.numel()
        self._alloc(L, T)
        Lp, d, H, hd, qd, dff = self.Lp, self.d, self.H, self.hd, self.q_dim, self.dff
        st = k.st
        scale = e.attn_scale
        tgt = target_ids.reshape(-1).to(torch.int64).contiguous()
        self.X[0].copy_(x0)
        layers = e.lm.layers
        # ------------------------------------------------------------------ forward
        for l, layer in enumerate(layers):
            X, XMID, QKV, P, GU = self.X[l], self.XMID[l], self.QKV[l], self.P[l], self.GU[l]
            cos, sin = e._layer_rope(l)
            self._rmsnorm(X, layer.input_layernorm.weight, self.XN, L)
            native.check(lib.bma_qkv_rope(self.XN.data_ptr(), d, e.w_qkv[l].data_ptr(), QKV.data_ptr(), self.qkv_dim,
                                          self.VT.data_ptr(), Lp, cos.data_ptr(), sin.data_ptr(), L, d, qd, qd, hd, 0,
                                          st), "qkv_rope")
            self._attn_forward(QKV, P, L, scale, causal=1)
            k.gemm(self.ATT, layer.self_attn.o_proj.weight, out=XMID, M=L, N=d, K=qd, R=X, ldr=d)
            self._rmsnorm(XMID, layer.post_attention_layernorm.weight, self.XN, L)
            k.gemm(self.XN, e.w_gate_up[l], out=GU, M=L, N=2 * dff, K=d)
            native.check(lib.bma_glu(GU.data_ptr(), 2 * dff, self.ACT.data_ptr(), dff, L, dff, e.act, st), "glu")
            k.gemm(self.ACT, layer.mlp.down_proj.weight, out=self.X[l + 1], M=L, N=d, K=dff, R=XMID, ldr=d)
        # ------------------------------------------------------------------ loss head (T rows)
        r0 = L - T - 1                              # logits[L-T-1 : L-1] predict the target (:1005-1008)
        XL = self.X[len(layers)]
        self._rmsnorm(XL[r0:r0 + T], e.lm.norm.weight, self.XF, T)
        # logits^T = W_head . XF^T with the table streamed once (the token-gradient GEMM's shape)
        k.gemm(e.head.weight, self.XF, out=self.LOGITS, M=self.V, N=T, K=d, ldc=self.V, transpose_out=True)
        native.check(lib.bma_ce_grad(self.LOGITS.data_ptr(), self.V, tgt.data_ptr(), self.DLOGITS.data_ptr(), self.V,
                                     self.LOSS.data_ptr(), T, self.V, st), "ce_grad")
        k.gemm(self.headT, self.DLOGITS, out=self.DXF, M=d, N=T, K=self.V, ldc=d, transpose_out=True)
        self.DX.zero_()
        self._rmsnorm_bwd(self.DXF, XL[r0:r0 + T], e.lm.norm.weight, None, self.DX[r0:r0 + T], T)
        # ------------------------------------------------------------------ backward
        for l in range(len(layers) - 1, -1, -1):
            layer = layers[l]
            X, XMID, QKV, P, GU = self.X[l], self.XMID[l], self.QKV[l], self.P[l], self.GU[l]
            cos, sin = e._layer_rope(l)
            k.gemm(self.DX, self.wT_down[l], out=self.DACT, M=L, N=dff, K=d)
            native.check(lib.bma_glu_bwd(GU.data_ptr(), 2 * dff, self.DACT.data_ptr(), dff, self.DGU.data_ptr(),
                                         2 * dff, L, dff, e.act, st), "glu_bwd")
            k.gemm(self.DGU, self.wT_gu[l], out=self.DXN, M=L, N=d, K=2 * dff)
            self._rmsnorm_bwd(self.DXN, XMID, layer.post_attention_layernorm.weight, self.DX, self.DXM, L)
            k.gemm(self.DXM, self.wT_o[l], out=self.DATT, M=L, N=qd, K=d)
            self._attn_backward(QKV, P, L, scale, causal=1)
            native.check(lib.bma_rope_bwd(self.DQKV.data_ptr(), self.qkv_dim, cos.data_ptr(), sin.data_ptr(), L, 2 * H,
                                          hd, 0, st), "rope_bwd")
            k.gemm(self.DQKV, self.wT_qkv[l], out=self.DXN, M=L, N=d, K=self.qkv_dim)
            self._rmsnorm_bwd(self.DXN, X, layer.input_layernorm.weight, self.DXM, self.DX, L)
        return self.DX

    @property
    def loss(self) -> torch.Tensor:
        """Mean CE of the last call (fp32, device)."""
        return self.LOSS.mean()


class NativeVisionTower(_AttentionProducts):
    """CLIP vision tower + LLaVA projector on this package's kernels: ``model.get_image_features`` as the
    reference calls it (bimodal_attack.py:526-536, ``vision_feature_layer=-2``, strategy "default") and
    its input gradient back to the raw fp32 pixels (the image half of compute_gradient, :1016-1025).
    The torchvision Normalize the reference applies first is folded into the im2col kernel; the patch
    convolution (stride == kernel, no bias) is a GEMM and so is its input gradient."""

    @staticmethod
    def supported(model, normalize) -> Optional[str]:
        from .attack import find_vision_parts
        parts = find_vision_parts(model)
        if parts is None:
            return "no CLIP vision tower + LLaVA projector found"
        tower, proj = parts
        vc = tower.config
        if type(tower).__name__ != "CLIPVisionModel":
            return f"vision tower {type(tower).__name__}"
        if getattr(normalize, "mean", None) is None or getattr(normalize, "std", None) is None:
            return "normalize is not a torchvision Normalize (mean/std unknown)"
        if len(normalize.mean) != 3:
            return "3-channel images only"
        hd = vc.hidden_size // vc.num_attention_heads
        if hd not in (64, 128):
            return f"vision head_dim {hd}"
        if vc.hidden_act not in ("quick_gelu", "gelu"):
            return f"vision activation {vc.hidden_act}"
        if getattr(model.config, "projector_hidden_act", "gelu") != "gelu":
            return "projector activation"
        if not isinstance(model.config.vision_feature_layer, int):
            return "several vision feature layers"
        if model.config.vision_feature_select_strategy not in ("default", "full"):
            return "vision feature strategy"
        if vc.hidden_size % 8 or vc.intermediate_size % 8:
            return "vision widths must be multiples of 8"
        if next(tower.parameters()).dtype != torch.bfloat16:
            return "bf16 towers only"
        return None

    def __init__(self, model, normalize):
        from .attack import find_vision_parts
        why = self.supported(model, normalize)
        if why is not None:
            raise RuntimeError(f"bimodal_b200 native vision tower: {why} is not covered")
        tower, proj = find_vision_parts(model)
        vm = tower.vision_model
        vc = tower.config
        self.dev = next(tower.parameters()).device
        self.k = _Lib(self.dev)
        self.D, self.H = vc.hidden_size, vc.num_attention_heads
        self.hd = self.D // self.H
        self.q_dim, self.qkv_dim = self.D, 3 * self.D
        self.F = vc.intermediate_size
        self.eps = float(vc.layer_norm_eps)
        self.patch, self.px = vc.patch_size, vc.image_size
        self.grid = self.px // self.patch
        self.n_patch = self.grid * self.grid
        self.L = self.n_patch + 1
        self.Lp = (self.L + 7) // 8 * 8
        self.act_kind = 0 if vc.hidden_act == "quick_gelu" else 1
        n_all = vc.num_hidden_layers
        self.n_run = model.config.vision_feature_layer % (n_all + 1)     # hidden_states[i] = after i layers
        self.drop_cls = model.config.vision_feature_select_strategy == "default"
        self.scale = float(self.hd) ** -0.5
        self.mean = (C.c_float * 3)(*[float(torch.tensor(m, dtype=torch.float32)) for m in normalize.mean])
        self.std = (C.c_float * 3)(*[float(torch.tensor(v, dtype=torch.float32)) for v in normalize.std])
        bf = torch.bfloat16
        emb = vm.embeddings
        self.Kp = (3 * self.patch * self.patch + 7) // 8 * 8              # 588 -> 592: 16-byte rows for TMA
        with torch.no_grad():
            wp = torch.zeros(self.D, self.Kp, dtype=bf, device=self.dev)
            wp[:, :3 * self.patch * self.patch] = emb.patch_embedding.weight.reshape(self.D, -1)
            self.w_patch, self.wT_patch = wp, wp.t().contiguous()
            self.cls = emb.class_embedding.detach().reshape(1, self.D).contiguous()
            self.pos = emb.position_embedding.weight.detach().contiguous()
            self.pre_ln = vm.pre_layrnorm
            self.layers = list(vm.encoder.layers)[:self.n_run]
            self.w_qkv, self.b_qkv, self.wT_qkv, self.wT_o, self.wT_fc1, self.wT_fc2 = [], [], [], [], [], []
            for lyr in self.layers:
                a = lyr.self_attn
                w = torch.cat([a.q_proj.weight, a.k_proj.weight, a.v_proj.weight], 0).contiguous()
                self.w_qkv.append(w)
                self.b_qkv.append(torch.cat([a.q_proj.bias, a.k_proj.bias, a.v_proj.bias], 0).contiguous())
                self.wT_qkv.append(w.t().contiguous())
                self.wT_o.append(a.out_proj.weight.t().contiguous())
                self.wT_fc1.append(lyr.mlp.fc1.weight.t().contiguous())
                self.wT_fc2.append(lyr.mlp.fc2.weight.t().contiguous())
            self.proj = proj
            self.d_out = proj.linear_2.weight.shape[0]
            self.wT_p1 = proj.linear_1.weight.t().contiguous()
            self.wT_p2 = proj.linear_2.weight.t().contiguous()
        if proj.linear_1.bias is None or self.layers and self.layers[0].self_attn.q_proj.bias is None:
            raise RuntimeError("bimodal_b200 native vision tower: bias-free projections are not covered")
        L, Lp, D, H, nl = self.L, self.Lp, self.D, self.H, self.n_run
        z = lambda *s, dt=bf: torch.zeros(s, dtype=dt, device=self.dev)     # noqa: E731
        self.PATCHES = z(self.n_patch, self.Kp)
        self.E = z(L, D)                      # embeddings before the pre-layernorm
        self.X = z(nl + 1, L, D)
        self.XMID = z(max(nl, 1), L, D)
        self.QKV = z(max(nl, 1), L, self.qkv_dim)
        self.P = z(max(nl, 1), H, L, Lp)
        self.H1 = z(max(nl, 1), L, self.F)
        self.XN = z(L, D)
        self.A1 = z(L, self.F)
        self._attn_buffers(L, z)
        n_feat = self.n_patch if self.drop_cls else L
        self.n_feat = n_feat
        self.Z1 = z(n_feat, self.d_out)
        self.G1 = z(n_feat, self.d_out)
        self.OUT = z(n_feat, self.d_out)
        self.DG = z(n_feat, self.d_out)
        self.DX = z(L, D)
        self.DXM = z(L, D)
        self.DXN = z(L, D)
        self.DA1 = z(L, self.F)
        self.DPATCH = z(self.n_patch, self.Kp)

    # -------------------------------------------------------------------------------------------
    def _ln(self, x, mod, y, rows):
        native.check(self.k.lib.bma_layernorm(x.data_ptr(), x.stride(0), mod.weight.data_ptr(), mod.bias.data_ptr(),
                                              y.data_ptr(), y.stride(0), rows, self.D, self.eps, self.k.st), "layernorm")

    def _ln_bwd(self, dy, x, mod, dres, dx, rows):
        native.check(self.k.lib.bma_layernorm_bwd(
            dy.data_ptr(), dy.stride(0), x.data_ptr(), x.stride(0), mod.weight.data_ptr(),
            None if dres is None else dres.data_ptr(), 0 if dres is None else dres.stride(0),
            dx.data_ptr(), dx.stride(0), rows, self.D, self.eps, self.k.st), "layernorm_bwd")

    def _act(self, x, dy, out, kind):
        native.check(self.k.lib.bma_act(x.data_ptr(), None if dy is None else dy.data_ptr(), out.data_ptr(),
                                        x.numel(), kind, self.k.st), "act")

    def forward(self, image: torch.Tensor, keep: bool = True) -> torch.Tensor:
        """image: [1, 3, px, px] fp32 raw pixels in [0, 1] -> image features [1, n_feat, d_llm] bf16 (a view of
        an internal buffer, valid until the next call). ``keep=False`` (scoring phase, no backward follows): the
        per-head products S = QK^T -> softmax -> PV, whose probabilities the backward needs, are replaced by ONE
        launch of the library's tcgen05 flash-attention kernel in its no-tail mode (bma_attention, S = 0)."""
        k, lib, st = self.k, self.k.lib, self.k.st
        L, Lp, D, F = self.L, self.Lp, self.D, self.F
        if tuple(image.shape[-3:]) != (3, self.px, self.px) or image.dtype != torch.float32:
            raise RuntimeError(f"bimodal_b200 native vision tower: expected a float32 3x{self.px}x{self.px} image")
        img = image.detach().contiguous()
        native.check(lib.bma_patchify(img.data_ptr(), self.PATCHES.data_ptr(), self.Kp, 3, self.px, self.px, self.patch,
                                      self.mean, self.std, st), "patchify")
        k.gemm(self.PATCHES, self.w_patch, out=self.E[1:], M=self.n_patch, N=D, K=self.Kp, ldc=D)
        self.E[0:1].copy_(self.cls)
        native.check(lib.bma_add_rows(self.E.data_ptr(), D, self.pos.data_ptr(), D, L, self.E.data_ptr(), D, L, D, st),
                     "add_rows")
        self._ln(self.E, self.pre_ln, self.X[0], L)
        for l, lyr in enumerate(self.layers):
            X, XMID, QKV, P, H1 = self.X[l], self.XMID[l], self.QKV[l], self.P[l], self.H1[l]
            a = lyr.self_attn
            self._ln(X, lyr.layer_norm1, self.XN, L)
            k.gemm(self.XN, self.w_qkv[l], out=QKV, M=L, N=self.qkv_dim, K=D, bias=self.b_qkv[l])
            self._transpose(QKV, self.qkv_dim, 0, self.VT, Lp, 0, L, D, 1, src_off=2 * D)
            if keep:
                self._attn_forward(QKV, P, L, self.scale, causal=0)
            else:
                kp = QKV.data_ptr() + 2 * D
                native.check(lib.bma_attention(QKV.data_ptr(), self.qkv_dim, kp, self.qkv_dim, self.VT.data_ptr(), Lp,
                                               kp, self.qkv_dim, self.VT.data_ptr(), Lp, self.ATT.data_ptr(), D,
                                               L, 0, L, self.H, 1, self.hd, self.scale, st), "attention(no tail)")
            k.gemm(self.ATT, a.out_proj.weight, out=XMID, M=L, N=D, K=D, bias=a.out_proj.bias, R=X, ldr=D)
            self._ln(XMID, lyr.layer_norm2, self.XN, L)
            k.gemm(self.XN, lyr.mlp.fc1.weight, out=H1, M=L, N=F, K=D, bias=lyr.mlp.fc1.bias)
            self._act(H1, None, self.A1, self.act_kind)
            k.gemm(self.A1, lyr.mlp.fc2.weight, out=self.X[l + 1], M=L, N=D, K=F, bias=lyr.mlp.fc2.bias, R=XMID, ldr=D)
        feat = self.X[self.n_run][1:] if self.drop_cls else self.X[self.n_run]
        pj = self.proj
        k.gemm(feat, pj.linear_1.weight, out=self.Z1, M=self.n_feat, N=self.d_out, K=D, lda=D, bias=pj.linear_1.bias)
        self._act(self.Z1, None, self.G1, 1)            # projector: erf-GELU
        k.gemm(self.G1, pj.linear_2.weight, out=self.OUT, M=self.n_feat, N=self.d_out, K=self.d_out, bias=pj.linear_2.bias)
        return self.OUT.unsqueeze(0)

    def backward(self, d_feats: torch.Tensor) -> torch.Tensor:
        """Gradient of the loss w.r.t. the raw image ([1, 3, px, px] fp32) given its gradient w.r.t. the
        features returned by the LAST forward() call."""
        k, lib, st = self.k, self.k.lib, self.k.st
        L, D, F = self.L, self.D, self.F
        dO = d_feats.reshape(self.n_feat, self.d_out).to(torch.bfloat16).contiguous()
        k.gemm(dO, self.wT_p2, out=self.DG, M=self.n_feat, N=self.d_out, K=self.d_out)
        self._act(self.Z1, self.DG, self.DG, 1)
        self.DX.zero_()
        dfeat = self.DX[1:] if self.drop_cls else self.DX
        k.gemm(self.DG, self.wT_p1, out=dfeat, M=self.n_feat, N=D, K=self.d_out, ldc=D)
        for l in range(self.n_run - 1, -1, -1):
            lyr = self.layers[l]
            X, XMID, QKV, P, H1 = self.X[l], self.XMID[l], self.QKV[l], self.P[l], self.H1[l]
            k.gemm(self.DX, self.wT_fc2[l], out=self.DA1, M=L, N=F, K=D)
            self._act(H1, self.DA1, self.DA1, self.act_kind)
            k.gemm(self.DA1, self.wT_fc1[l], out=self.DXN, M=L, N=D, K=F)
            self._ln_bwd(self.DXN, XMID, lyr.layer_norm2, self.DX, self.DXM, L)
            k.gemm(self.DXM, self.wT_o[l], out=self.DATT, M=L, N=D, K=D)
            self._attn_backward(QKV, P, L, self.scale, causal=0)
            k.gemm(self.DQKV, self.wT_qkv[l], out=self.DXN, M=L, N=D, K=self.qkv_dim)
            self._ln_bwd(self.DXN, X, lyr.layer_norm1, self.DXM, self.DX, L)
        self._ln_bwd(self.DX, self.E, self.pre_ln, None, self.DXM, L)
        k.gemm(self.DXM[1:], self.wT_patch, out=self.DPATCH, M=self.n_patch, N=self.Kp, K=D, lda=D)
        g = torch.empty((1, 3, self.px, self.px), dtype=torch.float32, device=self.dev)
        native.check(lib.bma_unpatchify_grad(self.DPATCH.data_ptr(), self.Kp, g.data_ptr(), 3, self.px, self.px,
                                             self.patch, self.std, st), "unpatchify_grad")
        return g
