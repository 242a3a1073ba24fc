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
             a_off=0, b_off=0, c_off=0):
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
        native.check(self.lib.bma_gemm_ex(C.byref(g), self.st), "gemm_ex")


class NativeGradient:
    """Forward + input-gradient backward of one [prefix | image | suffix | after | target] sequence."""

    @staticmethod
    def supported(engine: ScoringEngine) -> Optional[str]:
        """None when the decoder is covered, else the reason it is not."""
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
        self.VT = z(self.q_dim, Lp)
        self.S32 = z(H, L, Lp, dt=torch.float32)
        self.ATT = z(L, self.q_dim)
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
        self.DATT = z(L, self.q_dim)
        self.DATT_T = z(self.q_dim, Lp)
        self.QK_T = z(2 * self.q_dim, Lp)
        self.DS = z(H, L, Lp)
        self.DS_T = z(H, L, Lp)
        self.P_T = z(H, L, Lp)
        self.DQKV = z(L, self.qkv_dim)

    # -------------------------------------------------------------------------------------------
    def _rmsnorm(self, x, w, y, rows):
        native.check(self.k.lib.bma_rmsnorm(x.data_ptr(), x.stride(0), w.data_ptr(), y.data_ptr(), y.stride(0), rows,
                                            self.d, self.eps, 0, self.k.st), "rmsnorm")

    def _rmsnorm_bwd(self, dy, x, w, dres, dx, rows):
        native.check(self.k.lib.bma_rmsnorm_bwd(
            dy.data_ptr(), dy.stride(0), x.data_ptr(), x.stride(0), w.data_ptr(),
            None if dres is None else dres.data_ptr(), 0 if dres is None else dres.stride(0),
            dx.data_ptr(), dx.stride(0), rows, self.d, self.eps, 0, self.k.st), "rmsnorm_bwd")

    def _transpose(self, src, lds, src_batch, dst, ldd, dst_batch, rows, cols, batch, src_off=0):
        native.check(self.k.lib.bma_transpose_bf16(src.data_ptr() + 2 * src_off, lds, src_batch, dst.data_ptr(), ldd,
                                                   dst_batch, rows, cols, batch, self.k.st), "transpose")

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
            # S_h = Q_h K_h^T (fp32), heads side by side along the columns of QKV
            k.gemm(QKV, QKV, out32=self.S32, M=L, N=L, K=hd, lda=self.qkv_dim, ldb=self.qkv_dim, ldc=Lp, tile=(128, 1),
                   batch=H, a_col=hd, b_col=hd, b_off=qd, c_batch=L * Lp)
            native.check(lib.bma_softmax_rows(self.S32.data_ptr(), Lp, P.data_ptr(), Lp, H, L, scale, 1, st), "softmax")
            # O_h = P_h V_h: P stacked along the rows, V^T stacked along the rows
            k.gemm(P, self.VT, out=self.ATT, M=L, N=hd, K=L, lda=Lp, ldb=Lp, ldc=qd, tile=(128, 1),
                   batch=H, a_row=L, b_row=hd, c_batch=hd)
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
            # dP_h = dO_h V_h^T (fp32)
            k.gemm(self.DATT, QKV, out32=self.S32, M=L, N=L, K=hd, lda=qd, ldb=self.qkv_dim, ldc=Lp, tile=(128, 1),
                   batch=H, a_col=hd, b_col=hd, b_off=2 * qd, c_batch=L * Lp)
            native.check(lib.bma_softmax_bwd_rows(P.data_ptr(), Lp, self.S32.data_ptr(), Lp, self.DS.data_ptr(), Lp,
                                                  H, L, scale, 1, st), "softmax_bwd")
            # K-major operands of the three products: Q^T | K^T, dO^T, dS^T, P^T
            self._transpose(QKV, self.qkv_dim, 0, self.QK_T, Lp, 0, L, 2 * qd, 1)
            self._transpose(self.DATT, qd, 0, self.DATT_T, Lp, 0, L, qd, 1)
            self._transpose(self.DS, Lp, L * Lp, self.DS_T, Lp, L * Lp, L, L, H)
            self._transpose(P, Lp, L * Lp, self.P_T, Lp, L * Lp, L, L, H)
            # dQ_h = dS_h K_h ; dK_h = dS_h^T Q_h ; dV_h = P_h^T dO_h  -> the Q | K | V columns of DQKV
            k.gemm(self.DS, self.QK_T, out=self.DQKV, M=L, N=hd, K=L, lda=Lp, ldb=Lp, ldc=self.qkv_dim, tile=(128, 1),
                   batch=H, a_row=L, b_row=hd, b_off=qd * Lp, c_batch=hd)
            k.gemm(self.DS_T, self.QK_T, out=self.DQKV, M=L, N=hd, K=L, lda=Lp, ldb=Lp, ldc=self.qkv_dim, tile=(128, 1),
                   batch=H, a_row=L, b_row=hd, c_batch=hd, c_off=qd)
            k.gemm(self.P_T, self.DATT_T, out=self.DQKV, M=L, N=hd, K=L, lda=Lp, ldb=Lp, ldc=self.qkv_dim, tile=(128, 1),
                   batch=H, a_row=L, b_row=hd, c_batch=hd, c_off=2 * qd)
            native.check(lib.bma_rope_bwd(self.DQKV.data_ptr(), self.qkv_dim, cos.data_ptr(), sin.data_ptr(), L, 2 * H,
                                          hd, 0, st), "rope_bwd")
            k.gemm(self.DQKV, self.wT_qkv[l], out=self.DXN, M=L, N=d, K=self.qkv_dim)
            self._rmsnorm_bwd(self.DXN, X, layer.input_layernorm.weight, self.DXM, self.DX, L)
        return self.DX

    @property
    def loss(self) -> torch.Tensor:
        """Mean CE of the last call (fp32, device)."""
        return self.LOSS.mean()
