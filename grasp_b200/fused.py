"""Decoder-layer execution of the calibration passes on the C-ABI kernels, without autograd.

The reference runs `self.model(...)` + `loss.backward()` per calibration sample
(reference modeling_grasp.py:340-354), i.e. transformers' eager LlamaDecoderLayer and the
autograd graph behind it.  On a B200 the GEMMs of that step take about half of its time; the
other half is ~25 elementwise launches per layer and direction, a split pre-pass per GEMM
operand and autograd's gradient-accumulation adds.  This module runs the same mathematics as
an explicit forward and an explicit backward:

  * every linear is `grasp_gemm_f16x3_planes` on prepared operands: a weight is split once
    (tensor-scaled, so the same planes serve x W^T and dy W) and cached until it changes;
    an activation is split once per consumer group (q/k/v share one, gate/up share one);
  * RMSNorm / rotary / SwiGLU / cross-entropy are one row kernel each (layer_ops.cu), their
    backward the exact derivative, residual-gradient adds fused into the RMSNorm backward,
    sums over q/k/v and gate/up gradients folded into the GEMM epilogue (beta = 1);
  * the backward stops where no GRASPLayer below needs a gradient (autograd prunes the same
    way), and GRASPLayers harvest G += dY^T X exactly as engine.SigmaLinearFn does.

Attention is grasp_attn_fwd / grasp_attn_bwd (tcgen05 flash kernels on the same fp16-plane arithmetic) for
head dimensions 64 and 128, torch's scaled_dot_product_attention otherwise, behind `sdpa_fwd/bwd`.
All arithmetic sits behind a small backend object so that the orchestration (which tensors
are saved, every backward formula) is testable on CPU against autograd with a torch backend
that lives in tests/ -- the product backend below has no CPU path.
"""
from __future__ import annotations

import os
import weakref
from collections import OrderedDict
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, ops


# ---------------------------------------------------------------------------------- backend
class CudaBackend:
    """The product arithmetic: libgrasp_b200.so through grasp_b200.ops."""

    def __init__(self, weight_cache_bytes: Optional[int] = None, cap_fn=None):
        self._w = OrderedDict()       # (data_ptr, shape) -> Operand (holds a reference to the weight)
        self._w_bytes = 0
        self._w_cap = weight_cache_bytes
        self._cap_fn = cap_fn         # device -> byte cap, asked once (the runner's memory plan)
        self._tags = {}

    def cached_bytes(self) -> int:
        return self._w_bytes

    # -- operands
    def prep(self, x):
        return ops.split_f16(x, _lib.SCALE_ROWS)

    def wprep(self, w, tag=None):
        key = (w.data_ptr(), tuple(w.shape))
        op = self._w.get(key)
        if op is not None and op.version == w._version:
            self._w.move_to_end(key)
            return op
        if op is not None:
            self._drop(key)
        if self._w_cap is None:
            if self._cap_fn is not None:
                self._w_cap = int(self._cap_fn(w.device))
            else:
                self._w_cap = int(0.35 * torch.cuda.get_device_properties(w.device).total_memory)
        op = ops.split_f16(w.detach(), _lib.SCALE_TENSOR, keep_src=True)
        self._tags[key] = tag
        self._w[key] = op
        self._w_bytes += op.nbytes()
        while self._w_bytes > self._w_cap and len(self._w) > 1:
            self._drop(next(iter(self._w)))
        return op

    def _drop(self, key):
        op = self._w.pop(key)
        self._tags.pop(key, None)
        self._w_bytes -= op.nbytes()

    def drop_weights(self, tag=None):
        """Forget cached planes: all of them, or those registered under `tag` (a layer index)."""
        for key in [k for k in self._w if tag is None or self._tags.get(k) == tag]:
            self._drop(key)

    # -- GEMMs
    def mm_nt(self, xo, wo, out=None, beta=0.0):          # x W^T
        return ops.gemm_planes(xo, wo, b_kn=False, beta=beta, C_out=out)

    def mm_nn(self, dyo, wo, out=None, beta=0.0):         # dy W
        return ops.gemm_planes(dyo, wo, b_kn=True, beta=beta, C_out=out)

    # the rank-k intermediate of a factor pair leaves its GEMM as the operand of the next one
    def mm_nt_op(self, xo, wo):
        return ops.gemm_planes_to_operand(xo, wo, b_kn=False)

    def mm_nn_op(self, dyo, wo):
        return ops.gemm_planes_to_operand(dyo, wo, b_kn=True)

    def harvest(self, dy, x, G):                          # G (+)= dy^T x
        if G is None:
            return ops.gemm(dy, x, ta=True)
        return ops.gemm(dy, x, ta=True, beta=1.0, C_out=G)

    # -- row kernels that hand their result over as a prepared operand (the fp32 tensor only on request)
    @staticmethod
    def rmsnorm_fwd_op(x, w, eps, keep_y):
        return ops.rmsnorm_fwd(x, w, eps, want_y=keep_y, want_operand=True)

    @staticmethod
    def swiglu_fwd_op(g, u, keep_h):
        if g.shape[1] > 51200:                              # row does not fit in shared memory: two kernels
            h = ops.swiglu_fwd(g, u)
            return h, ops.split_f16(h, _lib.SCALE_ROWS)
        return ops.swiglu_fwd(g, u, want_h=keep_h, want_operand=True)

    @staticmethod
    def swiglu_bwd_op(dh, g, u, keep_grads):
        if g.shape[1] > 25600:
            dg, du = ops.swiglu_bwd(dh, g, u, inplace=True)
            return dg, du, ops.split_f16(dg, _lib.SCALE_ROWS), ops.split_f16(du, _lib.SCALE_ROWS)
        return ops.swiglu_bwd(dh, g, u, inplace=True, want_grads=keep_grads, want_operands=True)

    attn_prep = staticmethod(ops.attn_prep_qkv)

    # -- row kernels
    rmsnorm_fwd = staticmethod(ops.rmsnorm_fwd)
    rmsnorm_bwd = staticmethod(ops.rmsnorm_bwd)
    rope_ = staticmethod(ops.rope_)
    swiglu_fwd = staticmethod(ops.swiglu_fwd)

    @staticmethod
    def swiglu_bwd(dh, g, u):
        return ops.swiglu_bwd(dh, g, u, inplace=True)

    ce_loss_bwd_ = staticmethod(ops.ce_loss_bwd_)


# ------------------------------------------------------------------------------ attention
def _sdpa(q, k, v, B, S, H, Hkv, D, scale):
    q4 = q.view(B, S, H, D).transpose(1, 2)
    k4 = k.view(B, S, Hkv, D).transpose(1, 2)
    v4 = v.view(B, S, Hkv, D).transpose(1, 2)
    if Hkv != H:                                           # transformers repeat_kv
        rep = H // Hkv
        k4 = k4[:, :, None].expand(B, Hkv, rep, S, D).reshape(B, H, S, D)
        v4 = v4[:, :, None].expand(B, Hkv, rep, S, D).reshape(B, H, S, D)
    o = F.scaled_dot_product_attention(q4, k4, v4, attn_mask=None, dropout_p=0.0, is_causal=S > 1, scale=scale)
    return o.transpose(1, 2).reshape(B * S, H * D)


def _own_attention(q, D) -> bool:
    """The library's tensor-core attention applies (fp32 CUDA activations, head_dim 64 / 128); otherwise torch's
    scaled_dot_product_attention (library code) serves, e.g. for the 16 / 32-wide heads of the test models."""
    return (q.is_cuda and q.dtype == torch.float32 and ops.attn_supported(D)
            and os.environ.get("GRASP_B200_ATTN", "1") != "0")       # 0: torch's kernels everywhere (cross-check)


def sdpa_fwd(q, k, v, B, S, H, Hkv, D, scale, keep):
    if _own_attention(q, D):
        out, ctx = ops.attn_fwd(q, k, v, B, S, H, Hkv, D, scale)
        return out, (("grasp", ctx) if keep else None)
    if not keep:
        return _sdpa(q, k, v, B, S, H, Hkv, D, scale), None
    ql, kl, vl = (t.detach().requires_grad_(True) for t in (q, k, v))
    with torch.enable_grad():
        out = _sdpa(ql, kl, vl, B, S, H, Hkv, D, scale)
    return out.detach(), (out, ql, kl, vl)


def sdpa_bwd(ctx, d_out):
    if ctx[0] == "grasp":
        return ops.attn_bwd(ctx[1], d_out)
    out, ql, kl, vl = ctx
    dq, dk, dv = torch.autograd.grad(out, (ql, kl, vl), d_out)
    return dq.contiguous(), dk.contiguous(), dv.contiguous()


# ------------------------------------------------------------------------------- linears
def linear_kind(mod) -> Optional[str]:
    if hasattr(mod, "dense_weight") and hasattr(mod, "Vh"):
        return "grasp"                                     # modeling_grasp.GRASPLayer
    if hasattr(mod, "InLinear") and hasattr(mod, "OutLinear"):
        ok = type(mod.InLinear) is nn.Linear and type(mod.OutLinear) is nn.Linear and mod.InLinear.bias is None
        return "svd" if ok else None                       # modeling_grasp.SVDLinear
    if type(mod) is nn.Linear:
        return "dense"
    return None


def _f32_param(t) -> bool:
    return t is not None and t.dtype == torch.float32


def layer_supported(layer) -> bool:
    att, mlp = getattr(layer, "self_attn", None), getattr(layer, "mlp", None)
    if att is None or mlp is None:
        return False
    names = [(att, n) for n in ("q_proj", "k_proj", "v_proj", "o_proj")] + [(mlp, n) for n in
                                                                           ("gate_proj", "up_proj", "down_proj")]
    for owner, n in names:
        mod = getattr(owner, n, None)
        if mod is None or linear_kind(mod) is None:
            return False
    for n in ("input_layernorm", "post_attention_layernorm"):
        norm = getattr(layer, n, None)
        if norm is None or not _f32_param(getattr(norm, "weight", None)) or not hasattr(norm, "variance_epsilon"):
            return False
        if type(norm).__name__ not in ("LlamaRMSNorm",):
            return False
    act = getattr(mlp, "act_fn", None)
    if type(act).__name__ not in ("SiLU", "SiLUActivation"):
        return False
    if not all(hasattr(att, a) for a in ("head_dim", "scaling", "num_key_value_groups")):
        return False
    if getattr(att, "sliding_window", None):
        return False
    return att.head_dim % 8 == 0


class FusedLlama:
    """Explicit forward / backward over `runner.layers` (engine.LlamaRunner owns the modules)."""

    def __init__(self, runner, backend=None):
        # the runner owns this object: a proxy avoids a reference cycle that would keep the cached weight
        # planes (as large as the model) alive until the cycle collector runs
        self.r = weakref.proxy(runner)
        proxy = self.r                 # (a bound method would re-create the reference cycle)
        self.be = backend if backend is not None else CudaBackend(cap_fn=lambda dev: proxy.plane_cap_bytes(dev))

    def supported(self) -> bool:
        r = self.r
        if type(r.norm).__name__ != "LlamaRMSNorm" or linear_kind(r.head) != "dense":
            return False
        return all(layer_supported(l) for l in r.layers)

    # ---- linears ---------------------------------------------------------------------
    def _weight(self, mod, kind):
        return mod.dense_weight() if kind == "grasp" else mod.weight

    def lin_fwd(self, mod, xo, tag=None):
        be, kind = self.be, linear_kind(mod)
        if kind == "svd":
            if hasattr(be, "mm_nt_op"):
                to = be.mm_nt_op(xo, be.wprep(mod.InLinear.weight, tag))
            else:
                to = be.prep(be.mm_nt(xo, be.wprep(mod.InLinear.weight, tag)))
            y = be.mm_nt(to, be.wprep(mod.OutLinear.weight, tag))
            bias = mod.OutLinear.bias
        else:
            y = be.mm_nt(xo, be.wprep(self._weight(mod, kind), tag))
            bias = mod.bias if kind == "dense" else None    # GRASPLayer ignores its bias (reference :77-79)
        if bias is not None:
            y += bias
        return y

    def lin_bwd(self, mod, dyo, out=None, beta=0.0, tag=None):
        be, kind = self.be, linear_kind(mod)
        if kind == "svd":
            if hasattr(be, "mm_nn_op"):
                dto = be.mm_nn_op(dyo, be.wprep(mod.OutLinear.weight, tag))
            else:
                dto = be.prep(be.mm_nn(dyo, be.wprep(mod.OutLinear.weight, tag)))
            return be.mm_nn(dto, be.wprep(mod.InLinear.weight, tag), out=out, beta=beta)
        return be.mm_nn(dyo, be.wprep(self._weight(mod, kind), tag), out=out, beta=beta)

    def harvest(self, mod, dy, x):
        mod._G = self.be.harvest(dy, x, mod._G)

    @staticmethod
    def _is_grasp(mod) -> bool:
        return linear_kind(mod) == "grasp" and mod.S.requires_grad

    # ---- one decoder layer -------------------------------------------------------------
    def layer_fwd(self, i, x, B, S, cos, sin, keep: bool):
        """x [B*S, d] -> (x_out [B*S, d], saved or None)"""
        be = self.be
        L = self.r.layers[i]
        att, mlp = L.self_attn, L.mlp
        D = att.head_dim
        g_attn_in = any(self._is_grasp(m) for m in (att.q_proj, att.k_proj, att.v_proj))
        g_mlp_in = any(self._is_grasp(m) for m in (mlp.gate_proj, mlp.up_proj))
        g_down = self._is_grasp(mlp.down_proj)
        # the norms and the gated activation feed linears only: they hand over the operand form directly and
        # write the fp32 tensor only where a GRASPLayer needs it for G += dY^T X
        xn, rstd1, xo = be.rmsnorm_fwd_op(x, L.input_layernorm.weight, L.input_layernorm.variance_epsilon,
                                          keep and g_attn_in)
        q = self.lin_fwd(att.q_proj, xo, tag=i)
        k = self.lin_fwd(att.k_proj, xo, tag=i)
        v = self.lin_fwd(att.v_proj, xo, tag=i)
        del xo
        H, Hkv = q.shape[1] // D, k.shape[1] // D
        if _own_attention(q, D) and hasattr(be, "attn_prep"):
            # RoPE + operand planes of q, k, v in two passes, then the tensor-core attention
            qo, ko, vo = be.attn_prep(q, k, v, S, H, Hkv, D, cos, sin)
            a, actx = ops.attn_fwd_prepared(qo, ko, vo, B, S, H, Hkv, D, att.scaling)
            actx = ("grasp", actx) if keep else None
            del qo, ko, vo
        else:
            be.rope_(q, S, H, D, cos, sin)
            be.rope_(k, S, Hkv, D, cos, sin)
            a, actx = sdpa_fwd(q, k, v, B, S, H, Hkv, D, att.scaling, keep)
        x2 = self.lin_fwd(att.o_proj, be.prep(a), tag=i)
        x2 += x
        xn2, rstd2, xo2 = be.rmsnorm_fwd_op(x2, L.post_attention_layernorm.weight,
                                            L.post_attention_layernorm.variance_epsilon, keep and g_mlp_in)
        g = self.lin_fwd(mlp.gate_proj, xo2, tag=i)
        u = self.lin_fwd(mlp.up_proj, xo2, tag=i)
        del xo2
        h, ho = be.swiglu_fwd_op(g, u, keep and g_down)
        x3 = self.lin_fwd(mlp.down_proj, ho, tag=i)
        del ho
        x3 += x2
        if not keep:
            return x3, None
        saved = {"x": x, "rstd1": rstd1, "xn": xn if g_attn_in else None, "actx": actx, "H": H, "Hkv": Hkv,
                 "a": a if self._is_grasp(att.o_proj) else None, "x2": x2, "rstd2": rstd2,
                 "xn2": xn2 if g_mlp_in else None, "g": g, "u": u, "h": h if g_down else None}
        return x3, saved

    def layer_bwd(self, i, sv, dx3, B, S, cos, sin, need_dx: bool):
        """Gradient of the layer input (None when nothing below needs it); harvests G of this layer's GRASPLayers."""
        be = self.be
        L = self.r.layers[i]
        att, mlp = L.self_attn, L.mlp
        D = att.head_dim
        gq, gk, gv, go = (self._is_grasp(m) for m in (att.q_proj, att.k_proj, att.v_proj, att.o_proj))
        gg, gu, gd = (self._is_grasp(m) for m in (mlp.gate_proj, mlp.up_proj, mlp.down_proj))
        need_x2 = need_dx or gq or gk or gv or go          # does a gradient have to reach the attention block?
        if not (need_x2 or gg or gu or gd):
            return None
        # ---- MLP block:  x3 = x2 + down(silu(gate(n2)) * up(n2)),  n2 = rmsnorm(x2)
        if gd:
            self.harvest(mlp.down_proj, dx3, sv["h"])
        dx2 = None
        if need_x2 or gg or gu:
            dh = self.lin_bwd(mlp.down_proj, be.prep(dx3), tag=i)
            if need_x2:          # dg / du go on into the gate / up backward: operand form straight from the kernel
                dg, du, dgo, duo = be.swiglu_bwd_op(dh, sv["g"], sv["u"], gg or gu)
            else:
                dg, du = be.swiglu_bwd(dh, sv["g"], sv["u"])
            del dh
            if gg:
                self.harvest(mlp.gate_proj, dg, sv["xn2"])
            if gu:
                self.harvest(mlp.up_proj, du, sv["xn2"])
            if need_x2:
                dxn2 = self.lin_bwd(mlp.gate_proj, dgo, tag=i)
                self.lin_bwd(mlp.up_proj, duo, out=dxn2, beta=1.0, tag=i)
                del dgo, duo
                dx2 = be.rmsnorm_bwd(dxn2, sv["x2"], L.post_attention_layernorm.weight, sv["rstd2"], add=dx3)
        if not need_x2:
            return None
        # ---- attention block:  x2 = x + o(sdpa(rope(q(n1)), rope(k(n1)), v(n1))),  n1 = rmsnorm(x)
        if go:
            self.harvest(att.o_proj, dx2, sv["a"])
        if not (need_dx or gq or gk or gv):
            return None
        da = self.lin_bwd(att.o_proj, be.prep(dx2), tag=i)
        dq, dk, dv = sdpa_bwd(sv["actx"], da)
        del da
        be.rope_(dq, S, sv["H"], D, cos, sin, inverse=True)
        be.rope_(dk, S, sv["Hkv"], D, cos, sin, inverse=True)
        if gq:
            self.harvest(att.q_proj, dq, sv["xn"])
        if gk:
            self.harvest(att.k_proj, dk, sv["xn"])
        if gv:
            self.harvest(att.v_proj, dv, sv["xn"])
        if not need_dx:
            return None
        dxn = self.lin_bwd(att.q_proj, be.prep(dq), tag=i)
        self.lin_bwd(att.k_proj, be.prep(dk), out=dxn, beta=1.0, tag=i)
        self.lin_bwd(att.v_proj, be.prep(dv), out=dxn, beta=1.0, tag=i)
        return be.rmsnorm_bwd(dxn, sv["x"], L.input_layernorm.weight, sv["rstd1"], add=dx2)

    # ---- final norm + head + loss --------------------------------------------------------
    def final_norm(self, x):
        return self.be.rmsnorm_fwd(x, self.r.norm.weight, self.r.norm.variance_epsilon)[0]

    def loss_and_grad(self, x, B, S, labels, weights):
        """sum_s w_s * mean_t CE(logits[s, t], labels[s, t+1]) and its gradient w.r.t. x [B*S, d]
        (transformers' causal-LM loss on the loader's already shifted labels, reference :347-350)."""
        be, r = self.be, self.r
        d = x.shape[1]
        xs = x.view(B, S, d)[:, :-1].reshape(B * (S - 1), d)          # the last position has no target
        _, rstd, xno = be.rmsnorm_fwd_op(xs, r.norm.weight, r.norm.variance_epsilon, False)
        wo = be.wprep(r.head.weight, "head")
        logits = be.mm_nt(xno, wo)
        del xno
        if r.head.bias is not None:
            logits += r.head.bias
        lab = labels[:, 1:]
        valid = lab >= 0                                              # ignore_index = -100
        coef = (weights.view(B, 1).float() / valid.sum(dim=1, keepdim=True).clamp(min=1).float()) * valid.float()
        loss_rows = be.ce_loss_bwd_(logits, lab.reshape(-1).contiguous(), coef.reshape(-1).contiguous())
        dxn = be.mm_nn(be.prep(logits), wo)
        del logits
        dxs = be.rmsnorm_bwd(dxn, xs, r.norm.weight, rstd)
        dx = torch.zeros(B, S, d, dtype=x.dtype, device=x.device)
        dx[:, :-1] = dxs.view(B, S - 1, d)
        return loss_rows.sum(), dx.view(B * S, d)

    # ---- whole passes --------------------------------------------------------------------
    def run_layers(self, hidden, lo, hi):
        """[B, S, d] -> [B, S, d] through layers [lo, hi), no gradient."""
        B, S, d = hidden.shape
        _, (cos, sin) = self.r._pos(hidden)
        x = hidden.reshape(B * S, d)
        for i in range(lo, hi):
            x, _ = self.layer_fwd(i, x, B, S, cos, sin, keep=False)
        return x.view(B, S, d)

    def forward_backward(self, hidden, labels, weights, start_layer, lowest_grasp):
        """One micro-batch of a sigma-gradient pass: forward from `start_layer`, loss, backward down to
        the lowest layer that holds a GRASPLayer.  Returns the loss (a device scalar)."""
        B, S, d = hidden.shape
        _, (cos, sin) = self.r._pos(hidden)
        x = hidden.reshape(B * S, d)
        saved = []
        n = self.r.n_layers
        for i in range(start_layer, n):
            x, sv = self.layer_fwd(i, x, B, S, cos, sin, keep=i >= lowest_grasp)
            saved.append(sv)
        loss, dx = self.loss_and_grad(x, B, S, labels, weights)
        del x
        for i in range(n - 1, lowest_grasp - 1, -1):
            sv = saved.pop()
            dx = self.layer_bwd(i, sv, dx, B, S, cos, sin, need_dx=i > lowest_grasp)
            del sv
        return loss
