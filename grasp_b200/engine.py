"""Stage orchestration behind modeling_grasp.GRASPModel.

Holds the pieces that have no counterpart in the reference because the reference leaves
them to autograd / per-call library dispatch:

  BlockInfluence        device-side accumulator for compute_bi (one launch per forward pass,
                        one D2H copy per calibration run instead of one per layer per batch)
  batched_svd           groups same-shape weights into batched SVD launches
  SigmaLinearFn         forward/backward of a GRASPLayer against the dense weight; harvests
                        G = dY^T X instead of differentiating through U diag(S) Vh
  deferred_sigma_grads  accumulate G over a calibration pass, contract diag(U^T G V) once
"""
from __future__ import annotations

import contextlib
import logging
import os
from collections import OrderedDict
from typing import Iterable, List, Sequence

import torch

from . import dist, ops

logger = logging.getLogger(__name__)

# The three GEMMs of a GRASPLayer used on its own (forward, dx, G = dY^T X):
# "auto"  (default) grasp_gemm_f32 (tcgen05, fp16-plane arithmetic) for CUDA tensors; torch.matmul only for CPU
#         tensors, which occur in the host-logic tests alone
# "grasp" always grasp_gemm_f32;  "torch" always torch.matmul (cuBLAS fp32 on CUDA: the cross-check of the tests)
_LINEAR_BACKEND = "auto"


def set_linear_backend(name: str) -> None:
    global _LINEAR_BACKEND
    if name not in ("auto", "torch", "grasp"):
        raise ValueError(name)
    _LINEAR_BACKEND = name


def linear_backend() -> str:
    return _LINEAR_BACKEND


def _mm(a: torch.Tensor, b: torch.Tensor, ta=False, tb=False, out=None, accumulate=False) -> torch.Tensor:
    if _LINEAR_BACKEND == "grasp" or (_LINEAR_BACKEND == "auto" and a.is_cuda and a.dtype == torch.float32):
        return ops.gemm(a, b, ta=ta, tb=tb, beta=1.0 if accumulate else 0.0, C_out=out)
    A = a.t() if ta else a
    B = b.t() if tb else b
    if out is None:
        return A @ B
    if accumulate:
        return out.addmm_(A, B)
    return torch.mm(A, B, out=out)


class BlockInfluence:
    """Accumulates sum over batches of mean-over-tokens block influence per layer
    (reference modeling_grasp.py:148-167) in a device-resident float64 vector."""

    def __init__(self, n_layers: int, angular: bool = False, stride: int = 1):
        self.n_layers = n_layers
        self.angular = bool(angular)
        self.stride = max(int(stride or 1), 1) if angular else 1
        self.acc = None

    def add(self, hiddens: Sequence[torch.Tensor], scale: float = 1.0) -> None:
        hiddens = list(hiddens)
        if self.acc is None:
            n = max(self.n_layers, len(hiddens) - 1)
            self.acc = torch.zeros(n, dtype=torch.float64, device=hiddens[0].device)
        if not self.angular:
            ops.bi_chain(hiddens, self.acc, scale=scale)
            return
        # angular variant: last token only, layer i against layer i+stride
        for i in range(len(hiddens) - self.stride):
            ops.bi_accumulate(hiddens[i][:, -1:], hiddens[i + self.stride][:, -1:], self.acc[i:i + 1], angular=True,
                              scale=scale)

    def result(self) -> List[float]:
        if self.acc is None:
            return [0.0] * self.n_layers
        dist.all_reduce_sum_(self.acc)      # multi-GPU: ranks scored disjoint sample shards
        return self.acc[: self.n_layers].cpu().tolist()


def batched_svd(weights: Sequence[torch.Tensor], max_group: int = 8):
    """SVD of every weight; same-shape matrices share launches (latency-bound eigen-solves overlap).
    Multi-GPU: matrices are split over the ranks and the factors broadcast from their owners."""
    rank, world = dist.rank_world()
    if world > 1 and len(weights) > 0:
        shapes = [tuple(w.shape) for w in weights]
        owner = dist.owners_of(shapes, world)
        mine = [i for i in range(len(weights)) if owner[i] == rank]
        local = dict(zip(mine, _batched_svd_local([weights[i] for i in mine], max_group)))
        return dist.exchange_factors(local, shapes, owner, weights[0].device)
    return _batched_svd_local(weights, max_group)


class SvdNotConverged(RuntimeError):
    pass


def _batched_svd_local(weights: Sequence[torch.Tensor], max_group: int = 8):
    # one call per WORKING shape: a preconditioned 11008 x 4096 matrix runs its Jacobi phase on a 4096 x 4096 factor
    # and shares launches with the attention projections
    groups: "OrderedDict[tuple, list]" = OrderedDict()
    for i, w in enumerate(weights):
        groups.setdefault(dist.svd_working_shape(*w.shape), []).append(i)
    out = [None] * len(weights)
    infos = []
    for idxs in groups.values():
        # same ACTUAL shape side by side: a call whose matrices leave the tensor-core phase in the same sweep wastes
        # no sweeps on the early ones (mixed groups of 8: 292 ms per matrix, pure ones 268-289; profiles/r02_svd_batch_sizes.txt)
        idxs = sorted(idxs, key=lambda i: tuple(weights[i].shape))
        for s in range(0, len(idxs), max_group):
            chunk = idxs[s:s + max_group]
            usvs, info = ops.svd_batched([weights[i] for i in chunk], return_info=True)
            infos.append((chunk, info))
            for i, usv in zip(chunk, usvs):
                out[i] = usv
    # one synchronisation for the whole call: a Jacobi run that hit its sweep limit, or a CholeskyQR2
    # preconditioning that was not sound (cond^2 beyond fp32: ill-conditioned trained weights), would hand
    # non-orthogonal factors to scoring and compile.  Such a matrix is factored again as it is (no preconditioning),
    # then on the pure-fp32 path, and refused if that fails too.
    flags = torch.cat([info[:, 1] for _, info in infos]).cpu().tolist() if infos else []
    order = [i for chunk, _ in infos for i in chunk]
    for i, ok in zip(order, flags):
        if ok:
            continue
        shape = tuple(weights[i].shape)
        logger.warning("SVD of matrix %d %s: preconditioning unsound or sweep limit hit; factoring it as it is", i, shape)
        usvs, info = ops.svd_batched([weights[i]], return_info=True, precondition=False)
        if not int(info[0, 1].item()):
            logger.warning("SVD of matrix %d %s did not converge; retrying on the fp32 CUDA-core path", i, shape)
            usvs, info = ops.svd_batched([weights[i]], prec=ops.PREC_SIMT, max_sweeps=48, return_info=True)
            if not int(info[0, 1].item()):
                raise SvdNotConverged(f"SVD of a {shape} matrix did not converge")
        out[i] = usvs[0]
    return out


class SigmaLinearFn(torch.autograd.Function):
    """y = x W^T with W the dense weight of a GRASPLayer; the gradient w.r.t. the singular
    values is dS_i = u_i^T (dY^T X) v_i  (identical to autograd through U diag(S) Vh)."""

    @staticmethod
    def forward(ctx, x, S, layer):
        ctx.layer = layer
        ctx.save_for_backward(x)
        return _mm(x, layer.dense_weight(), tb=True)

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        layer = ctx.layer
        dy = dy.contiguous()
        dx = _mm(dy, layer.dense_weight()) if ctx.needs_input_grad[0] else None
        dS = None
        if ctx.needs_input_grad[1]:
            if layer._defer:
                if layer._G is None:
                    layer._G = _mm(dy, x, ta=True)
                else:
                    _mm(dy, x, ta=True, out=layer._G, accumulate=True)
            else:
                G = _mm(dy, x, ta=True)
                dS, _ = ops.sigma_score(layer.U.data, G, layer.Vh.data, None, metric="gradient", want_score=False)
        return dx, dS, None


@contextlib.contextmanager
def deferred_sigma_grads(layers: Iterable):
    layers = list(layers)
    for layer in layers:
        layer._defer = True
        layer._G = None
    try:
        yield
    finally:
        for layer in layers:
            layer._defer = False
            layer._G = None


def contract_sigma_grad(layer, dsigma: torch.Tensor = None) -> torch.Tensor:
    """dL/dS from the accumulated G of one calibration pass (zeros if the layer saw no batch)."""
    if layer._G is None:
        g = torch.zeros_like(layer.S.data)
        return g if dsigma is None else dsigma
    g, _ = ops.sigma_score(layer.U.data, layer._G, layer.Vh.data, None, metric="gradient", dsigma=dsigma,
                           want_score=False)
    return g


# =============================================================================
# Calibration engine: layer-wise LLaMA runner with a prefix-activation cache
# =============================================================================
# The reference runs the whole HF model forward+backward for every calibration sample of
# every block pass (reference modeling_grasp.py:340-354).  Layers below the block being
# compressed are untouched originals (grasp.py:75 walks layers deepest first), so their
# output is a constant of the run: it is computed once per selected layer and cached in HBM
# (n_samples x seq x hidden fp32, 4.3 GB at LLaMA-2-7B / 512 x 511).  A block pass then runs
# only layers [l, L) + norm + lm_head with autograd, in micro-batches of several samples
# (the summed per-sample mean losses give exactly the gradient sum of the reference's
# batch-size-1 loop).

import re
import torch.nn.functional as F

_F_LINEAR = F.linear


class _GemmLinearFn(torch.autograd.Function):
    """y = x W^T (+ b) for a frozen weight, on grasp_gemm_f32."""

    @staticmethod
    def forward(ctx, x2d, weight):
        ctx.weight = weight
        return ops.gemm(x2d, weight, tb=True)

    @staticmethod
    def backward(ctx, dy):
        if not ctx.needs_input_grad[0]:
            return None, None
        return ops.gemm(dy.contiguous(), ctx.weight), None


def _routed_linear(input, weight, bias=None):
    if (input.is_cuda and input.dtype == torch.float32 and weight.dtype == torch.float32
            and not weight.requires_grad and input.dim() >= 2 and input.shape[-1] == weight.shape[1]):
        x2d = input.reshape(-1, input.shape[-1])
        if not x2d.is_contiguous():
            x2d = x2d.contiguous()
        y = _GemmLinearFn.apply(x2d, weight)
        if bias is not None:
            y = y + bias
        return y.view(*input.shape[:-1], weight.shape[0])
    return _F_LINEAR(input, weight, bias)


@contextlib.contextmanager
def grasp_linear(enabled: bool = True):
    """Route every fp32 CUDA F.linear (and the GRASPLayer GEMMs) through grasp_gemm_f32."""
    if not enabled:
        yield
        return
    global _LINEAR_BACKEND
    prev_backend = _LINEAR_BACKEND
    F.linear = _routed_linear
    torch.nn.functional.linear = _routed_linear
    _LINEAR_BACKEND = "grasp"
    try:
        yield
    finally:
        F.linear = _F_LINEAR
        torch.nn.functional.linear = _F_LINEAR
        _LINEAR_BACKEND = prev_backend


class CalibrationSet:
    """All calibration batches of a DataLoader, resident on the device (tokens are tiny)."""

    def __init__(self, dataloader, device):
        ids, labels, weights = [], [], []
        self.supported = True
        for batch in dataloader:
            if len(batch) != 2:       # attention_mask present -> generic path
                self.supported = False
                break
            b = batch["input_ids"].shape[0]
            ids.append(batch["input_ids"])
            labels.append(batch["labels"])
            weights += [1.0 / b] * b  # the reference averages the loss over the whole batch
        if self.supported and ids and all(t.shape[1:] == ids[0].shape[1:] for t in ids):
            all_ids, all_labels = torch.cat(ids), torch.cat(labels)
            self.n_total = all_ids.shape[0]
            self.n_batches = len(ids)
            # multi-GPU: every rank keeps a contiguous shard of the samples (sums are all-reduced later)
            rank, world = dist.rank_world()
            lo, hi = dist.shard_range(self.n_total, rank, world)
            on_cuda = torch.device(device).type == "cuda"
            if on_cuda:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            self.input_ids = all_ids[lo:hi].to(device, non_blocking=True)
            self.labels = all_labels[lo:hi].to(device, non_blocking=True)
            self.weights = torch.tensor(weights[lo:hi], dtype=torch.float32).to(device, non_blocking=True)
            if on_cuda:
                e1.record()
                e1.synchronize()
                self.h2d_ms = e0.elapsed_time(e1)
        else:
            self.supported = False

    def __len__(self):
        return self.input_ids.shape[0]


def auto_micro_batch(seq_len: int, hidden: int, n_sms: int = 148, lo: int = 4, hi: int = 16) -> int:
    """Samples per micro-batch such that the most frequent GEMM of a block pass (tokens x hidden x hidden,
    128 x 256 tiles, persistent over n_sms CTAs) fills its last wave best; ties go to the larger batch."""
    best, best_eff = lo, 0.0
    n_tiles_n = -(-hidden // 256)
    for mb in range(lo, hi + 1):
        tiles = -(-(mb * seq_len) // 128) * n_tiles_n
        waves = -(-tiles // n_sms)
        eff = (mb * seq_len * hidden) / (waves * n_sms * 128 * 256)
        if eff >= best_eff - 1e-9:
            best, best_eff = mb, eff
    return best


class LlamaRunner:
    """Runs a HF LLaMA-family causal LM layer by layer (same modules, same math as model.forward)."""

    def __init__(self, hf_model, micro_batch: int = 8, use_grasp_gemm: bool = True, use_fused=None):
        self.hf = hf_model
        # explicit forward/backward on the row kernels + prepared-operand GEMMs (fused.py); the transformers
        # modules + autograd route below stays as the generic fallback and as the cross-check in the tests
        self.use_fused = (os.environ.get("GRASP_B200_FUSED", "1") != "0") if use_fused is None else bool(use_fused)
        self._fused = None
        m = hf_model.model
        self.embed, self.layers, self.norm, self.rotary, self.head = (m.embed_tokens, m.layers, m.norm, m.rotary_emb,
                                                                      hf_model.lm_head)
        self.micro_batch = micro_batch
        self.use_grasp_gemm = use_grasp_gemm
        # activation store: layer id -> [n_samples, S, d] input of that layer, valid while every layer below
        # is an untouched original (entries of the scoring pass are checkpoints, the others prefix caches)
        self.cache = {}
        self.cache_key = None    # id of the CalibrationSet the store belongs to
        self.store_key = None    # ... and of the one its byte budget was planned for
        self.store_budget = None
        self.store_per = 0
        gb = os.environ.get("GRASP_B200_MEM_BUDGET_GB")
        self.budget_bytes = int(float(gb) * 2**30) if gb else None
        mb = os.environ.get("GRASP_B200_STORE_MB")         # fixed size of the activation store (else planned)
        self.store_cap_bytes = int(float(mb) * 2**20) if mb else None

    # model families whose forward IS embed -> decoder layers (causal, no mask) -> norm -> head, verified
    # against transformers' own forward in tests/test_engine_cpu.py; anything else (embedding normalisers,
    # logit soft-capping, sliding-window / per-layer masks) takes the generic whole-model path
    SUPPORTED_MODEL_TYPES = ("llama",)

    @staticmethod
    def supports(hf_model) -> bool:
        m = getattr(hf_model, "model", None)
        ok = all(hasattr(m, a) for a in ("embed_tokens", "layers", "norm", "rotary_emb")) and hasattr(hf_model, "lm_head")
        cfg = getattr(hf_model, "config", None)
        impl = getattr(cfg, "_attn_implementation", "sdpa")
        if getattr(cfg, "model_type", None) not in LlamaRunner.SUPPORTED_MODEL_TYPES:
            return False
        if getattr(cfg, "sliding_window", None) or getattr(cfg, "layer_types", None) and \
                any(t != "full_attention" for t in cfg.layer_types):
            return False
        return bool(ok and impl in ("sdpa", None))

    @property
    def n_layers(self):
        return len(self.layers)

    def fused(self, like: torch.Tensor):
        """The fused executor when it applies to activations like `like` (fp32 on a CUDA device, every
        module of a known kind), else None."""
        if not (self.use_fused and self.use_grasp_gemm and like.is_cuda and like.dtype == torch.float32):
            return None
        if self._fused is None:
            from .fused import FusedLlama
            self._fused = FusedLlama(self)
        return self._fused if self._fused.supported() else None

    def _pos(self, hidden):
        position_ids = torch.arange(hidden.shape[1], device=hidden.device).unsqueeze(0)
        return position_ids, self.rotary(hidden, position_ids=position_ids)

    def _layer(self, i, hidden, position_ids, pos_emb):
        out = self.layers[i](hidden, attention_mask=None, position_ids=position_ids, position_embeddings=pos_emb,
                             use_cache=False)
        return out[0] if isinstance(out, tuple) else out

    def run_layers(self, hidden, lo, hi):
        position_ids, pos_emb = self._pos(hidden)
        for i in range(lo, hi):
            hidden = self._layer(i, hidden, position_ids, pos_emb)
        return hidden

    def hidden_states(self, input_ids, fused=None):
        """The L+1 states HF returns with output_hidden_states=True (last one after the final norm)."""
        hidden = self.embed(input_ids)
        position_ids, pos_emb = self._pos(hidden)
        states = [hidden]
        if fused is not None:
            B, S, d = hidden.shape
            x = hidden.reshape(B * S, d)
            for i in range(self.n_layers):
                x, _ = fused.layer_fwd(i, x, B, S, pos_emb[0], pos_emb[1], keep=False)
                states.append(x.view(B, S, d))
            states[-1] = fused.final_norm(x).view(B, S, d)
            return states
        for i in range(self.n_layers):
            hidden = self._layer(i, hidden, position_ids, pos_emb)
            states.append(hidden)
        states[-1] = self.norm(hidden)
        return states

    def loss_sum(self, hidden, labels, weights):
        """sum_samples w_s * mean_t CE(logits[s, t], labels[s, t+1]): HF's causal-LM loss on the loader's
        already shifted labels (the reference's double shift), one term per sample."""
        # only positions [0, S-1) enter the loss: drop the last one before the head instead of slicing the logits
        logits = self.head(self.norm(hidden[:, :-1]))
        B, S1, V = logits.shape
        lab = labels[:, 1:]
        per_tok = F.cross_entropy(logits.reshape(-1, V).float(), lab.reshape(-1), reduction="none", ignore_index=-100)
        valid = (lab != -100).sum(dim=1).clamp(min=1)          # HF averages over the non-ignored labels
        return (per_tok.view(B, S1).sum(dim=1) / valid * weights).sum()

    # ---- memory plan ---------------------------------------------------------------
    # Everything the runner keeps in HBM beyond the model is bounded by what is free when the plan is made:
    # the activation store (layer inputs of all calibration samples), the cached weight planes of the fused
    # executor and the per-micro-batch workspace of a pass.  GRASP_B200_MEM_BUDGET_GB caps the whole process
    # (torch-allocated bytes) below the physical memory -- used by the tests to force the bounded paths.
    def available_bytes(self, device) -> int:
        """Bytes this process may still allocate on `device` (free + torch's cached-but-unused blocks)."""
        device = torch.device(device)
        if device.type != "cuda":
            return 1 << 62
        free, _ = torch.cuda.mem_get_info(device)
        st = torch.cuda.memory_stats(device)
        allocated = st.get("allocated_bytes.all.current", 0)
        avail = free + st.get("reserved_bytes.all.current", 0) - allocated
        cap = self.budget_bytes
        if cap is not None:
            avail = min(avail, cap - allocated)
        return max(int(avail), 0)

    def _linear_weight_bytes(self) -> int:
        """fp32 bytes of every linear the fused executor keeps planes of (two fp16 planes = the same bytes)."""
        total = 0
        for p in list(self.layers.parameters()) + list(self.head.parameters()):
            if p.dim() == 2:
                total += p.numel() * 4
        return total

    def plan_store(self, calib: "CalibrationSet", hidden_shape, element_size: int) -> int:
        """Fix the byte budget of the activation store for this calibration set (once per set): half of what
        is free after the weight planes the passes will build, at least one full-size entry."""
        key = id(calib)
        if self.store_key == key and self.store_budget is not None:
            return self.store_budget
        per = len(calib) * hidden_shape[1] * hidden_shape[2] * element_size
        device = calib.input_ids.device
        avail = self.available_bytes(device)
        if device.type == "cuda" and self.use_fused and self.use_grasp_gemm:
            cached = self._fused.be.cached_bytes() if (self._fused is not None and hasattr(self._fused.be, "cached_bytes")) else 0
            avail -= max(min(self._linear_weight_bytes(), self.plane_cap_bytes(device)) - cached, 0)
        self.store_budget = max(int(0.5 * avail) if self.store_cap_bytes is None else self.store_cap_bytes, per)
        self.store_key, self.store_per = key, per
        return self.store_budget

    def plane_cap_bytes(self, device) -> int:
        """Upper bound of the cached weight planes: 35 % of the device, and never more than 40 % of the budget."""
        total = torch.cuda.get_device_properties(device).total_memory
        cap = int(0.35 * total)
        if self.budget_bytes is not None:
            cap = min(cap, int(0.4 * self.budget_bytes))
        return cap

    def _store_bytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self.cache.values())

    def pass_micro_batch(self, calib: "CalibrationSet", start_layer: int, hidden) -> int:
        """Micro-batch of a sigma-gradient pass from `start_layer`: the configured one, shrunk when the state
        the backward needs of layers [start_layer, L) would not fit in 70 % of what is free."""
        mb = self.micro_batch
        if not hidden.is_cuda:
            return mb
        cfg = self.hf.config
        d, ff = hidden.shape[-1], getattr(cfg, "intermediate_size", 4 * hidden.shape[-1])
        vocab = self.head.weight.shape[0]
        S = hidden.shape[1]
        per_sample = 4 * S * ((9 * d + 3 * ff) * (self.n_layers - start_layer) + 6 * ff + 2 * vocab + 4 * d)
        avail = self.available_bytes(hidden.device)
        fit = int(0.7 * avail // max(per_sample, 1))
        return max(1, min(mb, fit))

    # ---- prefix cache -------------------------------------------------------------
    def invalidate_above(self, layer_id: int):
        """Layer `layer_id` changed: cached inputs of deeper layers are stale."""
        for k in [k for k in self.cache if k > layer_id]:
            del self.cache[k]
        if self._fused is not None and hasattr(self._fused.be, "drop_weights"):
            self._fused.be.drop_weights(layer_id)          # its modules were replaced: cached planes are stale

    def invalidate_all(self):
        """The layer list itself changed (layers removed): nothing cached can be trusted."""
        self.cache, self.cache_key = {}, None
        self.store_key = self.store_budget = None
        if self._fused is not None and hasattr(self._fused.be, "drop_weights"):
            self._fused.be.drop_weights(None)

    def build_cache(self, calib: CalibrationSet, layer_ids, keep_only: bool = False):
        """Make the input of the layers in `layer_ids` resident for all calibration samples -- as many of them
        as the store budget holds (evenly spaced, always the lowest one); the others are recomputed from the
        nearest resident entry below when their pass starts (sigma_gradients calls this with one layer).
        Sources are whatever valid entries the store already has (checkpoints of the scoring pass, earlier
        entries), else the embeddings.  keep_only drops every entry that is not in `layer_ids` afterwards."""
        if self.cache_key != id(calib):
            self.cache, self.cache_key = {}, id(calib)
        wanted = sorted(set(layer_ids))
        need = [l for l in wanted if l not in self.cache]
        if need and keep_only:
            # make room first: of the entries that are not wanted only the sweep's starting point is of use
            sources = [c for c in self.cache if c <= need[0]]
            start = max(sources) if sources else None
            for k in [k for k in self.cache if k not in wanted and k != start]:
                del self.cache[k]
        if need:
            n = len(calib)
            probe = self.embed(calib.input_ids[:1])
            budget = self.plan_store(calib, probe.shape, probe.element_size())
            per = self.store_per
            slots = max(int((budget - self._store_bytes()) // max(per, 1)), 1)
            if len(need) > slots:
                # not all fit: keep slots-1 of them (one slot stays free for the entry recomputed per pass)
                m = max(slots - 1, 1)
                if m == 1:
                    need = [need[0]]
                else:
                    need = sorted({need[round(j * (len(need) - 1) / (m - 1))] for j in range(m)})
            sources = [c for c in self.cache if c <= need[0]]
            start = max(sources) if sources else None
            with torch.no_grad(), grasp_linear(self.use_grasp_gemm):
                fused = None
                for s in range(0, n, self.micro_batch):
                    ids = calib.input_ids[s:s + self.micro_batch]
                    if start is None:
                        hidden, first = self.embed(ids), 0
                    else:
                        hidden, first = self.cache[start][s:s + ids.shape[0]], start
                    position_ids, pos_emb = self._pos(hidden)
                    if s == 0:
                        fused = self.fused(hidden)
                    for i in range(first, need[-1] + 1):
                        if i in need:
                            if i not in self.cache:
                                self.cache[i] = torch.empty((n,) + tuple(hidden.shape[1:]), dtype=hidden.dtype,
                                                            device=hidden.device)
                            self.cache[i][s:s + ids.shape[0]] = hidden
                        if i < need[-1]:
                            if fused is not None:
                                B, S, d = hidden.shape
                                hidden = fused.layer_fwd(i, hidden.reshape(B * S, d), B, S, pos_emb[0], pos_emb[1],
                                                         keep=False)[0].view(B, S, d)
                            else:
                                hidden = self._layer(i, hidden, position_ids, pos_emb)
        if keep_only:
            for k in [k for k in self.cache if k not in wanted]:
                del self.cache[k]      # checkpoints of the scoring pass served their purpose

    def _plan_checkpoints(self, calib: CalibrationSet, hidden_shape, element_size):
        """Layers whose input is kept during the scoring pass.  The selected layers are unknown until scoring has
        finished, but block influence falls with depth (the reference's premise, and ShortGPT's finding): two
        thirds of the slots go to the deepest layers, one per layer, the rest is spread evenly below them.  A kept
        entry that turns out to be a selected layer IS its prefix cache (no recomputation); the others serve as
        starting points for the sweep.  All slots but one of the store budget are used (entries that are not
        needed are dropped before the sweep allocates anything)."""
        if not calib.input_ids.is_cuda:
            return []
        budget = self.plan_store(calib, hidden_shape, element_size)
        n_ck = int(min(self.n_layers - 1, budget // max(self.store_per, 1) - 1))
        if n_ck <= 0:
            return []
        top = min((2 * n_ck + 2) // 3, self.n_layers - 1)
        plan = list(range(self.n_layers - top, self.n_layers))
        rest, below = n_ck - top, self.n_layers - top          # `rest` more entries among layers [1, below)
        if rest > 0 and below > 1:
            stride = -(-below // (rest + 1))
            plan += list(range(stride, below, stride))[:rest]
        return sorted(set(plan))

    # ---- stage 1 ------------------------------------------------------------------
    def block_influence(self, calib: CalibrationSet, scorer: "BlockInfluence"):
        self.cache, self.cache_key = {}, id(calib)
        plan = None
        with torch.no_grad(), grasp_linear(self.use_grasp_gemm):
            fused = self.fused(self.embed(calib.input_ids[:1])) if len(calib) else None
            for s in range(0, len(calib), self.micro_batch):
                ids = calib.input_ids[s:s + self.micro_batch]
                states = self.hidden_states(ids, fused)
                if plan is None:
                    plan = self._plan_checkpoints(calib, states[0].shape, states[0].element_size())
                    for l in plan:
                        self.cache[l] = torch.empty((len(calib),) + tuple(states[0].shape[1:]), dtype=states[0].dtype,
                                                    device=states[0].device)
                for l in plan:
                    self.cache[l][s:s + ids.shape[0]] = states[l]     # states[l] is the input of layer l
                # the reference adds one mean per DataLoader batch: sum_s w_s * mean_t(sample s)
                w = calib.weights[s:s + ids.shape[0]]
                if bool((w == w[0]).all()):
                    scorer.add(states, scale=float(w[0]) * ids.shape[0])
                else:
                    for j in range(ids.shape[0]):
                        scorer.add([h[j:j + 1] for h in states], scale=float(w[j]))

    # ---- stage 3a -----------------------------------------------------------------
    def sigma_gradients(self, calib: CalibrationSet, layers: dict, start_layer: int):
        """dL/dS of the GRASPLayers in `layers` (name -> module), all of them at or above start_layer."""
        self.build_cache(calib, [start_layer])
        src = self.cache[start_layer]
        fused = self.fused(src)
        mb = self.pass_micro_batch(calib, start_layer, src)
        self.last_pass_micro_batch = mb
        if fused is not None:
            with deferred_sigma_grads(layers.values()), torch.no_grad():
                for s in range(0, len(calib), mb):
                    fused.forward_backward(src[s:s + mb], calib.labels[s:s + mb], calib.weights[s:s + mb],
                                           start_layer, start_layer)
                grads = {name: contract_sigma_grad(layer) for name, layer in layers.items()}
            dist.all_reduce_sum_many_(list(grads.values()))
            return grads
        with deferred_sigma_grads(layers.values()), grasp_linear(self.use_grasp_gemm):
            for s in range(0, len(calib), mb):
                hidden = self.run_layers(src[s:s + mb], start_layer, self.n_layers)
                loss = self.loss_sum(hidden, calib.labels[s:s + mb], calib.weights[s:s + mb])
                loss.backward()
            grads = {name: contract_sigma_grad(layer) for name, layer in layers.items()}
        # multi-GPU: each rank contracted the G of its own samples; dL/dS is linear in G
        dist.all_reduce_sum_many_(list(grads.values()))
        return grads


_LAYER_RE = re.compile(r"\.layers\.(\d+)\.")


def layer_index(name: str):
    m = _LAYER_RE.search("." + name)
    return int(m.group(1)) if m else None
