"""Drop-in `modeling_grasp` for the B200-native GRASP hot path.

Same classes, method names, argument meaning and error behaviour as the reference's
modeling_grasp.py (SVDLinear :25-59, GRASPLayer :62-79, GRASPModel :82-469), so callers
and pickles written against the reference keep working, but every numeric stage runs in
the sm_100a kernels of grasp_b200 (no torch.linalg.svd / torch.topk / CPU path):

  compute_bi                 -> fused cosine reduction over the chain of hidden states
  replace_with_GRASPLayer    -> batched block-Jacobi SVD
  get_svdlayer_gradients     -> G = dL/dW harvested per matrix, one diag(U^T G V) contraction per block
  dynamic_svd_selection      -> |g*S| score + radix-select top-k
  compile_grasp_model        -> low-rank rebuild GEMM / sqrt(S) factor pack
"""
import logging
import os
from typing import Dict, List, Literal, Optional, Union

import numpy as np
import torch
import torch.nn as nn
from torch.utils.data import DataLoader
from tqdm import tqdm

from grasp_b200 import engine, ops
from tools.utils_func import adaptive_rank_selection, block_influence

logger = logging.getLogger(__name__)


def setup_logger(log_file=None):
    """(Re)attach exactly one handler: a file handler when log_file is given, stderr otherwise."""
    for h in list(logger.handlers):
        logger.removeHandler(h)
    logger.setLevel(logging.INFO)
    handler = logging.FileHandler(log_file) if log_file else logging.StreamHandler()
    handler.setFormatter(logging.Formatter('%(asctime)s - %(name)s - %(levelname)s - %(message)s'))
    logger.addHandler(handler)


class SVDLinear(nn.Module):
    """Two skinny linears replacing one dense linear after truncation.

    U [out, k], S [k], Vh [k, in] are the retained triplets (reference modeling_grasp.py:36-38);
    sigma_fuse="UV" splits sqrt(S) between both factors, "U"/"V" put S on one side.
    As in the reference, "V" leaves OutLinear at its default initialisation.
    """

    def __init__(self, U: torch.Tensor, S: torch.Tensor, Vh: torch.Tensor, bias: Optional[torch.Tensor],
                 sigma_fuse: Literal["UV", "U", "V"] = "UV"):
        super().__init__()
        out_features, in_features, rank = U.shape[0], Vh.shape[1], S.shape[0]
        self.InLinear = nn.Linear(in_features, rank, bias=False, device=U.device, dtype=U.dtype)
        self.OutLinear = nn.Linear(rank, out_features, bias=bias is not None, device=U.device, dtype=U.dtype)
        if bias is not None:
            self.OutLinear.bias.data = bias
        if sigma_fuse == "UV":
            in_w, out_w = ops.factor_pack(U, S, Vh, torch.arange(rank, device=U.device))
            self.InLinear.weight.data = in_w
            self.OutLinear.weight.data = out_w
        elif sigma_fuse == "U":
            self.InLinear.weight.data = Vh.contiguous()
            self.OutLinear.weight.data = (U * S).contiguous()
        elif sigma_fuse == "V":
            self.InLinear.weight.data = (Vh * S.view(-1, 1)).contiguous()
        else:
            raise ValueError(f"value of sigma_fuse {sigma_fuse} not support")

    @classmethod
    def from_packed(cls, in_w: torch.Tensor, out_w: torch.Tensor, bias: Optional[torch.Tensor]):
        """Build from factors already scaled by the pack kernel (no extra copies)."""
        self = cls.__new__(cls)
        nn.Module.__init__(self)
        rank, in_features = in_w.shape
        out_features = out_w.shape[0]
        self.InLinear = nn.Linear(in_features, rank, bias=False, device="meta")
        self.OutLinear = nn.Linear(rank, out_features, bias=bias is not None, device="meta")
        self.InLinear.weight = nn.Parameter(in_w, requires_grad=False)
        self.OutLinear.weight = nn.Parameter(out_w, requires_grad=False)
        if bias is not None:
            self.OutLinear.bias = nn.Parameter(bias.detach(), requires_grad=False)
        return self

    def forward(self, x: torch.Tensor):
        return self.OutLinear(self.InLinear(x))


class GRASPLayer(nn.Module):
    """A linear layer held as its thin SVD with the singular values as the only trainable
    parameter (reference modeling_grasp.py:62-79).  The bias is stored but, as in the
    reference, not applied in forward.

    The reference re-materialises W = U diag(S) Vh on every forward and lets autograd push the
    loss through two dense r x r products.  Here forward is one GEMM against the dense weight
    and backward harvests G = dY^T X; dL/dS = diag(U^T G V) is contracted once per matrix by
    the sigma-score kernel (immediately, or once per calibration pass when the engine defers it).
    """

    def __init__(self, U: torch.Tensor, S: torch.Tensor, Vh: torch.Tensor, bias: Optional[torch.Tensor],
                 compression_ratio: Optional[float], weight: Optional[torch.Tensor] = None):
        super().__init__()
        self.U = nn.Parameter(U.detach(), requires_grad=False)
        self.S = nn.Parameter(S.detach().clone(), requires_grad=True)
        self.Vh = nn.Parameter(Vh.detach(), requires_grad=False)
        self.in_features = self.Vh.shape[1]
        self.out_features = self.U.shape[0]
        self.bias = bias
        self.compression_ratio = compression_ratio
        # dense weight used by forward; equals U diag(S) Vh up to the SVD's fp32 backward error
        self._dense = weight.detach() if weight is not None else None
        self._G = None          # accumulated dL/dW while the engine defers the contraction
        self._defer = False

    def dense_weight(self) -> torch.Tensor:
        if self._dense is None:
            k = self.S.shape[0]
            self._dense = ops.lowrank_rebuild(self.U.data, self.S.data, self.Vh.data,
                                              torch.arange(k, device=self.S.device))
        return self._dense

    def forward(self, x: torch.Tensor):
        b, s, _ = x.shape
        y = engine.SigmaLinearFn.apply(x.reshape(b * s, -1), self.S, self)
        return y.view(b, s, -1)


class GRASPModel(nn.Module):
    def __init__(self, model: nn.Module, *args, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self.model = model
        for p in self.model.parameters():
            p.requires_grad = False
        self.grasp_values_dict = {}
        self._svd_cache: Dict[str, tuple] = {}
        # B200 engine state (not part of the reference surface)
        self.use_engine = os.environ.get("GRASP_B200_ENGINE", "1") != "0"
        self.micro_batch = int(os.environ.get("GRASP_B200_MICRO_BATCH", "0"))   # 0 = pick per calibration set
        self._runner = None
        self._calib = None

    def __getstate__(self):
        # caches and the runner hold device tensors / module references that do not belong in a checkpoint
        state = self.__dict__.copy()
        state["_runner"], state["_calib"], state["_svd_cache"] = None, None, {}
        return state

    # ------------------------------------------------------------------ B200 engine plumbing
    def _engine_runner(self):
        """Layer-wise runner with the prefix-activation cache, or None when the wrapped model is not a
        LLaMA-family causal LM (then the generic whole-model path is used)."""
        if not self.use_engine or not engine.LlamaRunner.supports(self.model):
            return None
        if self._runner is None:
            self._runner = engine.LlamaRunner(self.model, micro_batch=self.micro_batch or 8)
        return self._runner

    def _calibration_set(self, dataloader, device):
        if self._calib is None or self._calib[0] is not dataloader:
            self._calib = (dataloader, engine.CalibrationSet(dataloader, device))
            calib = self._calib[1]
            if self._runner is not None and calib.supported:
                hidden = getattr(self.model.config, "hidden_size", 4096)
                self._runner.micro_batch = self.micro_batch or engine.auto_micro_batch(calib.input_ids.shape[1], hidden)
        return self._calib[1]

    def _touched(self, module_name: str):
        idx = engine.layer_index(module_name)
        if self._runner is not None and idx is not None:
            self._runner.invalidate_above(idx)

    def prepare_calibration(self, calibration_dataloader, layers_id, device="cuda"):
        """Cache, in one forward sweep, the hidden state entering every layer of `layers_id` for all
        calibration samples, so each later block pass starts at its own layer."""
        runner = self._engine_runner()
        if runner is None:
            return False
        calib = self._calibration_set(calibration_dataloader, device)
        if not calib.supported:
            return False
        runner.build_cache(calib, list(layers_id), keep_only=True)
        return True

    # ------------------------------------------------------------------ misc (API parity)
    def calculate_layer_compression_ratio(self, redundant_layers: Optional[List] = None):
        """Allocation-aware ratios are a stub in the reference (modeling_grasp.py:91-112); kept as a no-op."""
        return None

    @staticmethod
    def _extract_layer_index(module_name):
        parts = module_name.split('.')
        try:
            if "layers" in parts:
                return int(parts[parts.index("layers") + 1])
        except (ValueError, IndexError):
            return None

    def print_trainable_params(self, log_file: Optional[str] = None):
        setup_logger(log_file=log_file)
        total = sum(p.numel() for p in self.parameters())
        trainable = sum(p.numel() for p in self.parameters() if p.requires_grad)
        logger.info(f"trainable params: {trainable} || all params: {total} || trainable: {trainable / total * 100:.2f}%")

    def _set_module(self, model, submodule_key, module):
        *parents, leaf = submodule_key.split('.')
        owner = model
        for name in parents:
            owner = getattr(owner, name)
        setattr(owner, leaf, module)

    # ------------------------------------------------------------------ stage 1: layer scoring
    def compute_bi(self, num_prune_layers: Optional[int] = 1, calibration_dataloader: Optional[DataLoader] = None,
                   hiddens: Optional[List[torch.Tensor]] = None, angular: bool = False,
                   device: Literal["cpu", "cuda"] = "cuda", log_file: Optional[str] = None, *args, **kwargs):
        """Sum over batches of the per-batch mean block influence of every layer; returns
        (layer_importances, ids of the num_prune_layers least influential layers)."""
        setup_logger(log_file=log_file)
        assert hiddens is not None or calibration_dataloader is not None, \
            "please provide hidden_states or calibration dataloader to compute block influence"
        n_layers = len(self.model.model.layers)
        logger.info("=======>Compute Block Influence")
        scorer = engine.BlockInfluence(n_layers, angular=angular, stride=num_prune_layers if angular else 1)
        runner = self._engine_runner() if (hiddens is None and not angular) else None
        calib = self._calibration_set(calibration_dataloader, device) if runner is not None else None
        if hiddens is not None:
            scorer.add(hiddens)
        elif runner is not None and calib.supported:
            runner.block_influence(calib, scorer)
        else:
            for batch in tqdm(calibration_dataloader, desc="Compute BI", total=len(calibration_dataloader), leave=True):
                attention_mask = None if len(batch) == 2 else batch["attention_mask"].to(device=device)
                input_ids = batch["input_ids"].to(device=device)
                with torch.no_grad():
                    outputs = self.model(input_ids=input_ids, attention_mask=attention_mask, use_cache=False,
                                         output_hidden_states=True, return_dict=True)
                scorer.add(outputs.hidden_states)
        self.layer_importances = scorer.result()

        scores = np.array(self.layer_importances)
        if angular:
            start_layer = int(np.argsort(scores[:len(scores) - num_prune_layers + 1])[0])
            layers_to_remove = list(range(start_layer, start_layer + num_prune_layers))
        else:
            layers_to_remove = np.argsort(scores)[:num_prune_layers].tolist()
        self.redundant_layers = layers_to_remove
        return self.layer_importances, layers_to_remove

    def remove_layers(self, layers_to_remove: Optional[List[int]] = [], angular: Optional[bool] = False,
                      num_prune_layers: Optional[int] = None):
        if not layers_to_remove:
            assert self.layer_importances, "Need to compute importances with self.compute_bi()"
            assert num_prune_layers, "Need number of layers to prune"
            scores = np.array(self.layer_importances)
            if angular:
                start_layer = int(np.argsort(scores[:len(scores) - num_prune_layers + 1])[0])
                layers_to_remove = list(range(start_layer, start_layer + num_prune_layers))
            else:
                layers_to_remove = np.argsort(scores)[:num_prune_layers].tolist()
        if layers_to_remove is None:
            raise NotImplementedError("lack layers_to_remove")
        if self._runner is not None:
            self._runner.invalidate_all()      # cached layer inputs / weight planes are keyed by layer position
        self._calib = None
        for layer_idx in sorted(layers_to_remove, reverse=True):
            try:
                del self.model.model.layers[layer_idx]
            except IndexError:
                logger.info(f"layer {layer_idx} does not exist, function may have already been called")
                return []
        return layers_to_remove

    # ------------------------------------------------------------------ stage 2: SVD swap-in
    def precompute_svd(self, target_layers: List[str], device: str = "cuda"):
        """Factor many linears in batched launches ahead of the block loop (the SVDs use the
        ORIGINAL weights, so they do not depend on the order of the gradient passes)."""
        weights, names = [], []
        for name in target_layers:
            module = self.model.get_submodule(name)
            if not isinstance(module, nn.Linear):
                raise TypeError(f"target layer should be of Linear module, but got {type(module)}")
            if name not in self._svd_cache:
                names.append(name)
                weights.append(module.weight.data.to(device=device))
        for name, usv in zip(names, engine.batched_svd(weights)):
            self._svd_cache[name] = usv

    def svd_hoist_layers(self, layers_id: List[int], names_of, device="cuda") -> int:
        """How many of the leading layers of `layers_id` to factor in one batched call: their factors
        (U, S, Vh fp32) must fit in a quarter of what is free on the device; at least one layer.
        names_of(layer_id) -> target module names of that layer."""
        if not layers_id:
            return 0
        runner = self._engine_runner()
        dev = torch.device(device)
        if runner is None or dev.type != "cuda":
            return len(layers_id)
        avail = runner.available_bytes(dev)
        used, count = 0, 0
        for layer_id in layers_id:
            need = 0
            for name in names_of(layer_id):
                w = self.model.get_submodule(name).weight
                o, i = w.shape
                r = min(o, i)
                need += 4 * (o * r + r + r * i)
            if count and used + need > 0.25 * avail:
                break
            used += need
            count += 1
        return engine.dist.all_min_int(count, dev)       # every rank must hoist the same matrices

    def replace_with_GRASPLayer(self, target_layer: str, device: Literal["cuda", "cpu"] = "cuda",
                                log_file: Optional[str] = None):
        setup_logger(log_file=log_file)
        module = self.model.get_submodule(target=target_layer)
        if not isinstance(module, nn.Linear):
            raise TypeError(f"target layer should be of Linear module, but got {type(module)}")
        w = module.weight.data.to(device=device)
        if target_layer in self._svd_cache:
            U, S, Vh = self._svd_cache.pop(target_layer)
        else:
            U, S, Vh = ops.svd(w)
        grasp_layer = GRASPLayer(U=U, S=S, Vh=Vh, bias=module.bias,
                                 compression_ratio=getattr(module, "compression_ratio", None), weight=w)
        self._set_module(self.model, target_layer, grasp_layer)
        self._touched(target_layer)

    _BLOCKS = {
        "attention": ("self_attn.", ["q_proj", "k_proj", "v_proj", "o_proj"]),
        "mlp": ("mlp.", ["down_proj", "up_proj", "gate_proj"]),
    }

    def block_target_names(self, layer_id: int, block_type: str, target_layer_types) -> List[str]:
        if block_type not in self._BLOCKS:
            raise NotImplementedError(f"block type {block_type} not support")
        prefix, defaults = self._BLOCKS[block_type]
        if not target_layer_types:
            target_layer_types = defaults
        elif not all(t in defaults for t in target_layer_types):
            raise ValueError(f"values in target layer types is not valid, should be one of {defaults}")
        return [f"model.layers.{layer_id}.{prefix}{t}" for t in target_layer_types]

    def compress_block(self, layer_id: int, block_type: Literal["attention", "mlp"],
                       target_layer_types: Union[List[str], str] = ["q_proj", "k_proj", "v_proj", "o_proj",
                                                                    "down_proj", "up_proj", "gate_proj"],
                       device: Literal["cuda", "cpu"] = "cuda", allocation_aware: Optional[bool] = None,
                       verbose: bool = False, log_file: Optional[str] = None):
        """Swap the block's target linears for GRASPLayers. Returns True when there is nothing to
        compress (the caller then skips the gradient pass), None otherwise."""
        setup_logger(log_file=log_file)
        if layer_id is None:
            raise ValueError("Layer id should be given, but got None")
        if target_layer_types is None:
            return True
        names = self.block_target_names(layer_id, block_type, target_layer_types)
        if allocation_aware:
            ratios = []
            for name in names:
                module = self.model.get_submodule(name)
                if not isinstance(module, nn.Linear):
                    continue
                ratio = getattr(module, "compression_ratio", None)
                if isinstance(ratio, torch.Tensor):
                    ratio = ratio.cpu().item()
                ratios.append(ratio)
                if ratio != 0:
                    self.replace_with_GRASPLayer(target_layer=name, device=device)
            if np.all(np.array(ratios) == 0):
                return True
            return None
        todo = [n for n in names if n not in self._svd_cache and isinstance(self.model.get_submodule(n), nn.Linear)]
        if len(todo) > 1:
            self.precompute_svd(todo, device=device)  # one batched launch group for the block
        for name in names:
            self.replace_with_GRASPLayer(target_layer=name, device=device)
        return None

    def compute_preserve_rank(self, grasp_layer: GRASPLayer, compression_ratio: float):
        if compression_ratio is None:
            raise ValueError("Compression ratio should not be None")
        i, o = grasp_layer.in_features, grasp_layer.out_features
        return int(i * o * (1 - compression_ratio) / (i + o))

    def check_exists_grasp_layer(self, log_file: Optional[str] = None):
        setup_logger(log_file=log_file)
        names = [name for name, module in self.model.named_modules() if isinstance(module, GRASPLayer)]
        if not names:
            logger.info("GRASPLayer not found in current model, please use GRASPModel.replace_with_GRASPLayer first")
        return names

    # ------------------------------------------------------------------ stage 3a: sigma gradients
    def get_svdlayer_gradients(self, calibration_dataloader: DataLoader,
                               device: Literal["cuda:0", "cpu"] = "cuda:0", log_file: Optional[str] = None,
                               *args, **kwargs):
        """dL/dS of every GRASPLayer summed over the calibration batches (dict name -> [r] fp32)."""
        setup_logger(log_file=log_file)
        names = self.check_exists_grasp_layer()
        layers = {name: self.model.get_submodule(name) for name in names}
        self.model.to(device=device)
        runner = self._engine_runner() if names else None
        if runner is not None:
            calib = self._calibration_set(calibration_dataloader, device)
            starts = [engine.layer_index(n) for n in names]
            if calib.supported and all(i is not None for i in starts):
                grads = runner.sigma_gradients(calib, layers, min(starts))
                self.grasp_layer_grads = grads
                return grads
        with engine.deferred_sigma_grads(layers.values()):
            for batch in tqdm(calibration_dataloader, desc="Gradients Collection",
                              total=len(calibration_dataloader), leave=True):
                attention_mask = None if len(batch) == 2 else batch["attention_mask"].to(device=device)
                input_ids = batch["input_ids"].to(device=device)
                labels = batch["labels"].to(device=device)
                outputs = self.model(input_ids=input_ids, attention_mask=attention_mask, labels=labels,
                                     use_cache=False)
                loss = outputs[0]
                self.model.zero_grad()
                loss.backward()
            grads = {name: engine.contract_sigma_grad(layer) for name, layer in layers.items()}
        self.grasp_layer_grads = grads
        return grads

    # ------------------------------------------------------------------ stage 3b: selection
    def dynamic_svd_selection(self, grasp_layer_grads: dict, metric: Literal["gradient", "taylor"] = "taylor",
                              compression_ratio: Optional[float] = None, threshold_ratio: Optional[float] = None,
                              verbose: Optional[bool] = False, log_file: Optional[str] = None):
        setup_logger(log_file=log_file)
        if not grasp_layer_grads:
            raise ValueError("gradients of grasp_layer should be given, but got None")
        if metric not in ("gradient", "taylor"):
            raise RuntimeError(f"{metric} not support")

        scores, ks, order = {}, {}, []
        for name, grad in grasp_layer_grads.items():
            layer: GRASPLayer = self.model.get_submodule(name)
            scores[name] = ops.score_from_grad(grad, layer.S.data, metric)
            # a per-layer ratio overrides the argument and, as in the reference, sticks for later layers
            if layer.compression_ratio is not None:
                compression_ratio = layer.compression_ratio
            if compression_ratio is not None:
                ks[name] = self.compute_preserve_rank(layer, compression_ratio=compression_ratio)
            else:
                assert threshold_ratio, "Please provide Taylor threshold to select rank adaptively"
                ks[name] = None
            order.append(name)

        ratio_names = [n for n in order if ks[n] is not None]
        picked = dict(zip(ratio_names, ops.topk_batched([scores[n] for n in ratio_names],
                                                        [ks[n] for n in ratio_names])))
        indices_dict = {}
        for name in order:
            if ks[name] is not None:
                indices_dict[name] = picked[name]
            else:
                indices_dict[name] = adaptive_rank_selection(svd_importance_list=scores[name],
                                                             target_ratio=threshold_ratio)
            layer = self.model.get_submodule(name)
            self.grasp_values_dict[name] = {
                "svd_importance": torch.round(scores[name].cpu(), decimals=3).tolist(),
                "svd_value": torch.round(layer.S.data.cpu(), decimals=3).tolist(),
            }
        if verbose:
            logger.info("+" * 100)
            for name, indices in indices_dict.items():
                logger.info(f"{name}")
                shown = indices.detach().cpu().numpy().tolist() if torch.is_tensor(indices) else list(indices)
                logger.info(shown[:128])
            logger.info("+" * 100)
        self.indices_dict = indices_dict
        return indices_dict

    # ------------------------------------------------------------------ stage 3c: compile
    def compile_grasp_model(self, indices_dict: Optional[dict] = None, merge: Optional[bool] = False,
                            sigma_fuse: Literal["UV", "U", "V"] = "UV", device: Literal["cpu", "cuda"] = "cuda",
                            log_file: Optional[str] = None):
        setup_logger(log_file=log_file)
        if indices_dict is None:
            indices_dict = self.indices_dict
        for name, indices in indices_dict.items():
            layer: GRASPLayer = self.model.get_submodule(name)
            idx = torch.as_tensor(indices, dtype=torch.int64, device=layer.S.device)
            bias = layer.bias
            if merge:
                W = ops.lowrank_rebuild(layer.U.data, layer.S.data, layer.Vh.data, idx)
                linear = nn.Linear(layer.in_features, layer.out_features, bias=bias is not None, device="meta")
                linear.weight = nn.Parameter(W, requires_grad=False)
                if bias is not None:
                    linear.bias = bias
                linear.requires_grad_(False)
                self._set_module(self.model, name, linear)
            else:
                if sigma_fuse == "UV":
                    in_w, out_w = ops.factor_pack(layer.U.data, layer.S.data, layer.Vh.data, idx)
                    new = SVDLinear.from_packed(in_w, out_w, bias)
                else:
                    new = SVDLinear(U=layer.U.data[:, idx], S=layer.S.data[idx], Vh=layer.Vh.data[idx, :], bias=bias,
                                    sigma_fuse=sigma_fuse)
                new.requires_grad_(False)
                self._set_module(self.model, name, new)
            self._touched(name)
            del layer
        return
