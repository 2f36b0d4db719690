"""Drop-in `grasp.py`: the GRASP compression driver on the B200-native hot path.

`main(...)` keeps the reference's signature and stage order (reference grasp.py:29-152):
layer scoring -> deepest layer first -> per layer MLP block then attention block, each
block = SVD swap-in -> sigma-gradient pass -> selection -> compile.  The CLI accepts every
flag of the reference (grasp.py:155-244) so scripts/params_script.sh drives it unchanged.
Recovery fine-tuning and lm-eval evaluation are outside the hot path: they are delegated to
the reference's own `alpaca_grasp` / `evaluate_grasp` modules when those are importable.
"""
import argparse
import logging
import os
from typing import List, Literal, Optional, Union

import torch
from torch.utils.data import DataLoader
from tqdm import tqdm

from modeling_grasp import GRASPModel

logger = logging.getLogger(__name__)


def setup_logger(log_file=None):
    logger.setLevel(logging.INFO)
    handler = logging.FileHandler(log_file) if log_file else logging.StreamHandler()
    handler.setFormatter(logging.Formatter('%(asctime)s - %(name)s - %(levelname)s - %(message)s'))
    logger.addHandler(handler)


def _load_model_and_tokenizer(model_name_or_path):
    from transformers import AutoModelForCausalLM, AutoTokenizer
    model = AutoModelForCausalLM.from_pretrained(model_name_or_path)
    tokenizer = AutoTokenizer.from_pretrained(model_name_or_path)
    tokenizer.pad_token = tokenizer.eos_token
    return model, tokenizer


def _save(grasp_model, save_path):
    try:
        torch.save(grasp_model, save_path)
    except (AttributeError, TypeError) as exc:  # transformers 5.x hooks are not picklable
        from grasp_b200 import checkpoint
        logger.warning("whole-module pickle failed (%s); writing a state-dict checkpoint instead", exc)
        checkpoint.save(grasp_model, save_path)


def compress(grasp_model: GRASPModel, calibration_dataloader: DataLoader, layers_id=None, num_prune_layers=None,
             mlp_target_layer_types=("down_proj", "up_proj", "gate_proj"),
             attn_target_layer_types=("q_proj", "k_proj", "v_proj", "o_proj"), metric="taylor",
             compression_ratio=None, threshold_ratio=None, device="cuda", angular=False, allocation_aware=False,
             merge=False, verbose=False, log_file=None, hoist_svd=True):
    """The compression loop of reference grasp.py:61-126 on an already constructed GRASPModel."""
    if layers_id is None:
        layers_importance, layers_id = grasp_model.compute_bi(num_prune_layers=num_prune_layers,
                                                              calibration_dataloader=calibration_dataloader,
                                                              angular=angular, device=device)
        logger.info("Layer importance measure by BI:\n%s", layers_importance)
    if isinstance(layers_id, int):
        layers_id = [layers_id]
    grasp_model.redundant_layers = layers_id

    if allocation_aware:
        logger.info("=======> Start Compression ratio allocation with GRASP")
        grasp_model.calculate_layer_compression_ratio()

    layers_id.sort(reverse=True)  # deepest layer first
    logger.info("=======> Start Compressing model with GRASP")
    if threshold_ratio is not None:
        logger.info("=======> Adaptive rank selection by taylor threshold %s", threshold_ratio)

    def target_names(layer_id):
        names = []
        if mlp_target_layer_types is not None:
            names += grasp_model.block_target_names(layer_id, "mlp", mlp_target_layer_types)
        if attn_target_layer_types is not None:
            names += grasp_model.block_target_names(layer_id, "attention", attn_target_layer_types)
        return names

    def hoist(position):
        # the SVDs factor ORIGINAL weights, so those of the next layers can run ahead in batched launches;
        # how many layers at once is bounded by the memory their factors take (all of them at configs[1])
        if not (hoist_svd and not allocation_aware):
            return
        ahead = layers_id[position:]
        if not ahead or all(n in grasp_model._svd_cache for n in target_names(ahead[0])):
            return
        count = grasp_model.svd_hoist_layers(ahead, target_names, device=device)
        grasp_model.precompute_svd([n for l in ahead[:count] for n in target_names(l)], device=device)

    hoist(0)
    # one forward sweep caches the input of the selected layers for all calibration samples (as many as fit)
    grasp_model.prepare_calibration(calibration_dataloader, layers_id, device=device)

    blocks = (("mlp", mlp_target_layer_types), ("attention", attn_target_layer_types))
    for position, layer_id in enumerate(tqdm(layers_id, desc="GRASP Compressing", total=len(layers_id), leave=True)):
        hoist(position)
        for block_type, target_layer_types in blocks:
            skip_flag = grasp_model.compress_block(layer_id=layer_id, block_type=block_type,
                                                   target_layer_types=target_layer_types, verbose=verbose,
                                                   device=device, allocation_aware=allocation_aware,
                                                   log_file=log_file)
            if skip_flag:
                logger.info("=======> Skip Compressing This Block")
                continue
            grasp_layer_grads = grasp_model.get_svdlayer_gradients(calibration_dataloader, device, log_file)
            indices_dict = grasp_model.dynamic_svd_selection(grasp_layer_grads, metric=metric,
                                                             compression_ratio=compression_ratio,
                                                             threshold_ratio=threshold_ratio, verbose=verbose,
                                                             log_file=log_file)
            grasp_model.compile_grasp_model(indices_dict, merge=merge, device=device, log_file=log_file)
    logger.info("=======> Done!")
    return grasp_model


def main(
    model_name_or_path: str,
    calibration_dataloader: DataLoader,
    layers_id: Optional[Union[List[int], int]] = None,
    num_prune_layers: Optional[int] = None,
    mlp_target_layer_types: Union[List[str], str] = ["down_proj", "up_proj", "gate_proj"],
    attn_target_layer_types: Union[List[str], str] = ["q_proj", "k_proj", "v_proj", "o_proj"],
    metric: Literal["gradient", "taylor"] = "taylor",
    compression_ratio: Optional[float] = None,
    threshold_ratio: Optional[float] = None,
    device: Literal["cuda", "cpu"] = "cuda",
    save_path: Optional[str] = None,
    angular: Optional[bool] = False,
    allocation_aware: Optional[bool] = False,
    merge: Optional[bool] = False,
    verbose: Optional[bool] = False,
    recovery: Optional[bool] = True,
    log_file: Optional[str] = None,
    train_device: Optional[str] = None,
    *args, **kwargs
):
    setup_logger(log_file)
    model, tokenizer = _load_model_and_tokenizer(model_name_or_path)
    grasp_model = GRASPModel(model=model)
    grasp_model.model.to(device=device)

    compress(grasp_model, calibration_dataloader, layers_id=layers_id, num_prune_layers=num_prune_layers,
             mlp_target_layer_types=mlp_target_layer_types, attn_target_layer_types=attn_target_layer_types,
             metric=metric, compression_ratio=compression_ratio, threshold_ratio=threshold_ratio, device=device,
             angular=angular, allocation_aware=allocation_aware, merge=merge, verbose=verbose, log_file=log_file)

    if not save_path:
        os.makedirs("./checkpoint", exist_ok=True)
        model_id: str = grasp_model.model.config._name_or_path
        save_path = os.path.join("./checkpoint", f"{model_id.replace('/', '-')}.pth")
    _save(grasp_model, save_path)

    if recovery:
        logger.info("=======> Starting recovery with efficient finetuning")
        try:
            from alpaca_grasp import train  # the reference's recovery module, not part of the hot path
        except ImportError as exc:
            raise NotImplementedError(
                "recovery fine-tuning is outside the B200 hot path; put the reference's alpaca_grasp.py on "
                "PYTHONPATH or pass recovery=False") from exc
        grasp_model = train(grasp_model=grasp_model, tokenizer=tokenizer, output_dir=os.path.dirname(save_path),
                            log_file=log_file, train_device=train_device, **kwargs)
        _save(grasp_model, save_path.replace(".pth", "_recovered.pth"))
    return grasp_model


def parse_args(argv=None):
    p = argparse.ArgumentParser(description="GRASP Model Compression")
    p.add_argument("--model_name_or_path", type=str, required=True)
    p.add_argument("--dataset_name", type=str, default="wikitext2")
    # compression
    p.add_argument("--layers_id", type=int, nargs="+", default=None)
    p.add_argument("--num_prune_layers", type=int, default=None)
    p.add_argument("--mlp_target_layer_types", type=str, nargs="+", default=["down_proj", "up_proj", "gate_proj"])
    p.add_argument("--attn_target_layer_types", type=str, nargs="+", default=["q_proj", "k_proj", "v_proj", "o_proj"])
    p.add_argument("--metric", type=str, choices=["gradient", "taylor"], default="taylor")
    p.add_argument("--compression_ratio", type=float, default=None)
    p.add_argument("--threshold_ratio", type=float, default=None)
    p.add_argument("--device", type=str, choices=["cuda", "cpu"], default="cuda")
    p.add_argument("--save_path", type=str, default=None)
    p.add_argument("--angular", action="store_true")
    p.add_argument("--allocation_aware", action="store_true")
    p.add_argument("--merge", action="store_true")
    p.add_argument("--verbose", action="store_true")
    # calibration
    p.add_argument("--num_samples", type=int, default=1024)
    p.add_argument("--batch_size", type=int, default=1)
    p.add_argument("--seq_len", type=int, default=512)
    p.add_argument("--padding", type=str, default="max_length")
    p.add_argument("--recovery", action="store_true")
    p.add_argument("--log_file", type=str, default=None)
    # recovery (forwarded to the reference's alpaca_grasp.train)
    p.add_argument("--data_path", type=str, default='yahma/alpaca-cleaned')
    p.add_argument("--train_batch_size", type=int, default=32)
    p.add_argument("--micro_batch_size", type=int, default=4)
    p.add_argument("--num_epochs", type=int, default=1)
    p.add_argument("--learning_rate", type=float, default=3e-4)
    p.add_argument("--max_length", type=int, default=256)
    p.add_argument("--val_set_size", type=int, default=2000)
    p.add_argument("--train_on_inputs", action="store_true")
    p.add_argument("--add_eos_token", action="store_true")
    p.add_argument("--resume_from_checkpoint", type=str, default=None)
    p.add_argument("--prompt_template_name", type=str, default="alpaca")
    p.add_argument("--train_device", type=str, default="0")
    # evaluation (forwarded to the reference's evaluate_grasp.evaluate_model)
    p.add_argument("--evaluate", action="store_true")
    p.add_argument("--eval_ppl", type=str, default="wikitext2,ptb,c4")
    p.add_argument("--eval_tasks", type=str,
                   default="boolq,piqa,hellaswag,winogrande,arc_easy,arc_challenge,openbookqa,mathqa")
    p.add_argument("--num_fewshot", type=int, default=0)
    p.add_argument("--limit", type=int, default=-1)
    args = p.parse_args(argv)
    # run_grasp.sh joins the layer-type lists with commas; accept both spellings
    for key in ("mlp_target_layer_types", "attn_target_layer_types"):
        vals = getattr(args, key)
        if vals is not None:
            setattr(args, key, [t for v in vals for t in v.split(",") if t])
    return args


def _calibration_dataloader(args, tokenizer):
    """Batches in the reference loader's format (dataset/loader.py): the reference's own loader when it is importable,
    else grasp_b200.loader (same sampling, chunking and batch format; corpora read from ./datasets as there)."""
    if args.dataset_name != "synthetic":
        try:
            from dataset.loader import get_calibration_dataloader  # the reference's loader, if on PYTHONPATH
            return get_calibration_dataloader(dataset_name=args.dataset_name, tokenizer=tokenizer,
                                              num_samples=args.num_samples, batch_size=args.batch_size,
                                              seq_len=args.seq_len, padding=args.padding)
        except ImportError:
            pass
    from grasp_b200.loader import get_calibration_dataloader
    return get_calibration_dataloader(dataset_name=args.dataset_name, tokenizer=tokenizer, num_samples=args.num_samples,
                                      batch_size=args.batch_size, seq_len=args.seq_len, padding=args.padding)


if __name__ == "__main__":
    try:
        from setproctitle import setproctitle
        setproctitle("GRASP")
    except ImportError:
        pass
    args = parse_args()
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and args.device == "cuda":
        # torchrun: one process per GPU; the hot path shards matrices and calibration samples over the ranks
        import torch.distributed as td
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        td.init_process_group("nccl")
    from transformers import AutoTokenizer
    tokenizer = AutoTokenizer.from_pretrained(args.model_name_or_path)
    tokenizer.pad_token = tokenizer.eos_token
    calibration_dataloader = _calibration_dataloader(args, tokenizer)

    kwargs = {}
    if args.recovery:
        kwargs = {"data_path": args.data_path, "batch_size": args.train_batch_size,
                  "mirco_batch_size": args.micro_batch_size, "num_epochs": args.num_epochs,
                  "learning_rate": args.learning_rate, "max_length": args.max_length,
                  "val_set_size": args.val_set_size, "train_on_inputs": args.train_on_inputs,
                  "add_eos_token": args.add_eos_token, "resume_from_checkpoint": args.resume_from_checkpoint,
                  "prompt_template_name": args.prompt_template_name}

    grasp_model = main(model_name_or_path=args.model_name_or_path, calibration_dataloader=calibration_dataloader,
                       layers_id=args.layers_id, num_prune_layers=args.num_prune_layers,
                       mlp_target_layer_types=args.mlp_target_layer_types,
                       attn_target_layer_types=args.attn_target_layer_types, metric=args.metric,
                       compression_ratio=args.compression_ratio, threshold_ratio=args.threshold_ratio,
                       device=args.device, save_path=args.save_path, angular=args.angular,
                       allocation_aware=args.allocation_aware, merge=args.merge, verbose=args.verbose,
                       recovery=args.recovery, log_file=args.log_file, train_device=args.train_device, **kwargs)

    if args.evaluate:
        try:
            from evaluate_grasp import evaluate_model  # reference module (lm-eval harness), not the hot path
        except ImportError as exc:
            raise NotImplementedError("evaluation is outside the hot path; the reference's evaluate_grasp.py "
                                      "(and lm_eval) must be importable for --evaluate") from exc
        evaluate_model(model=grasp_model.model, tokenizer=tokenizer, model_name=args.model_name_or_path,
                       tasks=args.eval_tasks, eval_ppl=args.eval_ppl, num_fewshot=args.num_fewshot,
                       limit=args.limit, batch_size=args.batch_size, device=args.device, log_file=args.log_file)
