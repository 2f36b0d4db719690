"""TEST INFRASTRUCTURE -- BASELINE.json configs[0] timed IN FULL on the host cores: the oracle (oracle/restate.py =
the reference's torch CPU path, function by function) on random-init TinyLlama-1.1B shapes, NUM_PRUNE_LAYERS=2,
COMPRESSION_RATIO=0.9, 32 synthetic calibration samples x 512 tokens.  Writes one JSON record (stage seconds,
matrices/s, cores) -- the measured, not extrapolated, CPU figure that bench.py's extrapolation can be checked against.

    python oracle/time_config0.py [out.json]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from grasp_b200 import synth  # noqa: E402
from oracle import restate  # noqa: E402


def main(out_path):
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    model = synth.random_llama("tinyllama-1.1b", seed=0)
    tokens = synth.random_tokens(32, 512, model.config.vocab_size, seed=0)
    for p in model.parameters():
        p.requires_grad = False
    batches = restate.batches_from_tokens(tokens)
    t = {}
    t0 = time.perf_counter()
    imp, layers = restate.compute_bi(model, batches, 2)
    t["layer_scoring_s"] = time.perf_counter() - t0
    layers = sorted(layers, reverse=True)
    t["svd_s"] = t["sigma_gradients_s"] = t["selection_s"] = t["compile_s"] = 0.0
    for lid in layers:
        for block, types in (("mlp", ("down_proj", "up_proj", "gate_proj")), ("attention", ("q_proj", "k_proj", "v_proj", "o_proj"))):
            t0 = time.perf_counter()
            restate.compress_block(model, lid, block, types)
            t["svd_s"] += time.perf_counter() - t0
            names = restate.grasp_layer_names(model)
            mods = {n: model.get_submodule(n) for n in names}
            t0 = time.perf_counter()
            grads = restate.svdlayer_gradients(model, batches)
            t["sigma_gradients_s"] += time.perf_counter() - t0
            t0 = time.perf_counter()
            idx, _ = restate.select(grads, mods, "taylor", 0.9, None)
            t["selection_s"] += time.perf_counter() - t0
            t0 = time.perf_counter()
            restate.compile_model(model, idx, False)
            t["compile_s"] += time.perf_counter() - t0
            print(lid, block, {k: round(v, 1) for k, v in t.items()}, flush=True)
    total = sum(t.values())
    rec = {"workload": "random-init tinyllama-1.1b, NUM_PRUNE_LAYERS=2, COMPRESSION_RATIO=0.9, 32 samples x 512 tokens, fp32",
           "impl": "oracle/restate.py (reference algorithm, torch CPU fp32)", "cores": threads, "layers_chosen": layers,
           "total_s": total, "matrices": 14, "matrices_per_s": 14 / total, "stages": t, "torch": torch.__version__,
           "extrapolated": False}
    print(json.dumps(rec))
    if out_path:
        with open(out_path, "w") as f:
            json.dump(rec, f, indent=1)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "")
