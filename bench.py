#!/usr/bin/env python
"""bench.py -- headline benchmark of the GRASP compression hot path on B200.

    python bench.py --gpus N --steps K --warmup W            (our arm; torchrun launches it for N > 1)
    python bench.py --impl reference --gpus N --steps K ...  (the reference's CPU path on the host cores)

Workload = BASELINE.json configs[1], whatever K is: random-init LLaMA-2-7B shapes, NUM_PRUNE_LAYERS=8,
COMPRESSION_RATIO=0.9, 512 synthetic calibration samples x 512 tokens, fp32.  One STEP = one pruned decoder
layer compressed end to end (7 weight matrices: SVD, two sigma-gradient passes over all calibration samples,
selection, compile) plus its 1/8 share of the job's layer-scoring stage.  K steps are run as consecutive
executions of the 8-layer job on restored weights (the restore is outside the timed region); when K is not
a multiple of 8 the last execution stops after K mod 8 layers -- it still scores all layers over all
samples, so the figure errs on the slow side.  metric = weight matrices compressed per second = 7K / time.

The timed region is the public API call a user makes -- grasp.compress(GRASPModel, DataLoader, ...) --
with the calibration tokens in pinned HOST memory (e2e) and the retained index sets read back to the host.
`value` is the same run with the separately event-timed host->device copy of the tokens taken out.
Times are CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "weight_matrices_compressed_per_sec"
UNIT = "matrices/s"
MATRICES_PER_LAYER = 7
JOB_LAYERS = 8            # BASELINE.json configs[1]: NUM_PRUNE_LAYERS=8


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=8, help="pruned layers compressed in the timed region (8 = one job)")
    p.add_argument("--job-layers", type=int, default=JOB_LAYERS, help="NUM_PRUNE_LAYERS of the job (configs[1]: 8)")
    p.add_argument("--no-library-baseline", action="store_true")
    p.add_argument("--save-kept", default="", help="write the retained index sets + scores of the first execution (N=1 reference for the N>1 runs)")
    p.add_argument("--kept-reference", default=os.path.join(ROOT, "profiles", "r02_kept_indices_llama2-7b_1gpu.pt"))
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", choices=["ours", "reference"], default="ours")
    p.add_argument("--model", default="llama2-7b")
    p.add_argument("--samples", type=int, default=512)
    p.add_argument("--seq-len", type=int, default=512)
    p.add_argument("--ratio", type=float, default=0.9)
    p.add_argument("--micro-batch", type=int, default=0, help="0 = chosen from the tile/wave model")
    p.add_argument("--warmup-samples", type=int, default=32)
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--cpu-svd-dim", type=int, default=0, help="override the CPU sample's matrix size (debug)")
    return p.parse_args()


def workload_name(a):
    return (f"random-init {a.model}, NUM_PRUNE_LAYERS={a.job_layers}, COMPRESSION_RATIO={a.ratio}, "
            f"{a.samples} samples x {a.seq_len} tokens, fp32")


# ------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(names, r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        busy = [v for v in sm if v > 0.5 * (max(mx) if mx else 1)] or sm
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------- CPU reference arm
def cpu_reference_sample(a, threads: int):
    """Time the reference's CPU path (oracle port of modeling_grasp.py on torch CPU fp32) on a bounded
    sample of the workload and extrapolate to one execution of the job.  Returns (dict, seconds per job)."""
    from grasp_b200 import synth
    from oracle import restate
    torch.set_num_threads(threads)
    cfg = dict(synth.MODEL_CONFIGS[a.model])
    d, ff, L, V = cfg["hidden_size"], cfg["intermediate_size"], cfg["num_hidden_layers"], cfg["vocab_size"]
    if a.cpu_svd_dim:
        d, ff = a.cpu_svd_dim, a.cpu_svd_dim * 11 // 4
        cfg.update(hidden_size=d, intermediate_size=ff, num_attention_heads=max(1, d // 128),
                   num_key_value_heads=max(1, d // 128))
    # (1) the SVD of one attention-shaped matrix (reference modeling_grasp.py:231)
    W = torch.randn(d, d) * 0.02
    t0 = time.perf_counter()
    U, S, Vh = restate.svd(W)
    t_svd = time.perf_counter() - t0
    # (2) one calibration sample through a one-layer model of the same widths: dense forward, then the
    #     reference's GRASPLayer forward+backward for the attention block
    model = synth.random_llama(a.model, seed=0, num_hidden_layers=1, **{k: cfg[k] for k in
                               ("hidden_size", "intermediate_size", "num_attention_heads", "num_key_value_heads")})
    for p in model.parameters():
        p.requires_grad = False
    tok = synth.random_tokens(1, a.seq_len, V, seed=1)
    batch = restate.batches_from_tokens(tok)
    with torch.no_grad():
        t0 = time.perf_counter()
        model(input_ids=batch[0]["input_ids"], use_cache=False)
        t_fwd_layer_head = time.perf_counter() - t0
        t0 = time.perf_counter()
        h = model.model.embed_tokens(batch[0]["input_ids"])
        pos = torch.arange(h.shape[1]).unsqueeze(0)
        model.model.layers[0](h, position_ids=pos, position_embeddings=model.model.rotary_emb(h, position_ids=pos))
        t_fwd_layer = time.perf_counter() - t0
    # attention block with the already measured SVD reused for all four projections (same shape)
    for name in restate.block_names(0, "attention", ["q_proj", "k_proj", "v_proj", "o_proj"]):
        lin = model.get_submodule(name)
        if lin.weight.shape == W.shape:
            restate._set_module(model, name, restate.OracleGRASPLayer(U, S, Vh))
    t0 = time.perf_counter()
    restate.svdlayer_gradients(model, batch)
    t_pass_attn = time.perf_counter() - t0
    n_layers_in_attn = len(restate.grasp_layer_names(model))
    sample = {"svd_s": t_svd, "fwd_layer_s": t_fwd_layer, "fwd_layer_plus_head_s": t_fwd_layer_head,
              "attn_block_fwd_bwd_s": t_pass_attn, "grasp_layers_in_attn_sample": n_layers_in_attn}
    # extrapolation to one job (labelled as such): per selected layer 4 attention SVDs + 3 MLP SVDs
    # (MLP cost scaled by the nominal flop ratio), 2 passes x samples x (forward of the other layers + block
    # forward/backward; MLP passes cost at least the attention ones), plus the BI forward of every sample.
    from grasp_b200.dist import svd_cost
    mlp_ratio = svd_cost(ff, d) / svd_cost(d, d)
    t_svd_layer = t_svd * (4 + 3 * mlp_ratio)
    t_head = max(t_fwd_layer_head - t_fwd_layer, 0.0)
    t_full_fwd = L * t_fwd_layer + t_head
    t_pass_sample = (L - 1) * t_fwd_layer + t_pass_attn
    total = a.job_layers * (t_svd_layer + 2 * a.samples * t_pass_sample) + a.samples * t_full_fwd
    sample["extrapolated_job_s"] = total
    sample["extrapolation"] = ("job_s = NUM_PRUNE_LAYERS*(svd_s*(4+3*%.3f) + 2*samples*((L-1)*fwd_layer_s + attn_block_fwd_bwd_s))"
                               " + samples*(L*fwd_layer_s + head_s)") % mlp_ratio
    return sample, total


CPU_SAMPLE_DESC = ("oracle/restate.py (torch CPU fp32, the reference's own library calls) on the host cores: one %dx%d "
                   "torch.linalg.svd + 1 calibration sample through a 1-layer model of the named widths (dense "
                   "forward, reference GRASPLayer attention-block forward/backward); one job extrapolated: "
                   "4+3x(MLP flop ratio) SVDs per layer, 2 passes x %d samples x (31 dense layer forwards + block "
                   "pass), BI forward of every sample (formula in primitives.extrapolation); the job itself "
                   "would take ~a day on these cores")


def cpu_baseline_entry(a):
    threads = os.cpu_count() or 1
    t0 = time.perf_counter()
    sample, job_s = cpu_reference_sample(a, threads)
    return {"value": MATRICES_PER_LAYER * a.job_layers / job_s, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": CPU_SAMPLE_DESC % (a.cpu_svd_dim or 4096, a.cpu_svd_dim or 4096, a.samples),
            "sample_wall_s": time.perf_counter() - t0, "extrapolated": True, "primitives": sample}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    # every step times the same bounded sample of the job on the host cores; `value` is the metric of the whole
    # job extrapolated from the timed primitives (the job itself takes ~a day on CPU), `ms_per_step` the wall
    # time of what was actually run per step.  The number of steps is cut so that the arm ends within minutes.
    jobs, walls, sample, budget_s = [], [], None, 240.0
    planned = a.warmup + a.steps
    for i in range(planned):
        t0 = time.perf_counter()
        sample, job_s = cpu_reference_sample(a, threads)
        dt = time.perf_counter() - t0
        if i >= a.warmup or i == planned - 1:
            jobs.append(job_s)
            walls.append(dt)
        if i == 0 and dt * planned > budget_s:
            planned = max(1, int(budget_s / dt))
            if planned <= a.warmup:                # no room for untimed iterations: keep what was measured
                jobs, walls = [job_s], [dt]
        if i + 1 >= planned:
            break
    job_s = statistics.median(jobs)
    value = MATRICES_PER_LAYER * a.job_layers / job_s
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": 1000.0 * statistics.mean(walls), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": workload_name(a), "extrapolated": True, "timed_steps": len(walls),
                       "value_is": "matrices/s of the whole job extrapolated from the primitives timed in every step; "
                                   "ms_per_step is the wall time of one timed sample",
                       "extrapolated_job_s": job_s, "extrapolated_ms_per_layer": 1000.0 * job_s / a.job_layers,
                       "primitives": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "extrapolated": True,
                             "sample": CPU_SAMPLE_DESC % (a.cpu_svd_dim or 4096, a.cpu_svd_dim or 4096, a.samples)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------- library baseline on the same GPU
def gpu_library_sample(a, model, dev):
    """The reference's own torch path on THIS GPU (cuSOLVER torch.linalg.svd, cuBLAS fp32 GEMMs, autograd through
    the reference's GRASPLayer; oracle/restate.py moved to cuda) on a bounded sample, extrapolated to one job.
    This is the bar the kernels have to beat; the CPU figure only says what 16 host cores do."""
    from grasp_b200 import synth
    from oracle import restate
    cfg = synth.MODEL_CONFIGS[a.model]
    L, V = cfg["num_hidden_layers"], cfg["vocab_size"]

    def timed(fn, n=1):
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for _ in range(n):
            out = fn()
        torch.cuda.synchronize(dev)
        return (time.perf_counter() - t0) / n, out

    last = model.model.layers[L - 1]
    torch.linalg.svd(torch.randn(256, 256, device=dev), full_matrices=False)        # cuSOLVER handle
    t_svd_attn, usv_attn = timed(lambda: restate.svd(last.self_attn.o_proj.weight.data))
    t_svd_mlp, usv_mlp = timed(lambda: restate.svd(last.mlp.up_proj.weight.data))
    tok = synth.random_tokens(3, a.seq_len, V, seed=2).to(dev)
    batches = restate.batches_from_tokens(tok)
    with torch.no_grad():
        restate.compute_bi(model, batches[:1], 1)
        t_bi, _ = timed(lambda: restate.compute_bi(model, batches[1:], 1))
    t_bi /= 2
    res = {"svd_4096x4096_s": t_svd_attn, "svd_mlp_s": t_svd_mlp, "bi_sample_s": t_bi}
    saved = {}
    for block, types, usv in (("attention", ["o_proj"], usv_attn), ("mlp", ["up_proj"], usv_mlp)):
        # one GRASPLayer per block type carries the measured factors; the other matrices of the block reuse them
        # when the shape matches (the time of a pass does not depend on the values)
        all_types = ["q_proj", "k_proj", "v_proj", "o_proj"] if block == "attention" else ["gate_proj", "up_proj", "down_proj"]
        for name in restate.block_names(L - 1, block, all_types):
            lin = model.get_submodule(name)
            saved[name] = lin
            if tuple(lin.weight.shape) == (usv[0].shape[0], usv[2].shape[1]):
                f = usv
            else:
                f = restate.svd(lin.weight.data)
            restate._set_module(model, name, restate.OracleGRASPLayer(*f))
        restate.svdlayer_gradients(model, batches[:1])
        t_pass, _ = timed(lambda: restate.svdlayer_gradients(model, batches[1:]))
        res[f"{block}_pass_sample_s"] = t_pass / 2
        for name, lin in saved.items():
            restate._set_module(model, name, lin)
        saved = {}
        model.zero_grad(set_to_none=True)
        torch.cuda.empty_cache()
    job_s = (a.job_layers * (4 * t_svd_attn + 3 * t_svd_mlp) + a.samples * t_bi
             + a.job_layers * a.samples * (res["attention_pass_sample_s"] + res["mlp_pass_sample_s"]))
    res["extrapolated_job_s"] = job_s
    return {"value": MATRICES_PER_LAYER * a.job_layers / job_s, "unit": UNIT, "kind": "reference algorithm on torch "
            "CUDA library kernels (cuSOLVER svd, cuBLAS fp32, autograd GRASPLayer), same GPU", "extrapolated": True,
            "sample": "torch.linalg.svd of one 4096x4096 and one MLP weight on cuda; 2 calibration samples of the "
                      "reference's full-model forward (hidden states + BI) and of its forward/backward with the deepest "
                      "layer's attention / MLP block as GRASPLayers (the cheapest of the 16 passes: backward ends at "
                      "layer 31); job = 8*(4 svd_attn + 3 svd_mlp) + 512*bi + 8*512*(attn pass + mlp pass)",
            "primitives": res}


# ------------------------------------------------------------------------------------------- our arm
class StageTimer:
    """CUDA-event timers around the public GRASPModel stage methods (same stream as the kernels)."""

    def __init__(self):
        self.spans = []

    def wrap(self, obj, name, stage):
        fn = getattr(obj, name)

        def timed(*args, **kw):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn(*args, **kw)
            e1.record()
            self.spans.append((stage, e0, e1))
            return out
        setattr(obj, name, timed)

    def totals(self):
        out = {}
        for stage, e0, e1 in self.spans:
            out[stage] = out.get(stage, 0.0) + e0.elapsed_time(e1)
        return out


def run_ours(a):
    import gc
    import torch.distributed as td
    from grasp_b200 import ops, synth
    import grasp
    from modeling_grasp import GRASPModel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        td.init_process_group("nccl", device_id=dev)

    cfg = synth.MODEL_CONFIGS[a.model]
    V, L = cfg["vocab_size"], cfg["num_hidden_layers"]
    model = synth.random_llama(a.model, seed=0, device=dev)             # identical replicas (same seed)
    tokens = synth.random_tokens(a.samples, a.seq_len, V, seed=0).pin_memory()
    gm = GRASPModel(model)
    gm.micro_batch = a.micro_batch
    # the dense linears of every layer, to put the model back between executions of the job (outside the timing)
    originals = {n: m for n, m in model.named_modules() if isinstance(m, torch.nn.Linear) and ".layers." in n}

    def restore():
        for n, m in originals.items():
            if model.get_submodule(n) is not m:
                gm._set_module(model, n, m)
        gm._runner, gm._calib, gm._svd_cache = None, None, {}
        gm.grasp_values_dict = {}
        gc.collect()
        torch.cuda.empty_cache()

    def barrier():
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    # ---- warm-up: W untimed layer steps on the deepest layer (restored afterwards), fewer samples
    wtok = tokens[: min(a.warmup_samples, a.samples)]
    for _ in range(a.warmup):
        wdl = synth.calibration_dataloader(0, 0, 0, tokens=wtok)
        grasp.compress(gm, wdl, layers_id=[L - 1], compression_ratio=a.ratio, device=dev)
        restore()
    if a.warmup:
        # the layer-scoring forward also gets warm
        wdl = synth.calibration_dataloader(0, 0, 0, tokens=wtok)
        gm.compute_bi(num_prune_layers=1, calibration_dataloader=wdl, device=dev)
        restore()
    torch.cuda.reset_peak_memory_stats()

    # ---- timed region: the public API call with host-resident tokens, once per execution of the job
    timer = StageTimer()
    for name, stage in (("compute_bi", "layer_scoring"), ("precompute_svd", "svd"),
                        ("prepare_calibration", "prefix_cache"), ("get_svdlayer_gradients", "sigma_gradients"),
                        ("dynamic_svd_selection", "selection"), ("compile_grasp_model", "compile")):
        timer.wrap(gm, name, stage)
    picked = []                       # every block's retained index sets (device tensors until the read-back)
    select = gm.dynamic_svd_selection

    score_inputs = []                 # (name, sigma-gradient, singular values) per matrix of the FIRST execution

    def spy(grads, *args, **kw):
        out = select(grads, *args, **kw)
        picked.append(dict(out))
        if not job_ms:                # references only, no kernels: the scores are formed after the timing
            score_inputs.extend((n, grads[n], gm.model.get_submodule(n).S.data) for n in grads)
        return out
    gm.dynamic_svd_selection = spy

    ops.timers.reset(enabled=True)
    clocks = ClockSampler(local_rank)
    launches0 = ops.launch_count()
    job_ms, job_resident_ms, job_layers_done, d2h_bytes, checksums, layers_chosen = [], [], [], 0, [], None
    first_kept = None
    micro_batch, pass_mb = None, None
    remaining = a.steps
    clocks.start()
    while remaining > 0:
        n = min(a.job_layers, remaining)
        dl = synth.calibration_dataloader(0, 0, 0, tokens=tokens)
        del picked[:]
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if n == a.job_layers:
            grasp.compress(gm, dl, num_prune_layers=a.job_layers, compression_ratio=a.ratio, device=dev)
        else:
            # a truncated execution: the same calls grasp.compress makes, stopping after the n deepest layers
            _, ids = gm.compute_bi(num_prune_layers=a.job_layers, calibration_dataloader=dl, device=dev)
            grasp.compress(gm, dl, layers_id=sorted(ids, reverse=True)[:n], compression_ratio=a.ratio, device=dev)
        kept = [{k: (v if torch.is_tensor(v) else torch.as_tensor(v)).cpu() for k, v in blk.items()} for blk in picked]
        importances = list(gm.layer_importances) if getattr(gm, "layer_importances", None) else []   # D2H of the result
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        h2d_ms = getattr(gm._calib[1], "h2d_ms", 0.0)
        t = torch.tensor([ms, ms - h2d_ms], dtype=torch.float64, device=dev)
        if world > 1:
            td.all_reduce(t, op=td.ReduceOp.MAX)
        job_ms.append(t[0].item())
        job_resident_ms.append(t[1].item())
        job_layers_done.append(n)
        d2h_bytes += sum(v.numel() * 8 for blk in kept for v in blk.values()) + len(importances) * 8
        flat = {k: v for blk in kept for k, v in blk.items()}
        checksums.append(int(sum(int(flat[k].sum()) * (i + 1) for i, k in enumerate(sorted(flat)))))
        if len(job_ms) == 1:
            first_kept = flat
        if n == a.job_layers or layers_chosen is None:
            layers_chosen = list(gm.redundant_layers)
        micro_batch = gm._runner.micro_batch if gm._runner else a.micro_batch
        pass_mb = getattr(gm._runner, "last_pass_micro_batch", None) if gm._runner else None
        remaining -= n
        restore()                     # outside the timed region
    clk = clocks.stop()
    launches = ops.launch_count() - launches0
    total_ms, resident_ms = sum(job_ms), sum(job_resident_ms)

    if rank == 0:
        n_mat = MATRICES_PER_LAYER * a.steps
        stages = timer.totals()
        kt = ops.timers.summary()
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        roof = None
        if kt:
            # the dominant single kernel: the tensor-core GEMM (the SVD entry is a composite of ~10 kernels
            # per Jacobi round and is reported beside it)
            gemm_tags = {k: v for k, v in kt.items() if k.startswith("grasp_gemm_f")}
            tag, rec = max((gemm_tags or kt).items(), key=lambda kv: kv[1]["ms"])
            peak = peaks.get("bf16_tflops_sustained", 1400.0)
            achieved = rec["flops"] / (rec["ms"] * 1e-3) / 1e12
            traffic, traffic_note = None, None
            for fname in ("r02_ncu_traffic.json", "r01_ncu_traffic.json"):
                try:
                    tr = json.load(open(os.path.join(ROOT, "profiles", fname)))["tc_gemm_kernel"]
                    traffic = tr["dram_read_bytes"] + tr["dram_write_bytes"]
                    traffic_note = ("ncu dram__bytes_read+write of ONE launch, %s (algorithmic %d bytes); the run's launches "
                                    "mix shapes, see profiles/%s") % (tr["launch"], tr["algorithmic_bytes"], fname)
                    break
                except (OSError, KeyError, ValueError):
                    continue
            roof = {"bound": "tensor", "kernel": tag, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                    "frac": achieved / peak, "traffic": traffic, "traffic_note": traffic_note,
                    "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback",
                    "launches": rec["n"], "avg_launch_ms": rec["ms"] / max(rec["n"], 1),
                    "issued_mma_tflops": achieved * rec["mma_per_flop"],
                    "issued_mma_frac": achieved * rec["mma_per_flop"] / peak,
                    "share_of_step": rec["ms"] / total_ms,
                    "all_kernels_ms": {k: round(v["ms"], 3) for k, v in kt.items()}}
        n_jobs = len(job_ms)
        h2d = n_jobs * 2 * a.samples * (a.seq_len - 1) * 8
        line = {"metric": METRIC, "value": n_mat / (resident_ms * 1e-3), "unit": UNIT, "n_gpus": world,
                "steps": a.steps, "warmup": a.warmup, "ms_per_step": total_ms / a.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None,
                "dtype": "f32 (fp32 data; GEMMs as 3 fp16-plane tcgen05 products with fp32 accumulation, rel. err 3e-7; SVD bf16x6 planes + fp32 clean-up)",
                "data": "synthetic",
                "config": {"workload": workload_name(a), "step": "one pruned layer (7 matrices) of the job",
                           "executions": [{"layers": n, "seconds": round(ms / 1e3, 3)} for n, ms in
                                          zip(job_layers_done, job_ms)],
                           "micro_batch": micro_batch, "last_pass_micro_batch": pass_mb,
                           "l2": "inputs larger than L2 (27 GB of weights, 4.3 GB activation cache per layer)",
                           "warmup_samples": min(a.warmup_samples, a.samples), "layers_chosen": layers_chosen,
                           "kept_index_checksum": checksums[0], "kept_index_checksums": checksums,
                           "stages_ms": {k: round(v, 2) for k, v in stages.items()},
                           "end_to_end_s": total_ms / 1e3,
                           "full_job_s": (round(job_ms[0] / 1e3, 3) if job_layers_done[0] == a.job_layers else None)},
                "e2e": {"value": n_mat / (total_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d // a.steps,
                        "d2h_bytes_per_step": d2h_bytes // a.steps},
                "gpu_launches": int(launches), "clocks": clk, "roofline": roof,
                "memory": {"peak_allocated_gb": round(torch.cuda.max_memory_allocated() / 2**30, 1),
                           "peak_reserved_gb": round(torch.cuda.max_memory_reserved() / 2**30, 1),
                           "alloc_retries": torch.cuda.memory_stats().get("num_alloc_retries", 0)}}
        # retained index sets against the committed single-GPU run of the same job: they may differ only at ties
        # (multi-GPU changes the order of the fp32 sums), i.e. a swapped index scores within a hair of the k-th score
        scores = {n: (g * S).abs().cpu() for n, g, S in score_inputs}
        if a.save_kept and job_layers_done[0] == a.job_layers:
            torch.save({"workload": workload_name(a), "n_gpus": world, "kept": first_kept,
                        "scores": {n: v.half() for n, v in scores.items()}}, a.save_kept)
        if os.path.exists(a.kept_reference) and job_layers_done[0] == a.job_layers:
            try:
                refk = torch.load(a.kept_reference, map_location="cpu", weights_only=False)
                if refk.get("workload") == workload_name(a):
                    jmin, swapped, tie, total = 1.0, 0, 0.0, 0
                    for n, theirs in refk["kept"].items():
                        ours_set, theirs_set = set(first_kept[n].tolist()), set(theirs.tolist())
                        total += len(theirs_set)
                        jmin = min(jmin, len(ours_set & theirs_set) / max(len(ours_set | theirs_set), 1))
                        kth = scores[n][first_kept[n][-1]].item()
                        for i in ours_set ^ theirs_set:
                            swapped += 1
                            tie = max(tie, abs(scores[n][i].item() - kth) / max(kth, 1e-30))
                    line["config"]["index_sets_vs_1gpu_run"] = {
                        "reference": os.path.relpath(a.kept_reference, ROOT), "matrices": len(refk["kept"]),
                        "min_jaccard": jmin, "swapped_indices": swapped // 2, "of": total,
                        "worst_tie_margin_of_kth_score": tie}
            except Exception as exc:
                line["config"]["index_sets_vs_1gpu_run"] = {"unavailable": f"{type(exc).__name__}: {exc}"[:200]}
        if not a.no_library_baseline:
            try:
                line["library_baseline"] = gpu_library_sample(a, model, dev)
            except Exception as exc:                      # a baseline leg must never take the headline down
                line["library_baseline"] = {"unavailable": f"{type(exc).__name__}: {exc}"[:300]}
        if not a.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_entry(a)
        print(json.dumps(line))
    if world > 1:
        # the other ranks wait for rank 0's baseline legs on the rendezvous store (a blocking socket read): a NCCL
        # barrier would have them spin on the host cores the CPU baseline is timed on
        import datetime
        store = td.distributed_c10d._get_default_store()
        if rank == 0:
            store.set("grasp_bench_done", "1")
        else:
            store.wait(["grasp_bench_done"], datetime.timedelta(minutes=30))
        td.destroy_process_group()


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
