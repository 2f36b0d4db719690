"""Where does one sigma-gradient micro-batch spend its GPU time?  (torch profiler, 4-layer model of 7B widths)"""
import sys, torch
sys.path.insert(0, ".")
from grasp_b200 import synth, engine, ops
import grasp
from modeling_grasp import GRASPModel
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda")
model = synth.random_llama("llama2-7b", seed=0, device=dev, num_hidden_layers=4)
gm = GRASPModel(model); gm.micro_batch = 8
tokens = synth.random_tokens(16, 512, 32000, seed=0)
dl = synth.calibration_dataloader(0, 0, 0, tokens=tokens)
gm.prepare_calibration(dl, [0])
gm.compress_block(0, "attention", ["q_proj", "k_proj", "v_proj", "o_proj"], device=dev)
gm.get_svdlayer_gradients(dl, dev)          # warm
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
gm.get_svdlayer_gradients(dl, dev)
torch.cuda.synchronize()
print(f"wall time of one pass (16 samples, 4 layers): {(time.perf_counter() - t0) * 1e3:.1f} ms")
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    gm.get_svdlayer_gradients(dl, dev)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=70))
# ---- the layer-scoring forward (stage 1) on the same model
gm2 = GRASPModel(synth.random_llama("llama2-7b", seed=0, device=dev, num_hidden_layers=4)); gm2.micro_batch = 8
gm2.compute_bi(num_prune_layers=1, calibration_dataloader=dl, device=dev)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    gm2.compute_bi(num_prune_layers=1, calibration_dataloader=dl, device=dev)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=70))
