"""Device-time breakdown of a DEEP sigma-gradient pass: the block being compressed is dense, the 7 layers above it
are already compressed to rank-k factor pairs (what the passes of layers 24..30 look like in the 7B job)."""
import sys, torch
sys.path.insert(0, ".")
from grasp_b200 import synth, engine, ops
from modeling_grasp import GRASPModel, SVDLinear
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda")
model = synth.random_llama("llama2-7b", seed=0, device=dev, num_hidden_layers=8)
gm = GRASPModel(model); gm.micro_batch = 16
def fake_compress(lin, ratio=0.9):
    out_f, in_f = lin.weight.shape
    k = int(in_f * out_f * (1 - ratio) / (in_f + out_f))
    return SVDLinear.from_packed(torch.randn(k, in_f, device=dev) * 0.02, torch.randn(out_f, k, device=dev) * 0.02, None)
for i in range(1, 8):
    L = model.model.layers[i]
    for owner, names in ((L.self_attn, ("q_proj", "k_proj", "v_proj", "o_proj")), (L.mlp, ("gate_proj", "up_proj", "down_proj"))):
        for n in names:
            new = fake_compress(getattr(owner, n)); new.requires_grad_(False); setattr(owner, n, new)
tokens = synth.random_tokens(32, 512, 32000, seed=0)
dl = synth.calibration_dataloader(0, 0, 0, tokens=tokens)
gm.prepare_calibration(dl, [0])
for block, types in (("mlp", ["gate_proj", "up_proj", "down_proj"]),):
    gm.compress_block(0, block, types, device=dev)
    gm.get_svdlayer_gradients(dl, dev); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); gm.get_svdlayer_gradients(dl, dev); e1.record(); torch.cuda.synchronize()
    print(f"{block} pass, 32 samples, 1 dense + 7 compressed layers: {e0.elapsed_time(e1):.1f} ms")
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        gm.get_svdlayer_gradients(dl, dev); torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=24, max_name_column_width=64))
