"""grasp_b200 -- B200-native (sm_100a) implementation of the GRASP compression hot path.

Layout:
  _lib.py    ctypes binding of libgrasp_b200.so (the C ABI in include/grasp_b200.h)
  ops.py     torch-tensor wrappers (device pointers + current stream) over the C ABI
  engine.py  stage orchestration used by modeling_grasp.GRASPModel
  dist.py    matrix / sample partitioning and the NCCL exchange steps
  synth.py   random-init LLaMA configs and synthetic calibration tokens
There is no CPU or library fallback: every op raises if the CUDA library is
missing or a tensor is not on a CUDA device.
"""
__version__ = "0.1.0"
