"""Numeric helpers of the GRASP hot path, same names and call signatures as the
reference's tools/utils_func.py, computed by the sm_100a kernels of grasp_b200.

  block_influence          reference tools/utils_func.py:3-25
  jaccard_similarity       reference tools/utils_func.py:28-42 (pure set logic, host side)
  adaptive_rank_selection  reference tools/utils_func.py:45-57
"""
import torch

from grasp_b200 import ops


def block_influence(input_hidden_state: torch.Tensor, output_hidden_state: torch.Tensor, angular=False):
    """Per-token 1 - cos(in, out) (or arccos(cos)/pi) over [B, S, D] hidden states -> [B*S] fp32.

    NaN similarities (a zero-norm token) count as 0.5, as in the reference."""
    if input_hidden_state.dim() != 3 or output_hidden_state.dim() != 3:
        raise ValueError("hidden states must be [B, S, D]")
    return ops.bi_accumulate(input_hidden_state, output_hidden_state, acc=None, angular=bool(angular), per_row=True)


def jaccard_similarity(list1, list2):
    a = set(list1) if isinstance(list1, list) else list1
    b = set(list2) if isinstance(list2, list) else list2
    union = a | b
    return len(a & b) / len(union) if len(union) > 0 else 0


def adaptive_rank_selection(svd_importance_list, target_ratio):
    """Indices, by descending importance, of the shortest prefix reaching target_ratio of the total."""
    if not torch.is_tensor(svd_importance_list):
        raise TypeError("adaptive_rank_selection expects a CUDA tensor of importances")
    return ops.adaptive_rank(svd_importance_list, float(target_ratio)).tolist()
