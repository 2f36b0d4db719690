"""GPU: BASELINE.json's full sizes (LLaMA-2-7B matrices) checked through size-independent properties --
the CPU oracle would need minutes per matrix there."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def factors(cuda):
    """One 4096x11008 (down_proj-shaped) and one 4096x4096 matrix factored in ONE batched call."""
    from grasp_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(7)
    mats = [torch.randn(4096, 11008, device=cuda, generator=g) * 0.02,
            torch.randn(11008, 4096, device=cuda, generator=g) * 0.02,
            torch.randn(4096, 4096, device=cuda, generator=g) * 0.02]
    outs, info = ops.svd_batched(mats, return_info=True)
    return mats, outs, info.cpu()


def test_full_size_svd_properties(factors):
    mats, outs, info = factors
    assert torch.all(info[:, 1] == 1), f"SVD did not converge: {info}"
    for A, (U, S, Vh) in zip(mats, outs):
        r = min(A.shape)
        assert U.shape == (A.shape[0], r) and Vh.shape == (r, A.shape[1])
        assert torch.all(S[:-1] >= S[1:]) and torch.all(S >= 0)
        rec = (torch.linalg.norm((U * S) @ Vh - A) / torch.linalg.norm(A)).item()
        assert rec < 1e-5, rec                                              # LAPACK sgesdd: 2.5e-6 .. 3.3e-6
        eye = torch.eye(r, device=A.device)
        assert (U.T @ U - eye).abs().max().item() < 2e-5
        assert (Vh @ Vh.T - eye).abs().max().item() < 2e-5
        # Frobenius norm is the l2 norm of the singular values (checksum of checksums)
        assert abs(S.double().pow(2).sum().sqrt().item() / torch.linalg.norm(A.double()).item() - 1) < 1e-6
        # Marchenko-Pastur edge of a 0.02-scaled Gaussian matrix: sigma_max ~ 0.02 (sqrt(m) + sqrt(n))
        edge = 0.02 * (A.shape[0] ** 0.5 + A.shape[1] ** 0.5)
        assert abs(S[0].item() / edge - 1) < 0.02


def test_full_size_score_select_rebuild_properties(factors, cuda):
    from grasp_b200 import ops
    mats, outs, _ = factors
    A, (U, S, Vh) = mats[0], outs[0]
    out_f, in_f = A.shape
    r = S.numel()
    gen = torch.Generator(device="cuda").manual_seed(8)
    G1 = torch.randn(out_f, in_f, device=cuda, generator=gen)
    G2 = torch.randn(out_f, in_f, device=cuda, generator=gen)
    g1, s1 = ops.sigma_score(U, G1, Vh, S)
    g2, _ = ops.sigma_score(U, G2, Vh, S)
    g12, _ = ops.sigma_score(U, G1 + G2, Vh, S)
    assert ((g12 - (g1 + g2)).abs().max() / g12.abs().max()).item() < 1e-5          # linear in G
    gacc, _ = ops.sigma_score(U, G2, Vh, S, dsigma=g1.clone(), want_score=False)
    assert ((gacc - g12).abs().max() / g12.abs().max()).item() < 1e-5               # accumulation == sum
    # G = W itself: u_i^T W v_i = sigma_i
    gw, _ = ops.sigma_score(U, A, Vh, S, metric="gradient")
    assert ((gw - S).abs().max() / S[0]).item() < 1e-5
    assert torch.equal(s1, (g1 * S).abs())

    k = int(in_f * out_f * 0.1 / (in_f + out_f))
    assert k == 298
    idx = ops.topk(s1, k)
    vals = s1[idx]
    assert torch.all(vals[:-1] >= vals[1:]) and len(set(idx.tolist())) == k          # sorted, distinct
    mask = torch.ones(r, dtype=torch.bool, device=cuda); mask[idx] = False
    assert vals[-1] >= s1[mask].max()                                                # nothing better was left out

    W = ops.lowrank_rebuild(U, S, Vh, idx)
    # rank-k rebuild: its own factors reproduce it, and it is the projection of A onto the kept triplets
    proj = (U[:, idx] * S[idx]) @ Vh[idx]
    assert (torch.linalg.norm(W - proj) / torch.linalg.norm(proj)).item() < 1e-5
    in_w, out_w = ops.factor_pack(U, S, Vh, idx)
    assert (torch.linalg.norm(out_w @ in_w - proj) / torch.linalg.norm(proj)).item() < 1e-5
    full = ops.lowrank_rebuild(U, S, Vh, torch.arange(r, device=cuda))
    assert (torch.linalg.norm(full - A) / torch.linalg.norm(A)).item() < 1e-5          # k = r gives W back
    Wb = ops.lowrank_rebuild(U, S, Vh, idx, out_dtype=torch.bfloat16)
    assert (torch.linalg.norm(Wb.float() - W.to(torch.bfloat16).float()) / torch.linalg.norm(W)).item() < 1e-3
