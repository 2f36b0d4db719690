"""GPU parity tests of each kernel, called through the C ABI (grasp_b200.ops -> libgrasp_b200.so),
against the CPU oracle and the committed reference outputs."""
import math

import pytest
import torch

from oracle import restate

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = torch.as_tensor(a, dtype=torch.float64).cpu(), torch.as_tensor(b, dtype=torch.float64).cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


# ------------------------------------------------------------------------------- BI
def test_bi_matches_reference_fixture(cuda, golden):
    from grasp_b200 import ops
    import tools.utils_func as uf
    fx = golden("bi_small.pt")
    hs = [h.to(cuda) for h in fx["hiddens"]]
    acc = torch.zeros(len(hs) - 1, dtype=torch.float64, device=cuda)
    ops.bi_chain(hs, acc)
    ops.bi_chain(hs, acc)  # accumulates like the per-batch += of the reference
    assert rel(acc / 2, fx["means"]) < 1e-5          # tolerance: 1e-5 relative on fp32 inputs
    for i in range(len(hs) - 1):
        got = uf.block_influence(hs[i], hs[i + 1])
        assert (got.cpu() - fx["per_pair"][i]).abs().max().item() < 1e-5
        ang = uf.block_influence(hs[i][:, -1:], hs[i + 1][:, -1:], angular=True)
        assert (ang.cpu() - fx["angular_last_token"][i]).abs().max().item() < 1e-5
    assert abs(uf.block_influence(hs[1], hs[2])[3].item() - 0.5) < 1e-6   # zero row -> NaN -> 0.5


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("d", [4096, 1000, 8200])
def test_bi_chain_shapes_and_dtypes(cuda, dtype, d):
    from grasp_b200 import ops
    g = torch.Generator().manual_seed(d)
    hs = [torch.randn(1, 77, d, generator=g).to(dtype) for _ in range(5)]
    want = [restate.block_influence(hs[i].float(), hs[i + 1].float()).mean().item() for i in range(4)]
    acc = torch.zeros(4, dtype=torch.float64, device=cuda)
    ops.bi_chain([h.to(cuda) for h in hs], acc)
    assert rel(acc, want) < 2e-5


def test_bi_empty_and_full_size(cuda):
    from grasp_b200 import ops
    acc = torch.zeros(1, dtype=torch.float64, device=cuda)
    e = torch.empty(1, 0, 64, device=cuda)
    ops.bi_accumulate(e, e, acc)
    assert acc.item() == 0.0
    # LLaMA-2-7B shape: 33 states of [1, 511, 4096]; property: BI(x, x) == 0, BI(x, -x) == 2, BI(x, 2x) == 0
    x = torch.randn(1, 511, 4096, device=cuda)
    hs = [x, x.clone(), -x, 2 * x] + [torch.randn(1, 511, 4096, device=cuda) for _ in range(29)]
    acc = torch.zeros(32, dtype=torch.float64, device=cuda)
    ops.bi_chain(hs, acc)
    a = acc.cpu()
    assert abs(a[0]) < 1e-6 and abs(a[1] - 2) < 1e-6 and abs(a[2] - 2) < 1e-6
    assert all(abs(v - 1) < 0.01 for v in a[3:].tolist())   # independent gaussians are ~orthogonal


# ---------------------------------------------------------------------------- top-k
@pytest.mark.parametrize("r,k", [(4096, 204), (4096, 298), (8192, 637), (64, 22), (100, 100), (1, 1), (5000, 1),
                                 (65536, 300), (4096, 0)])
def test_topk_matches_torch(cuda, r, k):
    from grasp_b200 import ops
    g = torch.Generator().manual_seed(r + k)
    s = torch.rand(r, generator=g).abs()
    got = ops.topk(s.to(cuda), k).cpu()
    want = torch.topk(s, k).indices
    assert got.dtype == torch.int64 and got.shape == (k,)
    assert torch.equal(s[got], s[want])            # same scores in the same (descending) order
    assert len(set(got.tolist())) == k


def test_topk_ties_nan_negative_and_batch(cuda):
    from grasp_b200 import ops
    s = torch.tensor([1.0, 3.0, 3.0, -2.0, float("nan"), 3.0, 0.0, -0.0, float("inf"), 1.0])
    got = ops.topk(s.to(cuda), 6).tolist()
    assert got == [4, 8, 1, 2, 5, 0]               # NaN first (torch semantics), ties by lower index
    a, b = torch.randn(300), torch.randn(7000)
    ia, ib = ops.topk_batched([a.to(cuda), b.to(cuda)], [17, 300])
    assert torch.equal(a[ia.cpu()], torch.topk(a, 17).values)
    assert torch.equal(b[ib.cpu()], torch.topk(b, 300).values)
    many = [torch.randn(50 + i) for i in range(11)]   # > 8 matrices: more than one launch group
    outs = ops.topk_batched([m.to(cuda) for m in many], [5] * 11)
    for m, o in zip(many, outs):
        assert torch.equal(m[o.cpu()], torch.topk(m, 5).values)


def test_adaptive_rank_matches_reference(cuda, golden):
    import tools.utils_func as uf
    for case in golden("select_small.pt"):
        thr = case["threshold_0.6"]
        assert uf.adaptive_rank_selection(thr["score"].to(cuda), 0.6) == thr["idx"].tolist()


# ------------------------------------------------------------------------------ SVD
def check_svd(A, U, S, Vh, ref_S, tol_sigma=1e-5, tol_rec=1e-5, tol_orth=2e-5, full_rank=True):
    A, U, S, Vh = A.double().cpu(), U.double().cpu(), S.double().cpu(), Vh.double().cpu()
    r = min(A.shape)
    assert U.shape == (A.shape[0], r) and S.shape == (r,) and Vh.shape == (r, A.shape[1])
    assert torch.all(S >= 0) and torch.all(S[:-1] >= S[1:]), "S must be non-negative and descending"
    smax = ref_S[0].double()
    err = ((S - ref_S.double()).abs().max() / smax).item()
    assert err < tol_sigma, f"max|sigma - ref|/sigma_max = {err:.3e}"   # north star bar: 1e-4
    big = ref_S.double() >= 1e-3 * smax
    per = ((S - ref_S.double()).abs()[big] / ref_S.double()[big]).max().item()
    assert per < 1e-4, f"per-value relative sigma error {per:.3e}"
    rec = (torch.linalg.norm((U * S) @ Vh - A) / torch.linalg.norm(A)).item()
    assert rec < tol_rec, f"reconstruction {rec:.3e}"
    if full_rank:
        eye = torch.eye(r, dtype=torch.float64)
        ou = (U.T @ U - eye).abs().max().item()
        ov = (Vh @ Vh.T - eye).abs().max().item()
        assert ou < tol_orth and ov < tol_orth, f"orthogonality U {ou:.3e} V {ov:.3e}"


def test_svd_matches_reference_fixture(cuda, golden):
    from grasp_b200 import ops
    for case in golden("svd_small.pt"):
        A = case["A"]
        U, S, Vh = ops.svd(A.to(cuda))
        note = case.get("note", "")
        if note.startswith("rank"):
            check_svd(A, U, S, Vh, case["S"], full_rank=False)
        elif note.startswith("graded"):
            check_svd(A, U, S, Vh, case["S"], tol_orth=1e-4)
        else:
            check_svd(A, U, S, Vh, case["S"])


@pytest.mark.parametrize("m,n", [(256, 256), (704, 256), (256, 704), (512, 2048), (300, 200)])
def test_svd_random_weights(cuda, m, n):
    from grasp_b200 import ops
    g = torch.Generator().manual_seed(m * 7 + n)
    A = torch.randn(m, n, generator=g) * 0.02
    (U, S, Vh), = ops.svd_batched([A.to(cuda)])
    check_svd(A, U, S, Vh, restate.svd(A)[1])


@pytest.mark.parametrize("m,n", [(768, 2304), (2304, 768), (520, 1500), (1030, 520)])
def test_svd_preconditioned_wide_and_tall(cuda, m, n):
    """Wide / tall matrices go through the CholeskyQR2 reduction to a square factor; the factors must be as good as
    those of the direct route (precondition=False) and both must match the oracle."""
    from grasp_b200 import ops
    g = torch.Generator().manual_seed(m * 3 + n)
    A = torch.randn(m, n, generator=g) * 0.02
    S_ref = restate.svd(A)[1]
    outs, info = ops.svd_batched([A.to(cuda), A.to(cuda)], return_info=True)           # two of a kind: one launch group
    assert torch.all(info.cpu()[:, 1] == 1), info
    for U, S, Vh in outs:
        check_svd(A, U, S, Vh, S_ref)
    (U2, S2, Vh2), = ops.svd_batched([A.to(cuda)], precondition=False)
    check_svd(A, U2, S2, Vh2, S_ref)
    assert ((outs[0][1] - S2).abs().max() / S2[0]).item() < 2e-6


def _graded(r, L, decades, seed):
    g = torch.Generator().manual_seed(seed)
    Uo, _ = torch.linalg.qr(torch.randn(r, r, generator=g, dtype=torch.float64))
    Vo, _ = torch.linalg.qr(torch.randn(L, r, generator=g, dtype=torch.float64))
    sv = torch.logspace(0, -decades, r, dtype=torch.float64)
    return ((Uo * sv) @ Vo.T).float()


def test_svd_preconditioning_on_ill_conditioned_input(cuda):
    """cond(A) = 1e6: the shifted first Cholesky pass keeps the preconditioning sound (cond(A A^T) is beyond fp32).
    cond(A) = 1e10: it cannot be; the call reports that in info[1] (never silently wrong factors) and
    engine.batched_svd factors the matrix again without preconditioning."""
    from grasp_b200 import engine, ops

    def check(A, U, S, Vh):
        A64, S64 = A.double(), torch.linalg.svdvals(A.double())
        assert ((S.double().cpu() - S64).abs().max() / S64[0]).item() < 1e-5
        rec = (torch.linalg.norm((U.double().cpu() * S.double().cpu()) @ Vh.double().cpu() - A64) / torch.linalg.norm(A64)).item()
        assert rec < 1e-5, rec

    A6 = _graded(512, 1536, 6, 5)
    (usv,), info = ops.svd_batched([A6.to(cuda)], return_info=True)
    if int(info.cpu()[0, 1]):                         # reported sound -> it must be accurate
        check(A6, *usv)
    check(A6, *engine.batched_svd([A6.to(cuda)])[0])
    A10 = _graded(512, 1536, 10, 6)
    _, info = ops.svd_batched([A10.to(cuda)], return_info=True)
    assert int(info.cpu()[0, 1]) == 0, "an unsound preconditioning must be reported"
    check(A10, *engine.batched_svd([A10.to(cuda)])[0])


def test_svd_batched_mixed_shapes_and_info(cuda):
    from grasp_b200 import ops
    g = torch.Generator().manual_seed(11)
    mats = [torch.randn(m, n, generator=g) * 0.02 for (m, n) in [(128, 128), (352, 128), (128, 128), (128, 352),
                                                                 (128, 128)]]
    outs, info = ops.svd_batched([a.to(cuda) for a in mats], return_info=True)
    for A, (U, S, Vh) in zip(mats, outs):
        check_svd(A, U, S, Vh, restate.svd(A)[1])
    info = info.cpu()
    assert torch.all(info[:, 1] == 1), f"not converged: {info}"
    assert torch.all(info[:, 0] <= 24)


def test_svd_edge_cases(cuda):
    from grasp_b200 import ops
    Z = torch.zeros(40, 24)
    U, S, Vh = ops.svd(Z.to(cuda))
    assert torch.all(S == 0) and torch.isfinite(U).all() and torch.isfinite(Vh).all()
    I = torch.eye(70)
    U, S, Vh = ops.svd(I.to(cuda))
    assert (S.cpu() - 1).abs().max() < 1e-6
    assert (((U * S) @ Vh).cpu() - I).abs().max() < 1e-5
    D = torch.diag(torch.tensor([5.0, 1.0, 3.0, 0.0, 2.0]))
    U, S, Vh = ops.svd(D.to(cuda))
    assert torch.allclose(S.cpu(), torch.tensor([5.0, 3.0, 2.0, 1.0, 0.0]), atol=1e-6)


# ------------------------------------------------------------------- score / select
def test_sigma_score_matches_reference_fixture(cuda, golden):
    from grasp_b200 import ops
    for case in golden("select_small.pt"):
        U, S, Vh, G = (case[k].to(cuda) for k in ("U", "S", "Vh", "G"))
        for metric in ("taylor", "gradient"):
            g, sc = ops.sigma_score(U, G, Vh, S, metric=metric)
            assert rel(g, case["grad"]) < 2e-5       # tolerance: fp32 accumulation over out*in terms
            ref = case[f"{metric}_0.9"]
            assert rel(sc, ref["score"]) < 2e-5
        # accumulation over two calls == one call on 2G  (linearity in G)
        g1, _ = ops.sigma_score(U, G, Vh, S, want_score=False)
        g2, _ = ops.sigma_score(U, G, Vh, S, dsigma=g1.clone(), want_score=False)
        assert rel(g2, 2 * g1) < 1e-6
        assert rel(ops.score_from_grad(g1, S, "taylor"), (g1 * S).abs()) < 1e-6


def test_sigma_score_llama_shapes(cuda):
    from grasp_b200 import ops
    g = torch.Generator().manual_seed(5)
    for (o, i) in [(512, 512), (1376, 512), (512, 1376), (200, 328)]:
        r = min(o, i)
        U = torch.linalg.qr(torch.randn(o, r, generator=g))[0]
        Vh = torch.linalg.qr(torch.randn(i, r, generator=g))[0].T.contiguous()
        G = torch.randn(o, i, generator=g)
        S = torch.rand(r, generator=g)
        want = restate.sigma_grad_from_G(U.double(), G.double(), Vh.double())
        got, sc = ops.sigma_score(U.to(cuda), G.to(cuda), Vh.to(cuda), S.to(cuda))
        assert rel(got, want) < 2e-5
        assert rel(sc, (want * S.double()).abs()) < 2e-5


# -------------------------------------------------------------------------- compile
def test_rebuild_and_pack_match_reference_fixture(cuda, golden):
    from grasp_b200 import ops
    for case in golden("select_small.pt"):
        U, S, Vh = (case[k].to(cuda) for k in ("U", "S", "Vh"))
        for key in ("taylor_0.9", "taylor_0.5", "gradient_0.5"):
            ref = case[key]
            idx = ref["idx"].to(cuda)
            W = ops.lowrank_rebuild(U, S, Vh, idx)
            err = (torch.linalg.norm(W.cpu() - ref["merged"]) / torch.linalg.norm(ref["merged"])).item()
            assert err < 1e-6, err                     # bar: 1e-3 relative Frobenius (fp32 output)
            Wb = ops.lowrank_rebuild(U, S, Vh, idx, out_dtype=torch.bfloat16)
            assert torch.equal(Wb.cpu(), W.cpu().to(torch.bfloat16)) or \
                rel(Wb.float(), ref["merged"].to(torch.bfloat16).float()) < 8e-3
            in_w, out_w = ops.factor_pack(U, S, Vh, idx)
            # one gather, one correctly rounded sqrt, one multiply per element: at most 1 ulp from the reference
            assert rel(in_w, ref["in_w"]) < 2e-7, ("in_w", rel(in_w, ref["in_w"]))
            assert rel(out_w, ref["out_w"]) < 2e-7, ("out_w", rel(out_w, ref["out_w"]))
        # k = r reproduces the matrix
        r = S.numel()
        W = ops.lowrank_rebuild(U, S, Vh, torch.arange(r, device=cuda))
        assert (torch.linalg.norm(W.cpu() - case["W"]) / torch.linalg.norm(case["W"])).item() < 1e-5


@pytest.mark.parametrize("prec,tol", [(0, 1e-5), (16, 2e-6), (6, 2e-6), (3, 5e-5)])
def test_gemm_all_transposes(cuda, prec, tol):
    """grasp_gemm_f32 in every arithmetic (CUDA cores, fp16x3, bf16x6, bf16x3) against fp64."""
    from grasp_b200 import ops
    g = torch.Generator().manual_seed(9)
    for (M, N, K) in [(64, 64, 64), (130, 70, 33), (511, 300, 257), (1, 1, 1), (128, 384, 4096)]:
        for ta in (False, True):
            for tb in (False, True):
                A = torch.randn((K, M) if ta else (M, K), generator=g)
                B = torch.randn((N, K) if tb else (K, N), generator=g)
                want = (A.T if ta else A).double() @ (B.T if tb else B).double()
                got = ops.gemm(A.to(cuda), B.to(cuda), ta=ta, tb=tb, prec=prec)
                err = (torch.linalg.norm(got.double().cpu() - want) / torch.linalg.norm(want)).item()
                assert err < tol, (M, N, K, ta, tb, err)
                C0 = torch.randn(M, N, generator=g)
                got2 = ops.gemm(A.to(cuda), B.to(cuda), ta=ta, tb=tb, alpha=0.5, beta=2.0, C_out=C0.to(cuda), prec=prec)
                want2 = 0.5 * want + 2.0 * C0.double()
                denom = torch.linalg.norm(0.5 * want) + torch.linalg.norm(2.0 * C0.double())   # no cancellation blow-up
                assert (torch.linalg.norm(got2.double().cpu() - want2) / denom).item() < tol, (M, N, K, ta, tb)


def test_gemm_fp16_planes_survive_badly_scaled_rows(cuda):
    """Rows spanning 12 decades (gradients next to activations): the per-row power-of-two scaling keeps
    every row inside fp16 range; error measured per element against |a_m| |b_n|."""
    from grasp_b200 import ops
    g = torch.Generator().manual_seed(10)
    A = torch.randn(200, 333, generator=g) * torch.logspace(-9, 3, 200)[:, None]
    B = torch.randn(150, 333, generator=g) * torch.logspace(-6, 2, 150)[:, None]
    want = A.double() @ B.double().T
    scale = A.double().norm(dim=1)[:, None] * B.double().norm(dim=1)[None, :]
    for b_given_kn in (False, True):
        Bd = B.T.contiguous().to(cuda) if b_given_kn else B.to(cuda)
        got = ops.gemm(A.to(cuda), Bd, tb=not b_given_kn, prec=16)
        assert torch.isfinite(got).all()
        assert ((got.double().cpu() - want).abs() / scale).max().item() < 2e-6
    Z = torch.zeros(64, 64, device=cuda)
    assert torch.all(ops.gemm(Z, Z, prec=16) == 0)
