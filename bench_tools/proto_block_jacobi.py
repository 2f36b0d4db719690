"""Design prototype (numpy, CPU): block one-sided Jacobi SVD on row vectors.

Not product code and not the oracle -- it only fixes the algorithmic knobs
(block size, inner-sweep cap, ordering, convergence test) of the CUDA SVD in
grasp_b200/csrc/svd_jacobi.cu before any GPU time is spent.
"""
import sys, time
import numpy as np


def round_robin(p):
    """p even: list of p-1 rounds, each p/2 disjoint pairs (i<j)."""
    idx = list(range(p))
    rounds = []
    for _ in range(p - 1):
        pairs = []
        for t in range(p // 2):
            a, b = idx[t], idx[p - 1 - t]
            pairs.append((min(a, b), max(a, b)))
        rounds.append(pairs)
        idx = [idx[0]] + [idx[-1]] + idx[1:-1]
    return rounds


def jacobi_evd(G, max_sweeps, tol=1e-7, bipartite=False):
    """Parallel-order two-sided Jacobi on symmetric G (fp32). Returns E (cols = eigvecs), sweeps."""
    s = G.shape[0]
    G = G.astype(np.float32).copy()
    E = np.eye(s, dtype=np.float64)
    rr = round_robin(s)
    used = 0
    for sw in range(max_sweeps):
        d = np.sqrt(np.abs(np.diag(G)))
        off = np.abs(G - np.diag(np.diag(G))) / (np.outer(d, d) + 1e-30)
        if off.max() < tol:
            break
        used += 1
        for pairs in rr:
            J = np.eye(s, dtype=np.float32)
            J64 = np.eye(s, dtype=np.float64)
            for (p, q) in pairs:
                if bipartite and not (p < s // 2 <= q):
                    continue
                apq = float(G[p, q]); gpp = float(G[p, p]); gqq = float(G[q, q])
                if abs(apq) <= 1e-30 or abs(apq) < tol * np.sqrt(abs(gpp * gqq)):
                    continue
                tau = (gqq - gpp) / (2 * apq)
                t = np.sign(tau) / (abs(tau) + np.sqrt(1 + tau * tau)) if tau != 0 else 1.0
                c = 1 / np.sqrt(1 + t * t)
                sn = t * c
                J[p, p] = c; J[q, q] = c; J[p, q] = sn; J[q, p] = -sn
                J64[p, p] = c; J64[q, q] = c; J64[p, q] = sn; J64[q, p] = -sn
            G = J.T @ G @ J
            E = E @ J64
    return E.astype(np.float32), used, G


def block_jacobi_svd(A, b=32, inner_cap=30, max_sweeps=30, tol=1e-6, gemm=None, verbose=True):
    """Rows of Y are orthogonalised: Y = Q^T A. Returns U,S,Vh of A (A: r x L, r<=L)."""
    if gemm is None:
        gemm = lambda X, Y: X @ Y
    r, L = A.shape
    Y = A.astype(np.float32).copy()
    QT = np.eye(r, dtype=np.float32)
    p = r // b
    rr = round_robin(p)
    stats = []
    for sw in range(max_sweeps):
        maxoff = 0.0
        inner_total = 0
        for pairs in rr:
            for (I, J) in pairs:
                rows = np.r_[I * b:(I + 1) * b, J * b:(J + 1) * b]
                Yp = Y[rows]
                G = gemm(Yp, Yp.T)
                d = np.sqrt(np.abs(np.diag(G)))
                off = np.abs(G - np.diag(np.diag(G))) / (np.outer(d, d) + 1e-30)
                maxoff = max(maxoff, off.max())
                E, used, Gd = jacobi_evd(G, inner_cap, tol=tol)
                inner_total += used
                # order by descending diagonal so big rows migrate to low indices
                order = np.argsort(-np.diag(Gd), kind="stable")
                E = E[:, order]
                Y[rows] = gemm(E.T, Yp)
                QT[rows] = gemm(E.T, QT[rows])
        stats.append((sw, maxoff, inner_total))
        if verbose:
            print(f"sweep {sw}: max rel offdiag {maxoff:.3e} inner sweeps {inner_total}", flush=True)
        if maxoff < tol:
            break
    S = np.linalg.norm(Y.astype(np.float64), axis=1).astype(np.float32)
    order = np.argsort(-S, kind="stable")
    S = S[order]
    Vh = Y[order] / S[:, None]
    U = QT[order].T
    return U, S, Vh, stats


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    b = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    cap = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    rng = np.random.default_rng(0)
    A = (rng.standard_normal((n, n)) * 0.02).astype(np.float32)
    t = time.time()
    U, S, Vh, stats = block_jacobi_svd(A, b=b, inner_cap=cap)
    print("time", time.time() - t)
    Sref = np.linalg.svd(A.astype(np.float64), compute_uv=False)
    print("sigma err / smax", np.abs(S - Sref).max() / Sref[0])
    print("recon", np.linalg.norm((U * S) @ Vh - A) / np.linalg.norm(A))
    print("orthU", np.abs(U.T @ U - np.eye(n)).max(), "orthV", np.abs(Vh @ Vh.T - np.eye(n)).max())
