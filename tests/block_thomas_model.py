"""NumPy model of the GPU algorithm (two-sided block-Thomas with explicit block inverses
computed by an UNPIVOTED blocked Gauss-Jordan).  Used by CPU tests to validate the
numerics of the algorithm the CUDA kernels implement, independent of any GPU.
Not product code and not the oracle."""
import numpy as np


def gj_inverse_blocked(S, nb=64, pivot_in_block=True):
    """In-place-equivalent blocked Gauss-Jordan inverse without inter-block pivoting."""
    n = S.shape[0]
    X = S.copy()
    for k0 in range(0, n, nb):
        k1 = min(k0 + nb, n)
        Akk = X[k0:k1, k0:k1]
        P = np.linalg.inv(Akk) if pivot_in_block else _gj_unpivoted(Akk)
        P = P.astype(S.dtype)
        Xt = X.copy()
        Xt[:, k0:k1] = 0
        Xt[k0:k1, k0:k1] = np.eye(k1 - k0, dtype=S.dtype)
        Rk = (P @ Xt[k0:k1, :]).astype(S.dtype)
        Xn = (Xt - X[:, k0:k1] @ Rk).astype(S.dtype)
        Xn[k0:k1, :] = Rk
        X = Xn
    return X


def _gj_unpivoted(A):
    n = A.shape[0]
    X = A.copy()
    for k in range(n):
        p = 1 / X[k, k]
        row = X[k, :] * p
        row[k] = p
        col = X[:, k].copy()
        X[:, k] = 0
        X = X - np.outer(col, row)
        X[k, :] = row
    return X


def tri_from_planes(P, row, names):
    """Tridiagonal block (sub, dia, sup) for interior grid row `row` from planes."""
    lo, di, hi = (P[nm][row] for nm in names)
    return lo, di, hi


def tri_dense(lo, di, hi):
    n = di.size
    K = np.zeros((n, n), dtype=di.dtype)
    K[np.arange(n), np.arange(n)] = di
    K[np.arange(1, n), np.arange(0, n - 1)] = lo[1:]
    K[np.arange(0, n - 1), np.arange(1, n)] = hi[:-1]
    return K


class TwoSidedBlockThomas:
    def __init__(self, P, nb=64, inv=None):
        """P: dict of (M, n) planes (interior nodes)."""
        self.P = P
        M, n = P["c"].shape
        self.M, self.n = M, n
        self.m = M // 2
        dt = P["c"].dtype
        inv = inv or (lambda S: gj_inverse_blocked(S, nb))
        L = lambda i: tri_dense(P["dl"][i], P["d"][i], P["dr"][i])
        U = lambda i: tri_dense(P["ul"][i], P["u"][i], P["ur"][i])
        D = lambda i: tri_dense(P["l"][i], P["c"][i], P["r"][i])
        self.L, self.U = L, U
        T = [None] * M
        for i in range(0, self.m):
            S = D(i) if i == 0 else D(i) - L(i) @ T[i - 1] @ U(i - 1)
            T[i] = inv(S.astype(dt))
        for i in range(M - 1, self.m, -1):
            S = D(i) if i == M - 1 else D(i) - U(i) @ T[i + 1] @ L(i + 1)
            T[i] = inv(S.astype(dt))
        m = self.m
        S = D(m)
        if m > 0:
            S = S - L(m) @ T[m - 1] @ U(m - 1)
        if m < M - 1:
            S = S - U(m) @ T[m + 1] @ L(m + 1)
        T[m] = inv(S.astype(dt))
        self.T = T

    def solve(self, b, adjoint=False):
        """b: (M, n, nrhs)."""
        M, m, T = self.M, self.m, self.T
        dt = b.dtype
        H = lambda A: A.conj().T
        if not adjoint:
            opT = lambda i: T[i]
            lower = lambda i: self.L(i)      # couples row i to i-1
            upper = lambda i: self.U(i)      # couples row i to i+1
        else:
            opT = lambda i: H(T[i])
            lower = lambda i: H(self.U(i - 1))
            upper = lambda i: H(self.L(i + 1))
        z = np.zeros_like(b)
        for i in range(0, m):
            w = b[i] if i == 0 else b[i] - lower(i) @ z[i - 1]
            z[i] = (opT(i) @ w).astype(dt)
        for i in range(M - 1, m, -1):
            w = b[i] if i == M - 1 else b[i] - upper(i) @ z[i + 1]
            z[i] = (opT(i) @ w).astype(dt)
        w = b[m].copy()
        if m > 0:
            w = w - lower(m) @ z[m - 1]
        if m < M - 1:
            w = w - upper(m) @ z[m + 1]
        x = z
        x[m] = (opT(m) @ w).astype(dt)
        for i in range(m - 1, -1, -1):
            x[i] = (z[i] - opT(i) @ (upper(i) @ x[i + 1])).astype(dt)
        for i in range(m + 1, M):
            x[i] = (z[i] - opT(i) @ (lower(i) @ x[i - 1])).astype(dt)
        return x
