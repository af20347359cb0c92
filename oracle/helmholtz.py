"""Oracle restatement of ``Final_python/solve_helmholtz.py`` (TEST INFRASTRUCTURE ONLY).

Every function cites the reference lines it follows.  ``dtype="c64"`` reproduces
the reference's float32/complex64 arithmetic (JAX default, x64 disabled);
``dtype="c128"`` is the same algorithm in float64/complex128 and is the "truth"
that complex64 results are judged against.

The sparse solve is SciPy ``spsolve(csr_matrix(...), dense_rhs)`` exactly as
``solve_helmholtz.py:15-18`` does; SciPy is third-party (reference pins
scipy==1.15.2, this image has 1.18.1) -- see oracle/__init__.py, "parity unpinned".
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

_REAL = {"c64": np.float32, "c128": np.float64}
_CPLX = {"c64": np.complex64, "c128": np.complex128}


def stencil_opt_params(vmin, vmax, f, h, g, dtype="c128"):
    """Optimal 9-point stencil weights, ``solve_helmholtz.py:104-154``.

    b is fixed to 5/6 (``:141-143``); (d, e) solve the 2x2 normal equations of a
    1000 x 2 least-squares fit (``:144-147``).  ``dtype="c64"`` evaluates in float32
    like the reference does under JAX; "c128" in float64.
    """
    R = _REAL[dtype]
    vmin, vmax, f, h, g = (R(v) for v in (vmin, vmax, f, h, g))
    l, r = 100, 10  # :120-121
    Gmin = vmin / (f * h)
    Gmax = vmax / (f * h)
    m = np.arange(1, l + 1)
    n = np.arange(1, r + 1)
    theta = ((m - 1) * R(np.pi) / R(4 * (l - 1))).astype(R)  # :127
    G = (R(1) / (R(1) / Gmax + (n - 1).astype(R) / R(r - 1) * (R(1) / Gmin - R(1) / Gmax))).astype(R)  # :128
    TH, GG = np.meshgrid(theta, G)  # :131
    P = np.cos(g * R(2 * np.pi) * np.cos(TH) / GG)  # :133
    Q = np.cos(R(2 * np.pi) * np.sin(TH) / GG)  # :134
    S1 = (1 + 1 / g**2) * GG**2 * (1 - P - Q + P * Q)  # :136
    S2 = R(np.pi**2) * (2 - P - Q)  # :137
    S3 = R(2 * np.pi**2) * (1 - P * Q)  # :138
    S4 = R(2 * np.pi**2) + GG**2 * ((1 + 1 / g**2) * P * Q - P - Q / g**2)  # :139
    b = R(5.0 / 6.0)
    A = np.stack([S2.ravel(), S3.ravel()], axis=1).astype(R)
    y = (S4.ravel() - b * S1.ravel()).astype(R)
    params = np.linalg.solve(A.T @ A, A.T @ y)  # :146
    return float(b), float(params[0]), float(params[1])


def pml_profiles(x, y, a0, L_PML, dtype="c128"):
    """1-D half-grid PML stretch factors e_x, e_y (``solve_helmholtz.py:31-56``).

    The reference builds 2-D meshgrids; ``e_x`` varies only along x and ``e_y`` only
    along y, so the 1-D vectors carry everything.  f cancels (``:40-56``):
    e = 1 - i*a0*(max(|x-xc| - xspan + L, 0)/L)^2 with sign_convention = -1 (``:23``).
    Returns (ex[2Nx-1], ey[2Ny-1]).
    """
    R, C = _REAL[dtype], _CPLX[dtype]
    x = np.asarray(x, dtype=R)
    y = np.asarray(y, dtype=R)
    Nx, Ny = x.size, y.size
    xe = np.linspace(x[0], x[-1], 2 * (Nx - 1) + 1, dtype=R)
    ye = np.linspace(y[0], y[-1], 2 * (Ny - 1) + 1, dtype=R)
    xctr, xspan = (x[0] + x[-1]) / R(2), (x[-1] - x[0]) / R(2)
    yctr, yspan = (y[0] + y[-1]) / R(2), (y[-1] - y[0]) / R(2)
    L = R(L_PML)
    a0 = R(a0)
    px = (np.maximum(np.abs(xe - xctr) - xspan + L, R(0)) / L) ** 2
    py = (np.maximum(np.abs(ye - yctr) - yspan + L, R(0)) / L) ** 2
    ex = (R(1) - 1j * (a0 * px)).astype(C)
    ey = (R(1) - 1j * (a0 * py)).astype(C)
    return ex, ey


def _abc(ex, ey):
    """A=(ey/ex)[::2,1::2], B=(ex/ey)[1::2,::2], C=(ex*ey)[::2,::2] (``:58-60``)."""
    A = ey[::2][:, None] / ex[1::2][None, :]  # (Ny, Nx-1)
    B = ex[::2][None, :] / ey[1::2][:, None]  # (Ny-1, Nx)
    C = ey[::2][:, None] * ex[::2][None, :]  # (Ny, Nx)
    return A, B, C


def assemble_planes(Nx, Ny, g, b, d, e, h, A, B, C, k, stencil="python"):
    """Nine stencil-coefficient planes on the interior nodes.

    Follows ``assemble_Helmholtz`` (``solve_helmholtz.py:158-260``).  Returns a dict of
    (Ny-2, Nx-2) arrays keyed c,l,r,d,u,dl,dr,ul,ur.  ``stencil="python"`` reproduces
    the reference's out-of-bounds gathers ``A[., x+1]`` / ``B[y+1, .]`` with JAX's
    index clamping; ``"matlab"`` uses the MATLAB original's
    edge-adjacent coefficients (``solveHelmholtz.m:109,115,121``).
    """
    R = A.real.dtype.type
    g2 = R(g) ** 2
    h2 = R(h) ** 2
    b, d, e = R(b), R(d), R(e)
    beta = (R(1) - b) / R(2)
    ys = np.arange(1, Ny - 1)[:, None]
    xs = np.arange(1, Nx - 1)[None, :]

    def gA(yy, xx):  # JAX clamps out-of-range gather indices
        return A[np.clip(yy, 0, Ny - 1), np.clip(xx, 0, Nx - 2)]

    def gB(yy, xx):
        return B[np.clip(yy, 0, Ny - 2), np.clip(xx, 0, Nx - 1)]

    q = C * (k.astype(R) ** 2)

    A_c, A_l = gA(ys, xs), gA(ys, xs - 1)
    A_d, A_u = gA(ys - 1, xs), gA(ys + 1, xs)
    B_c, B_l, B_r = gB(ys, xs), gB(ys, xs - 1), gB(ys, xs + 1)
    B_d = gB(ys - 1, xs)
    A_dl, B_dl = gA(ys - 1, xs - 1), gB(ys - 1, xs - 1)
    B_dr = gB(ys - 1, xs + 1)
    A_ul = gA(ys + 1, xs - 1)
    if stencil == "python":
        A_dr = gA(ys - 1, xs + 1)  # :231 (clamped at x = Nx-2)
        B_ul = gB(ys + 1, xs - 1)  # :234 (clamped at y = Ny-2)
        A_ur = gA(ys + 1, xs + 1)  # :239
        B_ur = gB(ys + 1, xs + 1)  # :238
    elif stencil == "matlab":
        A_dr = gA(ys - 1, xs)
        B_ul = gB(ys, xs - 1)
        A_ur = gA(ys + 1, xs)
        B_ur = gB(ys, xs + 1)
    else:
        raise ValueError(stencil)

    def Q(dy, dx):
        return q[1 + dy:Ny - 1 + dy, 1 + dx:Nx - 1 + dx]

    P = {}
    P["c"] = (R(1) - d - e) * Q(0, 0) - b * (A_c + A_l + B_c / g2 + B_d / g2) / h2  # :242-244
    P["l"] = (b * A_l - beta * (B_l / g2 + B_dl / g2)) / h2 + (d / R(4)) * Q(0, -1)  # :245-247
    P["r"] = (b * A_c - beta * (B_r / g2 + B_dr / g2)) / h2 + (d / R(4)) * Q(0, 1)  # :248-250
    P["d"] = (b * B_d / g2 - beta * (A_d + A_dl)) / h2 + (d / R(4)) * Q(-1, 0)  # :251-253
    P["u"] = (b * B_c / g2 - beta * (A_u + A_ul)) / h2 + (d / R(4)) * Q(1, 0)  # :254-256
    P["dl"] = beta * (A_dl + B_dl / g2) / h2 + (e / R(4)) * Q(-1, -1)  # :257
    P["dr"] = beta * (A_dr + B_dr / g2) / h2 + (e / R(4)) * Q(-1, 1)  # :258
    P["ul"] = beta * (A_ul + B_ul / g2) / h2 + (e / R(4)) * Q(1, -1)  # :259
    P["ur"] = beta * (A_ur + B_ur / g2) / h2 + (e / R(4)) * Q(1, 1)  # :260
    return P


_OFFS = {"c": (0, 0), "l": (0, -1), "r": (0, 1), "d": (-1, 0), "u": (1, 0),
         "dl": (-1, -1), "dr": (-1, 1), "ul": (1, -1), "ur": (1, 1)}
PLANE_ORDER = ("c", "l", "r", "d", "u", "dl", "dr", "ul", "ur")  # :198-200 column order


def assemble_helmholtz(Nx, Ny, g, b, d, e, h, A, B, C, k, stencil="python"):
    """Sparse (N, N) CSR Helmholtz matrix, row-major unknowns ``y*Nx + x``.

    ``assemble_Helmholtz`` (``solve_helmholtz.py:158-290``): nine entries per interior
    row, identity on the Dirichlet ring (``:266-276``).
    """
    P = assemble_planes(Nx, Ny, g, b, d, e, h, A, B, C, k, stencil)
    ys, xs = np.meshgrid(np.arange(1, Ny - 1), np.arange(1, Nx - 1), indexing="ij")
    rows, cols, vals = [], [], []
    ctr = (ys * Nx + xs).ravel()
    for name in PLANE_ORDER:
        dy, dx = _OFFS[name]
        rows.append(ctr)
        cols.append(((ys + dy) * Nx + (xs + dx)).ravel())
        vals.append(P[name].ravel())
    x0 = np.arange(Nx)
    y0 = np.arange(1, Ny - 1)
    bdr = np.concatenate([x0, (Ny - 1) * Nx + x0, y0 * Nx, y0 * Nx + Nx - 1])  # :269-273
    rows.append(bdr)
    cols.append(bdr)
    vals.append(np.ones(bdr.size, dtype=A.dtype))
    H = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                      shape=(Nx * Ny, Nx * Ny))
    return H.tocsr()


def _setup(x, y, vel, f, a0, L_PML, dtype, bde, stencil):
    """Everything ``solve_helmholtz`` computes before the solve (``:23-64``)."""
    R, Cx = _REAL[dtype], _CPLX[dtype]
    x = np.asarray(x, dtype=R)
    y = np.asarray(y, dtype=R)
    vel = np.asarray(vel, dtype=R)
    f = R(f)
    h = np.mean(np.diff(x)).astype(R)  # :24
    gh = np.mean(np.diff(y)).astype(R)  # :25
    g = R(gh / h)  # :26
    Nx, Ny = x.size, y.size  # :27
    k = (R(2 * np.pi) * f / vel).astype(R)  # :28
    ex, ey = pml_profiles(x, y, a0, L_PML, dtype)
    A, B, C = _abc(ex, ey)
    if bde is None:
        bde = stencil_opt_params(vel.min(), vel.max(), f, h, g, dtype)  # :62
    b, d, e = bde
    H = assemble_helmholtz(Nx, Ny, g, b, d, e, h, A.astype(Cx), B.astype(Cx), C.astype(Cx), k, stencil)
    return H.astype(Cx), (Nx, Ny), bde


def solve_helmholtz(x, y, vel, src, f, a0, L_PML, adjoint, dtype="c64", bde=None, stencil="python"):
    """``solve_helmholtz(x, y, vel, src, f, a0, L_PML, adjoint)`` -> (Ny, Nx, nrhs).

    Restates ``solve_helmholtz.py:21-101``.  The adjoint system is conj(H)^T
    (``:66-73``); the right-hand side is ``src`` reshaped C-order to (Nx*Ny, nrhs)
    (``:78``); the solve is ``spsolve(csr, dense)`` (``:15-18``, ``:85-93``).
    ``bde`` injects (b, d, e) so that two precisions can share identical weights.
    """
    Cx = _CPLX[dtype]
    H, (Nx, Ny), _ = _setup(x, y, vel, f, a0, L_PML, dtype, bde, stencil)
    if adjoint:
        H = H.conj().T.tocsr()
    src = np.asarray(src)
    rhs = np.ascontiguousarray(src.reshape(Nx * Ny, -1).astype(Cx))
    sol = spla.spsolve(H, rhs)
    return np.asarray(sol, dtype=Cx).reshape(Ny, Nx, -1)


class HelmholtzFactor:
    """Oracle-side convenience: one SuperLU factorisation reused for forward and
    adjoint solves of the same operator.  Not how the reference does it (it calls
    ``spsolve`` -- i.e. refactorises -- three times per iteration,
    ``nonlinearcg.py:213,263,279``); mathematically the same solves.  Used by tests
    to keep the CPU suite fast; the timed ``cpu_baseline`` uses ``solve_helmholtz``.
    """

    def __init__(self, x, y, vel, f, a0, L_PML, dtype="c128", bde=None, stencil="python"):
        self.dtype = dtype
        H, (self.Nx, self.Ny), self.bde = _setup(x, y, vel, f, a0, L_PML, dtype, bde, stencil)
        self.H = H
        # SciPy's spsolve hands a CSR matrix to SuperLU as the CSC storage of H^T and solves the
        # transposed system; factorising H^T here mirrors that (in complex64 the CSC form of the
        # unscaled H -- identity Dirichlet rows next to ~1/h^2 interior rows -- loses all accuracy,
        # SURVEY.md section 0.6).
        self.lu = spla.splu(H.T.tocsc())

    def solve(self, src, adjoint=False, threads=1):
        """``threads`` > 1 splits the columns over a thread pool (SuperLU's triangular solves release the GIL); the
        columns are independent, so the result is the same -- used by the large-grid tests to stay within seconds."""
        Cx = _CPLX[self.dtype]
        rhs = np.ascontiguousarray(np.asarray(src).reshape(self.Nx * self.Ny, -1).astype(Cx))

        def block(r):
            if adjoint:  # conj(H)^T x = b  <=>  H^T conj(x) = conj(b)
                return np.conj(self.lu.solve(np.conj(r), trans="N"))
            return self.lu.solve(r, trans="T")

        ncol = rhs.shape[1]
        if threads > 1 and ncol > 1:
            from concurrent.futures import ThreadPoolExecutor
            edges = np.linspace(0, ncol, min(threads, ncol) + 1).astype(int)
            with ThreadPoolExecutor(len(edges) - 1) as ex:
                parts = list(ex.map(lambda i: block(np.ascontiguousarray(rhs[:, edges[i]:edges[i + 1]])), range(len(edges) - 1)))
            sol = np.concatenate(parts, axis=1)
        else:
            sol = block(rhs)
        return np.asarray(sol, dtype=Cx).reshape(self.Ny, self.Nx, -1)
