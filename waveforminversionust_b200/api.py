"""Reference-facing surface: the same names, argument order and meaning as the reference's
``solve_helmholtz.py``, ``fwi_loss_function.py`` and ``nonlinearcg.py``, running on libustfwi.so.

    solve_helmholtz(x, y, vel, src, f, a0, L_PML, adjoint)            solve_helmholtz.py:21-101
    fwi_loss_function(params, xi, yi, REC_DATA, SRC, f, a0, L_PML,
                      tx_include, ind_matlab, mask_indices, num_elements) -> (loss, grad)
                                                                       fwi_loss_function.py:29-103
    nonlinear_conjugate_gradient[_vectorized](xi, yi, numElements, REC_DATA, SRC, tx_include,
                      ind_matlab, c_init, f, Niter, a0, L_PML, mask_indices)
                                                                       nonlinearcg.py:41-308

Inputs may be NumPy arrays (host buffers: copies happen inside the C call) or torch CUDA tensors
(device buffers: zero-copy).  There is no CPU fallback: without the CUDA library / a GPU these raise.
Extensions over the reference (keyword-only): ``dtype`` ("c64" default = the reference's precision,
or "c128"), ``engine`` ("auto" | "simt" | "tc2": block-GEMM engine; tcgen05 is complex64 only), ``bde`` (inject the stencil weights), ``stencil`` ("python" | "matlab"), and ``f`` /
``REC_DATA`` may carry a leading frequency axis for the joint multi-frequency objective.
"""
from __future__ import annotations

import hashlib

import numpy as np

from .plan import HelmholtzPlan

_PLANS = {}


def _is_torch(a):
    return type(a).__module__.startswith("torch")


def _to_np(a):
    if _is_torch(a):
        return a.detach().cpu().numpy()
    return np.asarray(a)


def _device_of(*arrs):
    for a in arrs:
        if _is_torch(a) and a.is_cuda:
            return a.device.index
    return 0


def get_plan(nx, ny, dtype, device, max_freq, max_nrhs, stencil, fwi_buffers, engine="auto"):
    """Plan cache: plans are reused (and grown) per (grid, precision, device, stencil, engine)."""
    key = (nx, ny, dtype, device, stencil, engine)
    p = _PLANS.get(key)
    if p is not None and (p.max_freq < max_freq or p.max_nrhs < max_nrhs or (fwi_buffers and not p.fwi_buffers)):
        max_freq, max_nrhs = max(max_freq, p.max_freq), max(max_nrhs, p.max_nrhs)
        fwi_buffers = fwi_buffers or p.fwi_buffers
        p.close()
        p = None
    if p is None:
        p = HelmholtzPlan(nx, ny, dtype=dtype, max_freq=max_freq, max_nrhs=max_nrhs, device=device,
                          stencil=stencil, fwi_buffers=fwi_buffers, engine=engine)
        p.fwi_buffers = bool(fwi_buffers)
        _PLANS[key] = p
    return p


def clear_plans():
    for p in _PLANS.values():
        p.close()
    _PLANS.clear()


def _fingerprint(*parts):
    h = hashlib.blake2b(digest_size=16)
    for q in parts:
        if isinstance(q, np.ndarray):
            h.update(np.ascontiguousarray(q).tobytes())
        else:
            h.update(repr(q).encode())
    return h.digest()


def solve_helmholtz(x, y, vel, src, f, a0, L_PML, adjoint, *, dtype="c64", bde=None, stencil="python", engine="auto"):
    """Solve H(vel, f) u = src (or conj(H)^T u = src when ``adjoint``) for every column of ``src``.

    ``src`` is (Ny, Nx, nrhs) (anything that reshapes C-order to (Ny*Nx, nrhs), solve_helmholtz.py:78);
    returns (Ny, Nx, nrhs) complex (solve_helmholtz.py:101).  The factorisation is cached and reused
    while (vel, f, grid, PML, weights) stay the same, so the forward / adjoint / perturbation solves of
    one iteration (nonlinearcg.py:213,263,279) factorise once instead of three times.
    """
    xh, yh = _to_np(x).astype(np.float64).ravel(), _to_np(y).astype(np.float64).ravel()
    nx, ny = xh.size, yh.size
    f = float(np.asarray(_to_np(f)).reshape(-1)[0])
    adjoint = bool(np.asarray(_to_np(adjoint)).reshape(-1)[0]) if not isinstance(adjoint, bool) else adjoint
    nrhs = int(np.prod(src.shape)) // (nx * ny)
    dev = _device_of(src, vel)
    plan = get_plan(nx, ny, dtype, dev, 1, nrhs, stencil, False, engine)
    plan.set_grid(xh, yh, float(a0), float(L_PML))
    if _is_torch(src) and src.is_cuda:
        import torch
        velt = vel if _is_torch(vel) else torch.as_tensor(_to_np(vel))
        velt = velt.to(device=src.device, dtype=plan.treal).contiguous()
        key = _fingerprint(velt.cpu().numpy(), f, None if bde is None else np.asarray(bde, dtype=np.float64))
        if plan._factor_key != key:
            plan.factor(velt, [f], bde=None if bde is None else [bde])
            plan._factor_key = key
        out = src.to(plan.tcplx).reshape(ny * nx, nrhs).contiguous().clone()
        plan.solve(out, 0, adjoint)
        return out.reshape(ny, nx, nrhs)
    velh = _to_np(vel).astype(plan.real)
    key = _fingerprint(velh, f, None if bde is None else np.asarray(bde, dtype=np.float64))
    refactor = plan._factor_key != key
    out = plan.solve_helmholtz_host(velh, _to_np(src), f, adjoint=adjoint, bde=bde, refactor=refactor)
    plan._factor_key = key
    return out


def _acquisition(SRC, ind_matlab, mask_indices, nx, ny):
    """Convert the reference's arrays to what the C ABI takes: one-hot source nodes and row-major
    receiver nodes.  ``ind_matlab`` indexes the column-major (order='F') flattening of a (Ny, Nx) field
    (nonlinearcg.py:220-222): p = x*Ny + y."""
    ind = _to_np(ind_matlab).astype(np.int64).ravel()
    xr, yr = ind // ny, ind % ny
    rx_lin = (yr * nx + xr).astype(np.int32)
    mask = _to_np(mask_indices).astype(np.int32)
    if hasattr(SRC, "src_lin"):
        src_lin = np.asarray(SRC.src_lin, dtype=np.int32)
    else:
        S = _to_np(SRC)
        nt = S.shape[2]
        flat = S.reshape(ny * nx, nt)
        nz_row, nz_col = np.nonzero(flat)
        if nz_col.size != nt or not np.array_equal(np.sort(nz_col), np.arange(nt)) or not np.allclose(flat[nz_row, nz_col], 1.0):
            raise NotImplementedError("fwi_loss_function: SRC must be one-hot unit sources (fwi_script.py:72-74)")
        src_lin = np.empty(nt, dtype=np.int32)
        src_lin[nz_col] = nz_row
    return src_lin, rx_lin, mask


class OneHotSources:
    """Sparse stand-in for the reference's dense (Ny, Nx, Nt) one-hot ``SRC`` array."""

    def __init__(self, src_lin, shape):
        self.src_lin = np.asarray(src_lin, dtype=np.int32)
        self.shape = tuple(shape)


def _fwi_plan(xi, yi, REC_DATA, SRC, f, a0, L_PML, ind_matlab, mask_indices, dtype, stencil, device, engine="auto"):
    xh, yh = _to_np(xi).astype(np.float64).ravel(), _to_np(yi).astype(np.float64).ravel()
    nx, ny = xh.size, yh.size
    freqs = np.atleast_1d(np.asarray(_to_np(f), dtype=np.float64)).ravel()
    src_lin, rx_lin, mask = _acquisition(SRC, ind_matlab, mask_indices, nx, ny)
    plan = get_plan(nx, ny, dtype, device, freqs.size, src_lin.size, stencil, True, engine)
    plan.set_grid(xh, yh, float(a0), float(L_PML))
    plan.set_acquisition(src_lin, rx_lin, mask)
    return plan, freqs


def fwi_loss_function(params, xi, yi, REC_DATA, SRC, f, a0, L_PML, tx_include, ind_matlab, mask_indices,
                      num_elements, *, dtype="c64", bde=None, stencil="python", engine="auto"):
    """(loss, grad): loss of fwi_loss_function.py:29-103 and the adjoint-state gradient with respect to
    the slowness ``params`` (nonlinearcg.py:243-265), ``grad.shape == params.shape``.

    With ``f`` of length Nf and ``REC_DATA`` of shape (Nf, Nt, E) the joint objective sum_f loss_f is
    returned.  Usable as ``jaxopt.LBFGS(fun, value_and_grad=True)`` / ``scipy.optimize.minimize(jac=True)``.
    """
    dev = _device_of(params, REC_DATA)
    plan, freqs = _fwi_plan(xi, yi, REC_DATA, SRC, f, a0, L_PML, ind_matlab, mask_indices, dtype, stencil, dev, engine)
    if _is_torch(params) and params.is_cuda:
        rec = REC_DATA.to(plan.tcplx).reshape(freqs.size, plan.nt, plan.nelem).contiguous()
        slow = params.to(plan.treal).reshape(plan.ny, plan.nx).contiguous()
        loss, grad = plan.fwi_loss_grad(slow, rec, freqs, bde=bde)
        return loss[0], grad.reshape(params.shape)
    p = _to_np(params)
    loss, grad = plan.fwi_loss_grad_host(p.reshape(plan.ny, plan.nx), _to_np(REC_DATA), freqs, bde=bde)
    return loss, grad.reshape(p.shape)


def nonlinear_conjugate_gradient(xi, yi, numElements, REC_DATA, SRC, tx_include, ind_matlab, c_init, f, Niter,
                                 a0, L_PML, mask_indices, *, dtype="c64", bde=None, stencil="python",
                                 device=0, history=None, return_fields=True, engine="auto"):
    """The reference's NCG loop (nonlinearcg.py:41-180 / 184-308: Hestenes-Stiefel beta, forced to 0 at
    the first iteration; linearised exact step) on the GPU path: per iteration one factorisation,
    forward + adjoint + perturbation sweeps, all on device.

    Returns (VEL, sd, grad, ADJ_WV, WV) as NumPy arrays like the reference; ADJ_WV / WV (each
    (Ny, Nx, Nt) complex, WV scaled by the source estimates, nonlinearcg.py:227) are only materialised
    when ``return_fields`` and refer to the last iteration / first frequency.
    """
    import torch
    plan, freqs = _fwi_plan(xi, yi, REC_DATA, SRC, f, a0, L_PML, ind_matlab, mask_indices, dtype, stencil, device, engine)
    dv = torch.device(f"cuda:{device}")
    ny, nx = plan.ny, plan.nx
    rec = torch.as_tensor(_to_np(REC_DATA)).to(device=dv, dtype=plan.tcplx).reshape(freqs.size, plan.nt, plan.nelem).contiguous()
    VEL = (torch.as_tensor(np.asarray(_to_np(c_init), dtype=np.float64)) * torch.ones((ny, nx), dtype=torch.float64)).to(dv, plan.treal)
    SLOW = (1.0 / VEL).contiguous()
    sd = torch.zeros((ny, nx), dtype=plan.treal, device=dv)
    gprev = torch.zeros_like(sd)
    grad = torch.zeros_like(sd)
    for it in range(int(Niter)):
        loss, grad = plan.fwi_loss_grad(SLOW, rec, freqs, bde=bde)  # nonlinearcg.py:213-265
        dg = grad - gprev  # :268
        if it == 0:  # :274
            beta = torch.zeros((), dtype=plan.treal, device=dv)
        else:
            beta = torch.sum(grad * dg) / torch.sum(sd * dg)  # :270-272
        sd = (beta * sd - grad).contiguous()  # :276
        if return_fields and it == int(Niter) - 1:
            ADJ_WV = plan.adjoint_wavefield(0)  # before the perturbation solve reuses the buffer
        nd = plan.ncg_linesearch(sd)  # :279-298, :24-27
        step = (nd[0] / nd[1]).to(plan.treal)  # :28
        SLOW = (SLOW + step * sd).contiguous()  # :29
        VEL = 1.0 / SLOW  # :30
        if history is not None:
            history.append(dict(it=it, loss=float(loss[0]), grad_norm=float(torch.linalg.norm(grad.double())),
                                beta=float(beta), step=float(step), vel_min=float(VEL.min()), vel_max=float(VEL.max())))
        gprev = grad
    out_adj = out_wv = None
    if return_fields and int(Niter) > 0:
        alpha = torch.as_tensor(plan.src_est(0)).to(dv)
        out_wv = (plan.wavefield(0) * alpha[None, None, :]).cpu().numpy()
        out_adj = ADJ_WV.cpu().numpy()
    return VEL.cpu().numpy(), sd.cpu().numpy(), grad.cpu().numpy(), out_adj, out_wv


nonlinear_conjugate_gradient_vectorized = nonlinear_conjugate_gradient


def run_lbfgs_fwi(xi, yi, REC_DATA, SRC, tx_include, ind_matlab, c_init, f, a0, L_PML, mask_indices, *, maxiter=1, tol=1e-5,
                  history_size=10, dtype="c64", bde=None, stencil="python", engine="auto", history=None, loss_grad=None,
                  pert_scale=1e-2):
    """``run_lbfgs_fwi`` of the reference (``fwi_loss_function.py:106-132``): L-BFGS on the slowness map, returning the
    final sound speed ``(Ny, Nx)``.  The reference wires ``jaxopt.LBFGS(fun=loss_fn, maxiter=1, tol=1e-5)`` around a
    loss-only function (which JAX cannot differentiate through ``pure_callback``); here the objective is this package's
    ``fwi_loss_function -> (loss, grad)`` (the ``value_and_grad=True`` contract: ``grad.shape == params.shape``, also for the
    2-D ``init_params`` of ``:110-111``) and the optimiser is SciPy's L-BFGS (jaxopt is not installable in this image; with
    jaxopt present use ``jaxopt.LBFGS(fun, value_and_grad=True, jit=False)`` on the same function).

    Deliberate deviation: in the reference's units ``|grad| ~ 3e-11`` and useful steps are ``~3e7``, so
    ``tol=1e-5`` would stop at iteration 0.  The problem is non-dimensionalised: the unknown is the relative slowness
    perturbation ``p = (s / s0 - 1) / pert_scale`` (a unit step is a ``pert_scale`` = 1 % change) and the objective
    ``loss / loss(s0)``; ``tol`` applies to the projected gradient of that scaled problem.

    ``loss_grad(slow_2d) -> (loss, grad_2d)`` replaces the objective (the parity test drives the same optimiser with the
    oracle's); ``history`` receives ``(loss, params.copy())`` of every evaluation.
    """
    from scipy.optimize import minimize
    ny, nx = _to_np(yi).size, _to_np(xi).size
    num_elements = int(SRC.shape[2]) if SRC is not None else 0  # unused when ``loss_grad`` is injected
    c0 = np.asarray(_to_np(c_init), dtype=np.float64)
    s0 = 1.0 / float(c0.mean())
    real = np.float32 if dtype == "c64" else np.float64
    state = {"loss0": None}
    if loss_grad is None:
        def loss_grad(slow):
            return fwi_loss_function(slow.astype(real), xi, yi, REC_DATA, SRC, f, a0, L_PML, tx_include, ind_matlab, mask_indices,
                                     num_elements, dtype=dtype, bde=bde, stencil=stencil, engine=engine)

    def fun(p):
        slow = (1.0 + pert_scale * p.reshape(ny, nx)) * s0
        loss, grad = loss_grad(slow)
        if np.shape(grad) != slow.shape:
            raise ValueError("loss_grad must return a gradient with the shape of its argument")
        if state["loss0"] is None:
            state["loss0"] = float(loss)
        if history is not None:
            history.append((float(loss), slow.copy()))
        scale = 1.0 / state["loss0"]
        return float(loss) * scale, np.asarray(grad, dtype=np.float64).ravel() * (s0 * pert_scale * scale)

    p0 = ((np.ones((ny, nx)) / (c0 * s0) - 1.0) / pert_scale).ravel()
    res = minimize(fun, p0, jac=True, method="L-BFGS-B", options=dict(maxiter=int(maxiter), maxcor=int(history_size), gtol=float(tol)))
    final_slow = (1.0 + pert_scale * res.x.reshape(ny, nx)) * s0
    return 1.0 / final_slow


# -------------------------------------------------------------------------------------------------
# time-domain synthesis and the low-to-high frequency schedule
# -------------------------------------------------------------------------------------------------
def hanning(n):
    """MATLAB ``hanning(n)`` (symmetric Hann window without the zero end points), the frequency response of
    ``Lecture19_Fwi/TimeDomainSimulation.m:33``."""
    k = np.arange(1, int(n) + 1, dtype=np.float64)
    return 0.5 * (1.0 - np.cos(2.0 * np.pi * k / (int(n) + 1)))


def idtft(WVFIELD_F, f, resp_freq, time, df=None, *, device=0):
    """Inverse discrete-time Fourier transform of a stack of frequency-domain wavefields on an arbitrary time axis
    (``TimeDomainSimulation.m:54-56``; not an inverse FFT): ``WVFIELD_F`` (Ny, Nx, Nf) complex -> (Ny, Nx, Nt) with
    ``out[..., t] = sum_k exp(2j*pi*f[k]*time[t]) * df * resp_freq[k] * WVFIELD_F[..., k]`` (``ust_idtft``).

    A torch CUDA tensor in gives a torch CUDA tensor out (a permuted view of the (Nt, Ny*Nx) result); NumPy in, NumPy out.
    """
    import ctypes as C
    import torch
    from . import _lib
    f = np.ascontiguousarray(np.asarray(_to_np(f), dtype=np.float64).ravel())
    tm = np.ascontiguousarray(np.asarray(_to_np(time), dtype=np.float64).ravel())
    resp = np.ascontiguousarray(np.asarray(_to_np(resp_freq), dtype=np.float64).ravel())
    if resp.size != f.size:
        raise ValueError("resp_freq and f must have the same length")
    if df is None:
        df = float(f[1] - f[0]) if f.size > 1 else 1.0
    on_dev = _is_torch(WVFIELD_F) and WVFIELD_F.is_cuda
    W = WVFIELD_F if on_dev else torch.as_tensor(_to_np(WVFIELD_F)).to(f"cuda:{device}")
    if W.dtype not in (torch.complex64, torch.complex128):
        W = W.to(torch.complex64)
    ny, nx, nf = W.shape
    if nf != f.size:
        raise ValueError("last axis of WVFIELD_F must match f")
    U = W.permute(2, 0, 1).reshape(nf, ny * nx).contiguous()  # frequency-major stack
    out = torch.empty((tm.size, ny * nx), dtype=U.dtype, device=U.device)
    pd = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    with torch.cuda.device(U.device):
        rc = _lib.lib().ust_idtft(0 if U.dtype == torch.complex64 else 1, C.c_void_p(U.data_ptr()), nf, ny * nx, pd(f), pd(resp),
                                  float(df), pd(tm), tm.size, C.c_void_p(out.data_ptr()),
                                  C.c_void_p(torch.cuda.current_stream(U.device).cuda_stream))
    _lib.check(rc, "ust_idtft")
    res = out.reshape(tm.size, ny, nx).permute(1, 2, 0)
    return res if on_dev else res.cpu().numpy()


def time_domain_simulation(xi, yi, C_map, src, f, resp_freq, time, a0, L_PML, *, dtype="c64", stencil="python",
                           engine="auto", device=0, batch=16):
    """``Lecture19_Fwi/TimeDomainSimulation.m:36-56`` on the GPU path: one forward solve per frequency for the source
    ``src`` (Ny, Nx) of one transmitting element, ``batch`` frequencies factorised per launch sequence, then the inverse
    discrete-time Fourier transform to the time axis ``time``.  Returns NumPy ``(WVFIELD_F (Ny, Nx, Nf), WVFIELD_T (Ny, Nx, Nt))``.
    """
    import torch
    xh, yh = _to_np(xi).astype(np.float64).ravel(), _to_np(yi).astype(np.float64).ravel()
    nx, ny = xh.size, yh.size
    f = np.asarray(_to_np(f), dtype=np.float64).ravel()
    batch = max(1, min(int(batch), f.size))
    plan = get_plan(nx, ny, dtype, device, batch, 1, stencil, False, engine)
    plan.set_grid(xh, yh, float(a0), float(L_PML))
    plan._factor_key = None
    dv = torch.device(f"cuda:{device}")
    vel = torch.as_tensor(_to_np(C_map)).to(device=dv, dtype=plan.treal).reshape(ny, nx).contiguous()
    rhs0 = torch.as_tensor(_to_np(src)).to(device=dv, dtype=plan.tcplx).reshape(ny * nx, 1).contiguous()
    U = torch.empty((f.size, ny * nx), dtype=plan.tcplx, device=dv)
    for k0 in range(0, f.size, batch):
        fb = f[k0:k0 + batch]
        plan.factor(vel, fb)
        for i in range(fb.size):
            X = rhs0.clone()
            plan.solve(X, i, False)
            U[k0 + i] = X[:, 0]
    WF = U.reshape(f.size, ny, nx).permute(1, 2, 0)
    WT = idtft(WF, f, resp_freq, time)
    return WF.cpu().numpy(), WT.cpu().numpy()


def channel_data(WVFIELD_T, x_idx, y_idx):
    """``channelData(t, e) = WVFIELD_T(y_idx(e), x_idx(e), t)`` (``TimeDomainSimulation.m:72-75``; 0-based indices)."""
    W = _to_np(WVFIELD_T)
    return W[np.asarray(y_idx), np.asarray(x_idx), :].T


def continuation_stages(f, nstages):
    """Low-to-high frequency schedule: ``nstages`` contiguous groups of ascending frequency (index arrays into ``f``).
    The reference states the rule only (``SimulateData.m:29-30`` "Use 0.1-0.6 MHz ... cycle skipping"; slides p.24)."""
    order = np.argsort(np.asarray(_to_np(f), dtype=np.float64), kind="stable")
    return [np.sort(g) for g in np.array_split(order, int(nstages)) if g.size]


def frequency_continuation(xi, yi, numElements, REC_DATA, SRC, tx_include, ind_matlab, c_init, f, stages, Niter,
                           a0, L_PML, mask_indices, *, dtype="c64", stencil="python", device=0, engine="auto", history=None):
    """Multi-frequency FWI by frequency continuation: for every stage (an index array into ``f``, see
    ``continuation_stages``) ``Niter`` NCG iterations (``nonlinear_conjugate_gradient``) on the joint objective of that
    stage's frequencies, each stage starting from the sound speed the previous one ended with.  ``REC_DATA`` is
    (Nf, Nt, E).  Returns the final sound speed (Ny, Nx); ``history`` (a list) receives one list of per-iteration
    records per stage.
    """
    f = np.asarray(_to_np(f), dtype=np.float64).ravel()
    rec = _to_np(REC_DATA)
    if rec.ndim != 3 or rec.shape[0] != f.size:
        raise ValueError("REC_DATA must be (Nf, Nt, E) with Nf == len(f)")
    vel = c_init
    prev = None
    for st in stages:
        idx = np.asarray(st, dtype=np.int64).ravel()
        if prev is not None and prev != idx.size:
            clear_plans()  # a plan is sized for its number of frequencies; do not keep two factor stores alive
        prev = idx.size
        h = [] if history is not None else None
        vel, _, _, _, _ = nonlinear_conjugate_gradient(xi, yi, numElements, rec[idx], SRC, tx_include, ind_matlab, vel, f[idx], Niter,
                                                       a0, L_PML, mask_indices, dtype=dtype, stencil=stencil, device=device,
                                                       history=h, return_fields=False, engine=engine)
        if history is not None:
            history.append(h)
    return vel
