"""Oracle restatement of ``nonlinearcg.py`` (vectorised form) and
``fwi_loss_function.py`` (TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py).
"""
from __future__ import annotations

import numpy as np

from .helmholtz import _CPLX, _REAL, HelmholtzFactor, solve_helmholtz


def estimate_src_strength_batched(rec_sim, rec):
    """alpha_t = <sim_t, rec_t> / <sim_t, sim_t>, vdot conjugates its first argument
    (``nonlinearcg.py:14-20``, ``fwi_loss_function.py:18-26``)."""
    num = np.sum(np.conj(rec_sim) * rec, axis=1)
    den = np.sum(np.conj(rec_sim) * rec_sim, axis=1)
    return (num / den).astype(rec_sim.dtype)


def receiver_gather(WV, ind_matlab, mask_indices):
    """rec_sim[t, j] = WV.ravel(order='F' per source)[ind_matlab[mask_indices[t, j]], t]
    (``nonlinearcg.py:220-222``, ``fwi_loss_function.py:67-74``)."""
    N1, N2, Nt = WV.shape
    flat = np.transpose(WV, (1, 0, 2)).reshape(N1 * N2, Nt)
    global_inds = np.take(ind_matlab, mask_indices)  # (Nt, Nmask)
    return np.take_along_axis(flat.T, global_inds, axis=1), global_inds


def _solver(xi, yi, VEL, f, a0, L_PML, dtype, bde, stencil, reuse_factor, threads=1):
    """Returns solve(src, adjoint).  ``reuse_factor=False`` is the reference's behaviour
    (fresh ``spsolve`` per call); True shares one SuperLU factorisation."""
    if reuse_factor:
        fac = HelmholtzFactor(xi, yi, VEL, f, a0, L_PML, dtype=dtype, bde=bde, stencil=stencil)
        return lambda src, adjoint=False: fac.solve(src, adjoint, threads=threads)
    return lambda src, adjoint=False: solve_helmholtz(
        xi, yi, VEL, src, f, a0, L_PML, adjoint, dtype=dtype, bde=bde, stencil=stencil)


def fwi_loss_and_grad(params, xi, yi, REC_DATA, SRC, f, a0, L_PML, tx_include, ind_matlab,
                      mask_indices, num_elements, dtype="c64", bde=None, stencil="python",
                      reuse_factor=True, return_fields=False, threads=1):
    """(loss, grad) for one frequency.

    loss: ``fwi_loss_function.py:29-103``.  grad: the adjoint-state gradient with
    respect to slowness that ``nonlinearcg.py:243-265`` forms (residual scatter,
    virtual source ``2 w^2 s u``, adjoint solve, ``sum_t -Re(conj(VIRT) * ADJ_WV)``),
    returned with ``params``' shape.
    """
    R, Cx = _REAL[dtype], _CPLX[dtype]
    Nyi, Nxi = np.asarray(yi).size, np.asarray(xi).size
    Nt = np.asarray(tx_include).size
    params = np.asarray(params, dtype=R)
    SLOW = params.reshape(Nyi, Nxi)
    VEL = (R(1) / SLOW).astype(R)
    REC_DATA = np.asarray(REC_DATA).astype(Cx)
    solve = _solver(xi, yi, VEL, f, a0, L_PML, dtype, bde, stencil, reuse_factor, threads)

    WV = solve(SRC, False)  # fwi_loss_function.py:53
    rec_sim, global_inds = receiver_gather(WV, ind_matlab, mask_indices)
    rec = np.take_along_axis(REC_DATA, mask_indices, axis=1)
    SRC_EST = estimate_src_strength_batched(rec_sim, rec)  # :78
    WV = (WV * SRC_EST[None, None, :]).astype(Cx)  # :81
    rec_sim, _ = receiver_gather(WV, ind_matlab, mask_indices)  # :84-93
    diff = rec_sim - rec
    loss = R(0.5) * np.sum(np.abs(diff) ** 2)  # :102

    # adjoint source, nonlinearcg.py:248-254
    flat_adj = np.zeros((Nt, Nyi * Nxi), dtype=Cx)
    flat_adj[np.arange(Nt)[:, None], global_inds] = diff
    ADJ_SRC = np.transpose(flat_adj.reshape(Nt, Nxi, Nyi), (2, 1, 0))
    w = R(2 * np.pi) * R(f)
    VIRT = (R(2) * w**2) * SLOW[:, :, None] * WV  # :258
    ADJ_WV = solve(ADJ_SRC, True)  # :263
    grad = np.sum(-np.real(np.conj(VIRT) * ADJ_WV), axis=2)  # :264-265
    out = (float(loss), grad.astype(R).reshape(params.shape))
    if return_fields:
        return out + (dict(WV=WV, ADJ_WV=ADJ_WV, SRC_EST=SRC_EST, rec_sim=rec_sim, VIRT=VIRT),)
    return out


def fwi_loss_function(params, xi, yi, REC_DATA, SRC, f, a0, L_PML, tx_include, ind_matlab,
                      mask_indices, num_elements, dtype="c64", bde=None, stencil="python"):
    """The reference surface: scalar loss only (``fwi_loss_function.py:29-103``)."""
    R, Cx = _REAL[dtype], _CPLX[dtype]
    Nyi, Nxi = np.asarray(yi).size, np.asarray(xi).size
    SLOW = np.asarray(params, dtype=R).reshape(Nyi, Nxi)
    VEL = (R(1) / SLOW).astype(R)
    REC_DATA = np.asarray(REC_DATA).astype(Cx)
    WV = solve_helmholtz(xi, yi, VEL, SRC, f, a0, L_PML, False, dtype=dtype, bde=bde, stencil=stencil)
    rec_sim, _ = receiver_gather(WV, ind_matlab, mask_indices)
    rec = np.take_along_axis(REC_DATA, mask_indices, axis=1)
    SRC_EST = estimate_src_strength_batched(rec_sim, rec)
    WV = (WV * SRC_EST[None, None, :]).astype(Cx)
    rec_sim, _ = receiver_gather(WV, ind_matlab, mask_indices)
    return float(R(0.5) * np.sum(np.abs(rec_sim - rec) ** 2))


def nonlinear_conjugate_gradient_vectorized(xi, yi, numElements, REC_DATA, SRC, tx_include, ind_matlab,
                                            c_init, f, Niter, a0, L_PML, mask_indices, dtype="c64",
                                            bde=None, stencil="python", reuse_factor=True, history=None):
    """``nonlinear_conjugate_gradient_vectorized`` (``nonlinearcg.py:184-308``); the loop
    form (``:41-180``) computes the same quantities with unrolled Python loops.

    Returns (VEL, sd, grad, ADJ_WV, WV) like the reference.  ``history`` (a list) receives
    per-iteration dicts (loss, grad norm, beta, step, VEL range) for the known-answer tests.
    ``c_init`` may be a scalar (reference) or an (Ny, Nx) array.
    """
    R, Cx = _REAL[dtype], _CPLX[dtype]
    xi = np.asarray(xi, dtype=R)
    yi = np.asarray(yi, dtype=R)
    Nyi, Nxi = yi.size, xi.size
    REC_DATA = np.asarray(REC_DATA).astype(Cx)
    Nt = len(tx_include)
    VEL = (R(1) * np.asarray(c_init, dtype=R) * np.ones((Nyi, Nxi), dtype=R)).astype(R)  # :202
    SLOW = (R(1) / VEL).astype(R)
    sd = np.zeros((Nyi, Nxi), dtype=R)
    gprev = np.zeros((Nyi, Nxi), dtype=R)
    ADJ_WV = np.zeros((Nyi, Nxi, Nt), dtype=Cx)
    WV = np.zeros((Nyi, Nxi, Nt), dtype=Cx)
    batch_idx = np.arange(Nt)[:, None]
    w = R(2 * np.pi) * R(f)
    for it in range(Niter):
        solve = _solver(xi, yi, VEL, f, a0, L_PML, dtype, bde, stencil, reuse_factor)
        WV = solve(SRC, False)  # :213
        rec_sim, global_inds = receiver_gather(WV, ind_matlab, mask_indices)  # :220-222
        rec = np.take_along_axis(REC_DATA, mask_indices, axis=1)  # :223
        SRC_EST = estimate_src_strength_batched(rec_sim, rec)  # :224
        WV = (WV * SRC_EST[None, None, :]).astype(Cx)  # :227
        rec_sim, _ = receiver_gather(WV, ind_matlab, mask_indices)  # :230-239
        rec_obs = rec
        REC_SIM = np.zeros((Nt, numElements), dtype=Cx)  # :243-245
        REC_SIM[batch_idx, mask_indices] = rec_sim
        diff = rec_sim - rec_obs  # :248
        flat_adj = np.zeros((Nt, Nyi * Nxi), dtype=Cx)
        flat_adj[batch_idx, global_inds] = diff  # :249-250
        ADJ_SRC = np.transpose(flat_adj.reshape(Nt, Nxi, Nyi), (2, 1, 0))  # :253-254
        VIRT = ((R(2) * w**2) * SLOW[:, :, None] * WV).astype(Cx)  # :258
        ADJ_WV = solve(ADJ_SRC, True)  # :263
        grad = np.sum(-np.real(np.conj(VIRT) * ADJ_WV), axis=2).astype(R)  # :264-265
        dg = grad - gprev  # :268
        if it == 0:  # :274
            beta = R(0)
        else:
            beta = R(np.vdot(grad.ravel(order="F"), dg.ravel(order="F"))
                     / np.vdot(sd.ravel(order="F"), dg.ravel(order="F")))  # :270-272
        sd = (beta * sd - grad).astype(R)  # :276
        PERT = solve((-VIRT * sd[:, :, None]).astype(Cx), False)  # :279-281
        rec_vals, _ = receiver_gather(PERT, ind_matlab, mask_indices)  # :284-293
        dREC = np.zeros((Nt, numElements), dtype=Cx)
        dREC[batch_idx, mask_indices] = rec_vals  # :297-298
        num = np.real(np.vdot(dREC.ravel(order="F"), (REC_DATA - REC_SIM).ravel(order="F")))  # :24-26
        den = np.real(np.vdot(dREC.ravel(order="F"), dREC.ravel(order="F")))  # :27
        step = R(num / den)  # :28
        SLOW = (SLOW + step * sd).astype(R)  # :29
        VEL = (R(1) / SLOW).astype(R)  # :30
        if history is not None:
            history.append(dict(it=it, loss=float(0.5 * np.sum(np.abs(diff) ** 2)),
                                grad_norm=float(np.linalg.norm(grad)), beta=float(beta), step=float(step),
                                vel_min=float(VEL.min()), vel_max=float(VEL.max())))
        gprev = grad  # scan carry :303
    return VEL, sd, grad, ADJ_WV, WV
