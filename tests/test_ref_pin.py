"""The oracle against fixtures produced by EXECUTING the reference's own source (tests/golden/make_ref_golden.py:
the unmodified Final_python/*.py on a NumPy-backed ``jax`` stand-in, SciPy SuperLU as in the reference).

``x64`` fixtures (JAX with x64 enabled: float64 assembly + SuperLU, the reference's explicit complex64 casts kept) sit ~1e-7
from exact arithmetic and pin the algorithm; ``x32`` fixtures (the reference's default single precision) pin its actual
outputs to within the complex64 SuperLU noise floor (forward ~1e-5, adjoint / gradient ~1e-3).
"""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from common import cfg1_inputs, rel, small_case
from oracle import fwi as ofwi
from oracle import helmholtz as oh

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MODE = {"x32": ("c64", np.float32), "x64": ("c128", np.float64)}


def _load(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def _case(G):
    n = int(G["n"])
    geom, f, vel = small_case(n, int(G["nelem"]), seed=int(G["seed"]), pml_cells=float(G["pml_cells"]))
    assert f == pytest.approx(float(G["f"]), rel=1e-12)
    return n, geom, f, vel


@pytest.mark.parametrize("mode", ["x32", "x64"])
def test_stencil_weights_match_the_executed_reference(mode):
    G = _load("ref_solve_" + mode)
    n, geom, f, vel = _case(G)
    dt, R = MODE[mode]
    h = np.mean(np.diff(geom.xi.astype(R)))
    bde = oh.stencil_opt_params(R(vel).min(), R(vel).max(), R(f), h, R(1), dt)
    # float32 normal equations of a 1000 x 2 fit: ~5e-5 scatter between summation orders; float64: exact
    assert np.allclose(bde, G["bde"], rtol=0, atol=2e-5 if mode == "x32" else 1e-13)


@pytest.mark.parametrize("mode", ["x32", "x64"])
def test_assembled_matrix_is_the_one_the_reference_hands_to_superlu(mode):
    """CSR (data, indices, indptr) captured at the reference's ``scipy_solve`` call, forward and adjoint."""
    G = _load("ref_solve_" + mode)
    n, geom, f, vel = _case(G)
    dt, R = MODE[mode]
    H, _, _ = oh._setup(geom.xi, geom.yi, vel.astype(R), f, geom.a0, geom.L_PML, dt, tuple(G["bde"]), "python")
    Href = sp.csr_matrix((G["csr_data"], G["csr_indices"], G["csr_indptr"]), shape=(n * n, n * n))
    assert np.array_equal(H.indptr, Href.indptr) and np.array_equal(H.indices, Href.indices)  # same pattern, same order
    tol = 3e-7 if mode == "x32" else 1e-15
    assert abs(H - Href).max() / abs(Href).max() < tol
    Hadj = sp.csr_matrix((G["csr_adj_data"], G["csr_adj_indices"], G["csr_adj_indptr"]), shape=(n * n, n * n))
    assert abs(H.conj().T.tocsr() - Hadj).max() / abs(Href).max() < tol


@pytest.mark.parametrize("mode", ["x32", "x64"])
def test_wavefields_match_the_executed_reference(mode):
    G = _load("ref_solve_" + mode)
    n, geom, f, vel = _case(G)
    dt, R = MODE[mode]
    onehot = geom.dense_src()[:, :, :int(G["nrhs"])]
    for tag, src in (("onehot", onehot), ("dense", G["dense_src"])):
        for adj in (False, True):
            u = oh.solve_helmholtz(geom.xi, geom.yi, vel.astype(R), src, f, geom.a0, geom.L_PML, adj, dtype=dt,
                                   bde=tuple(G["bde"]) if mode == "x32" else None)
            ref = G["wv_%s_%s" % (tag, "adj" if adj else "fwd")]
            assert ref.dtype == np.complex64 and ref.shape == (n, n, int(G["nrhs"]))
            # x64: only the reference's final complex64 cast separates the two; x32: complex64 SuperLU noise floor
            tol = 1e-7 if mode == "x64" else (2e-5 if not adj else 5e-3)
            assert rel(u, ref) < tol, (tag, adj, rel(u, ref))


@pytest.mark.parametrize("mode", ["x32", "x64"])
def test_ncg_iterations_match_the_executed_reference(mode):
    """nonlinear_conjugate_gradient_vectorized, 1 and 2 iterations (loss, gradient, search direction, updated sound
    speed, wavefields), and the loop form the script calls."""
    G = _load("ref_ncg_" + mode)
    n, geom, f, vel = _case(G)
    dt, R = MODE[mode]
    rec = G["rec"].astype(np.complex64)  # fwi_script.py:26
    assert np.array_equal(G["grad1_loop"], G["grad1"]) or rel(G["grad1_loop"], G["grad1"]) < 1e-6
    tg, tv, tw, ta = (5e-6, 1e-4, 2e-7, 2e-6) if mode == "x64" else (2e-2, 0.1, 2e-4, 5e-3)
    for it, sfx in ((1, "1"), (2, "")):
        h = []
        VEL, sd, grad, ADJ, WV = ofwi.nonlinear_conjugate_gradient_vectorized(
            geom.xi, geom.yi, geom.num_elements, rec, geom.dense_src(), geom.tx_include, geom.ind_matlab, 1480.0, f, it,
            geom.a0, geom.L_PML, geom.mask_indices, dtype=dt, reuse_factor=False, history=h)
        assert rel(grad, G["grad" + sfx]) < tg and rel(sd, G["sd" + sfx]) < tg
        assert np.sqrt(np.mean((VEL - G["VEL" + sfx]) ** 2)) < tv  # m/s
        assert rel(WV[:, :, :2], G["WV" + sfx]) < tw and rel(ADJ[:, :, :2], G["ADJ_WV" + sfx]) < ta
    loss = ofwi.fwi_loss_function(np.full((n, n), 1 / 1480.0), geom.xi, geom.yi, rec, geom.dense_src(), f, geom.a0, geom.L_PML,
                                  geom.tx_include, geom.ind_matlab, geom.mask_indices, geom.num_elements, dtype=dt)
    assert abs(loss - float(G["loss0"])) / float(G["loss0"]) < (1e-6 if mode == "x64" else 2e-4)
    assert abs(h[0]["loss"] - float(G["loss0"])) / float(G["loss0"]) < (1e-6 if mode == "x64" else 2e-4)


def test_cfg1_script_run_matches_the_oracle_fixture():
    """BASELINE configs[0]: ``fwi_script.main()`` executed on the shipped RecordedData.mat (single precision, loop-form NCG,
    1 iteration) against the complex128 oracle results stored in cfg1_shipped.npz (made by make_cfg1_golden.py)."""
    S = _load("ref_script_cfg1")
    g = _load("cfg1_shipped")
    geom, rec = cfg1_inputs(g["rec"], g["x_circ"], g["y_circ"])
    assert np.array_equal(S["xi"], geom.xi) and np.array_equal(S["ind_matlab"], geom.ind_matlab)  # fwi_script.py:46-68
    assert np.array_equal(S["mask_indices"], geom.mask_indices) and float(S["f"]) == float(g["f"])
    assert (float(S["a0"]), float(S["L_PML"]), float(S["c_init"])) == (geom.a0, geom.L_PML, 1480.0)
    # complex64 reference vs complex128 oracle: SURVEY C gives 6.3e-3 gradient rel-L2, 0.047 m/s RMS after one iteration
    assert abs(float(S["grad_norm"]) - float(g["grad_norm0"])) / float(g["grad_norm0"]) < 2e-3
    assert rel(S["grad_dec2"][::2, ::2], g["grad0_dec4"]) < 2e-2
    assert np.sqrt(np.mean((S["VEL_dec2"][::2, ::2] - g["vel1_dec4"]) ** 2)) < 0.1
    assert abs(float(S["vel_min"]) - float(g["vel_min1"])) < 0.5 and abs(float(S["vel_max"]) - float(g["vel_max1"])) < 0.5
