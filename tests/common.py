"""Shared builders for the parity tests (seeded inputs for both the oracle and the CUDA path)."""
import numpy as np

from oracle import helmholtz as oh
from waveforminversionust_b200 import geometry as G


def rel(a, b):
    a = np.asarray(a); b = np.asarray(b)
    return float(np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel()))


def small_case(n=48, nelem=32, seed=0, dwnsmp=1, lr=None, contrast=60.0, pml_cells=11.25):
    """Ring-array problem on an n x n grid with a smooth random model."""
    geom = G.ring_geometry(n, nelem, dwnsmp=dwnsmp, num_elem_lr=lr, pml_cells=pml_cells)
    f = G.frequency_for_grid(n)
    vel = G.blob_model(geom, dc=contrast, seed=seed + 7)
    return geom, f, vel


def bde_for(geom, vel, f):
    h = float(np.mean(np.diff(geom.xi.astype(np.float64))))
    return oh.stencil_opt_params(float(vel.min()), float(vel.max()), f, h, 1.0, "c128")


def observed_data(geom, f, vel_true, bde=None, seed=1234):
    """REC_DATA[t, e] = amplitude_t * u_t(element e), SimulateData.m:26,55-59 (complex128 oracle)."""
    fac = oh.HelmholtzFactor(geom.xi, geom.yi, vel_true, f, geom.a0, geom.L_PML, "c128", bde=bde)
    WV = fac.solve(geom.dense_src(np.complex128))
    amp = G.source_amplitudes(geom.tx_include.size, seed)
    rec = WV[geom.y_idx, geom.x_idx, :].T * amp[:, None]  # (Nt, E)
    return rec
