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


def cfg1_inputs(rec_c64, x_circ, y_circ, dwnsmp=1):
    """The inputs ``fwi_script.py:31-85`` builds from the shipped ``RecordedData.mat`` (BASELINE configs[0]): float32 grid
    ``arange(-0.12, 0.12 + dxi, 0.8e-3)`` (301 points), nearest-node element snapping, +-31 element exclusion,
    ``a0 = 10``, ``L_PML = 9 mm``.  Returns a ``RingGeometry`` and the (Nt, E) complex64 data."""
    dxi, xmax = 0.8e-3, 120e-3
    xi = np.arange(-xmax, xmax + dxi, dxi, dtype=np.float32)  # fwi_script.py:46-50
    yi = xi.copy()
    xc = np.asarray(x_circ, dtype=np.float32).ravel()
    yc = np.asarray(y_circ, dtype=np.float32).ravel()
    x_idx, y_idx = G.snap_elements(xi, yi, xc, yc)  # :65-66
    ne = xc.size
    tx_include = np.arange(0, ne, dwnsmp)  # :35
    mask = G.build_masks(ne, tx_include, 31)  # :39-44, :79-85
    geom = G.RingGeometry(xi=xi, yi=yi, x_idx=x_idx, y_idx=y_idx, ind_matlab=x_idx * xi.size + y_idx, tx_include=tx_include,
                          mask_indices=mask, num_elements=ne, a0=10.0, L_PML=9.0e-3)
    return geom, np.asarray(rec_c64).astype(np.complex64)[tx_include, :]
