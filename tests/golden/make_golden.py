"""Generates tests/golden/*.npz from the oracle (complex128).  Run from the repo root:
    python tests/golden/make_golden.py
The fixtures pin the oracle against regressions and give the GPU tests a committed, box-independent
set of expected outputs (the reference itself ships no expected outputs; see oracle/__init__.py)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from common import bde_for, observed_data, small_case  # noqa: E402
from oracle import fwi as ofwi  # noqa: E402
from oracle import helmholtz as oh  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
PML_CELLS = 4.0  # keeps the ring elements outside the absorbing layer on these tiny grids


def make(name, n, nelem, stencil="python", seed=0):
    geom, f, vel_true = small_case(n, nelem, seed=seed, pml_cells=PML_CELLS)
    c0 = np.full((n, n), 1480.0)
    bde = bde_for(geom, c0, f)
    rec = observed_data(geom, f, vel_true, bde=bde_for(geom, vel_true, f))
    slow = 1.0 / c0
    args = (geom.xi, geom.yi, rec, geom.dense_src(), f, geom.a0, geom.L_PML, geom.tx_include, geom.ind_matlab,
            geom.mask_indices, geom.num_elements)
    loss, grad, fl = ofwi.fwi_loss_and_grad(slow, *args, dtype="c128", bde=bde, stencil=stencil, return_fields=True)
    # heterogeneous-model wavefields (forward and adjoint) for 4 sources, full field
    src4 = geom.dense_src(np.complex128)[:, :, :4]
    fac = oh.HelmholtzFactor(geom.xi, geom.yi, vel_true, f, geom.a0, geom.L_PML, "c128", bde=bde, stencil=stencil)
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"), n=n, nelem=nelem, f=f, pml_cells=PML_CELLS, bde=np.array(bde), seed=seed, stencil=stencil,
        vel_true=vel_true, rec=rec, loss=loss, grad=grad, src_est=fl["SRC_EST"],
        wv_fwd=fac.solve(src4, False), wv_adj=fac.solve(src4, True))
    print(name, "loss", loss, "|grad|", np.linalg.norm(grad))


if __name__ == "__main__":
    make("ring40_python", 40, 16)
    make("ring56_matlab", 56, 32, stencil="matlab", seed=3)
