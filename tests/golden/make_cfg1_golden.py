"""Fixture for BASELINE configs[0] (the reference's shipped ring-array dataset).  Run HERE (needs /root/reference):
    python tests/golden/make_cfg1_golden.py
Reads Final_python/RecordedData.mat with the repo's own MAT-v7.3 reader (mat73/h5py are not installable), keeps only what
fwi_script.py consumes (REC_DATA as complex64, element coordinates, f -- the 801x801 true map is not needed) and stores
next to it the complex128 oracle's first NCG iteration (loss, source estimates, gradient on a 4x-decimated grid, step,
updated sound speed).  If the survey-time independent restatement's outputs are still on disk (/tmp/survey, not part of
the repo) the two are compared and the differences printed: two separately written restatements of the reference agree.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from common import cfg1_inputs, rel  # noqa: E402
from oracle import fwi as ofwi  # noqa: E402
from waveforminversionust_b200.matfile import load_mat73  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
MAT = "/root/reference/Final_python/RecordedData.mat"

if __name__ == "__main__":
    d = load_mat73(MAT)
    f = float(np.float32(d["f"].ravel()[0]))  # fwi_script.py:24 casts to float32
    rec = d["REC_DATA"].astype(np.complex64)  # :26
    xc, yc = d["x_circ"].astype(np.float32).ravel(), d["y_circ"].astype(np.float32).ravel()
    geom, rec_tx = cfg1_inputs(rec, xc, yc)
    hist = []
    t0 = time.time()
    VEL, sd, grad, ADJ_WV, WV = ofwi.nonlinear_conjugate_gradient_vectorized(
        geom.xi, geom.yi, geom.num_elements, rec_tx, geom.dense_src(np.complex128), geom.tx_include, geom.ind_matlab, 1480.0, f, 1,
        geom.a0, geom.L_PML, geom.mask_indices, dtype="c128", history=hist)
    print("oracle c128, 1 NCG iteration: %.0f s" % (time.time() - t0), hist[0])
    # source estimates of iteration 0 (WV is scaled by them: recover from the one-hot source node is not possible, recompute)
    loss0, grad0, fl = ofwi.fwi_loss_and_grad(1.0 / np.full((geom.Ny, geom.Nx), 1480.0), geom.xi, geom.yi, rec_tx, geom.dense_src(np.complex128),
                                              f, geom.a0, geom.L_PML, geom.tx_include, geom.ind_matlab, geom.mask_indices,
                                              geom.num_elements, dtype="c128", return_fields=True)
    print("loss0", loss0, "|grad0|", np.linalg.norm(grad0), "rel(grad NCG vs loss_and_grad)", rel(grad, grad0))
    np.savez_compressed(os.path.join(HERE, "cfg1_shipped.npz"), rec=rec, x_circ=xc, y_circ=yc, f=f,
                        loss0=hist[0]["loss"], grad_norm0=hist[0]["grad_norm"], step0=hist[0]["step"],
                        vel_min1=hist[0]["vel_min"], vel_max1=hist[0]["vel_max"], src_est0=fl["SRC_EST"],
                        grad0_dec4=grad0[::4, ::4], vel1_dec4=VEL[::4, ::4], grad0_sum=grad0.sum(), grad0_abs_sum=np.abs(grad0).sum())
    sv = "/tmp/survey/ncg_c128_1_python.npz"
    if os.path.exists(sv):
        s = np.load(sv)
        print("vs survey-time independent restatement (complex128, python stencil): grad0 rel %.3e, VEL after it.1 rel %.3e (RMS %.3e m/s), "
              "src_est rel %.3e, loss rel %.3e" % (rel(grad0, s["grad0"]), rel(VEL, s["VEL"]), float(np.sqrt(np.mean((VEL - s["VEL"]) ** 2))),
                                                     rel(fl["SRC_EST"], s["est0"]), abs(loss0 - float(s["losses"][0])) / loss0))
