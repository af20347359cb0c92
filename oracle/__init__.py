"""CPU oracle for the WaveformInversionUST hot path.

TEST INFRASTRUCTURE ONLY.  A NumPy/SciPy restatement of the reference's
``solve_helmholtz.py`` / ``nonlinearcg.py`` / ``fwi_loss_function.py`` used as the
parity checker.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The
product package (``waveforminversionust_b200``) never does.

PARITY UNPINNED: the reference ships no tests, no golden vectors and no stored
outputs for this path, and JAX / jaxopt / mat73 are not installable in the build
image, so the oracle cannot be checked against outputs of the reference itself.
It is pinned only by (a) line-by-line restatement with file:line citations,
(b) the reference's own arithmetic back-end, SciPy ``spsolve`` -> SuperLU ``gssv``,
being called exactly as the reference calls it, (c) mathematical properties
(residual, adjoint dot-test, reciprocity) and (d) the shipped ``RecordedData.mat``
reconstruction reproducing the phantom (tests/golden/cfg1_known_answers.json).
"""
from .helmholtz import (  # noqa: F401
    stencil_opt_params,
    pml_profiles,
    assemble_helmholtz,
    assemble_planes,
    solve_helmholtz,
    HelmholtzFactor,
)
from .fwi import (  # noqa: F401
    estimate_src_strength_batched,
    receiver_gather,
    fwi_loss_function,
    fwi_loss_and_grad,
    nonlinear_conjugate_gradient_vectorized,
)
from . import timedomain  # noqa: F401
