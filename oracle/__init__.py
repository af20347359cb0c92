"""CPU oracle for the WaveformInversionUST hot path.

TEST INFRASTRUCTURE ONLY.  A NumPy/SciPy restatement of the reference's
``solve_helmholtz.py`` / ``nonlinearcg.py`` / ``fwi_loss_function.py`` used as the
parity checker.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The
product package (``waveforminversionust_b200``) never does.

PINNED against the reference's own source text: the reference ships no tests or stored outputs and its runtime
(JAX) is not installable here, so ``tests/golden/make_ref_golden.py`` EXECUTES the unmodified
``Final_python/{solve_helmholtz,nonlinearcg,fwi_loss_function,fwi_script}.py`` on ``oracle/jax_shim.py`` (a
NumPy-backed stand-in for the ``jax`` names they use; SciPy ``spsolve`` -> SuperLU is the real thing, as in the
reference) and commits what they return (``tests/golden/ref_*.npz``).  ``tests/test_ref_pin.py`` holds this oracle to
those fixtures: assembled CSR matrix identical in pattern and to 1e-15 in value (x64 mode; 3e-7 in the reference's
single precision), stencil weights to all digits, wavefields to 1e-7, two NCG iterations to 5e-6 in the gradient and
1e-4 m/s in the sound speed, and ``fwi_script.main()`` on the shipped RecordedData.mat to the complex64 noise floor.
What stays third-party: SciPy's SuperLU (reference pins scipy==1.15.2, this image has 1.18.1) and XLA's own
float32 instruction selection, which the NumPy stand-in cannot reproduce bit for bit.
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
