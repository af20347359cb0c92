"""waveforminversionust_b200 -- B200-native (sm_100a) Helmholtz forward/adjoint solves and
adjoint-state gradient for ring-array ultrasound FWI, behind the reference's own Python surface.
See DESIGN.md and include/ustfwi.h."""
from .api import (  # noqa: F401
    OneHotSources,
    clear_plans,
    fwi_loss_function,
    nonlinear_conjugate_gradient,
    nonlinear_conjugate_gradient_vectorized,
    run_lbfgs_fwi,
    solve_helmholtz,
)
from .plan import HelmholtzPlan  # noqa: F401
