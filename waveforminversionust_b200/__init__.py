"""waveforminversionust_b200 -- B200-native (sm_100a) Helmholtz forward/adjoint solves and
adjoint-state gradient for ring-array ultrasound FWI, behind the reference's own Python surface.
See DESIGN.md and include/ustfwi.h."""
from .api import (  # noqa: F401
    OneHotSources,
    channel_data,
    clear_plans,
    continuation_stages,
    frequency_continuation,
    fwi_loss_function,
    hanning,
    idtft,
    nonlinear_conjugate_gradient,
    nonlinear_conjugate_gradient_vectorized,
    run_lbfgs_fwi,
    solve_helmholtz,
    time_domain_simulation,
)
from .plan import HelmholtzPlan  # noqa: F401
