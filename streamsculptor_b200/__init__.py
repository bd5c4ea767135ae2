"""streamsculptor_b200 - B200-native (sm_100a) implementation of the data-parallel hot path of
jnibauer/streamsculptor behind the reference's Python API (same class / function names as
`streamsculptor`, star-exported the way /root/reference/streamsculptor/__init__.py does).
"""
from .units import usys, dimensionless, UnitSystem, G_KPC_MYR_MSUN  # noqa: F401
from .solvers import Dopri5, Dopri8  # noqa: F401
from .main import Potential, Solution  # noqa: F401
from . import potential  # noqa: F401
from .potential import LinearTrack, CubicTrack  # noqa: F401
from .streamhelpers import (custom_release_model, gen_stream_ics_pert, gen_stream_vmapped_with_pert,  # noqa: F401
                            gen_stream_scan_with_pert, gen_stream_vmapped_with_pert_fixed_prog, gen_stream_ics_Chen25,
                            gen_stream_ics_pert_Chen25, gen_stream_vmapped_Chen25, gen_stream_vmapped_with_pert_Chen25,
                            gen_stream_vmapped_with_pert_Chen25_fixed_prog, eval_dense_stream, eval_dense_stream_id, get_Streakline_ICs,
                            gen_streakline, computed_binned_track, compute_stream_length, compute_length_oscillations, sample_from_1D_pdf)
from . import fields  # noqa: F401
from .fields import integrate_field  # noqa: F401
from . import perturbative  # noqa: F401
from . import RestrictedNbody  # noqa: F401
from . import GenerateImpactParams, generate_derivs  # noqa: F401
from .GenerateImpactParams import ImpactGenerator  # noqa: F401
from .RestrictedNbody import RestrictedNbody_generator, integrate_restricted_Nbody, initialize_prog_params  # noqa: F401

__all__ = ["usys", "Potential", "potential", "fields", "perturbative", "integrate_field", "Dopri5", "Dopri8", "LinearTrack", "CubicTrack"]
