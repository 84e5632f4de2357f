"""sponet_b200 -- B200-native ensemble path for SPoNet's continuous-time spreading models.

Drop-in names for the hot path of lueckem/SPoNet (``CNVM``, ``CNTM``, their parameter classes, the
OpinionShares family of collective variables, ``sample_many_runs``, ``sample_moments``,
``estimate_steady_state``).  The event loops run as sm_100a CUDA kernels inside
``libsponet_b200.so`` (C ABI: include/sponet_b200.h); there is no CPU fallback.
"""
from .cntm import CNTM, CNTMParameters
from .cnvm import CNVM, CNVMParameters
from .collective_variables import (
    CollectiveVariable,
    CompositeCollectiveVariable,
    DegreeWeightedOpinionShares,
    Interfaces,
    OpinionShares,
    OpinionSharesByDegree,
    Propensities,
)
from .engine import set_default_mode
from .multiprocessing import sample_many_runs
from .network_generator import (
    BarabasiAlbertGenerator,
    ErdosRenyiGenerator,
    NetworkGenerator,
    RandomRegularGenerator,
    WattsStrogatzGenerator,
)
from .parameters import Parameters
from .sample_moments import sample_histogram, sample_moments, sample_moments_sweep
from .states import sample_states_target_shares, sample_states_uniform, sample_states_uniform_shares
from .steady_state import estimate_steady_state
from .stochastic_approximation import sample_stochastic_approximation

__all__ = [
    "CNTM", "CNTMParameters", "CNVM", "CNVMParameters", "CollectiveVariable",
    "CompositeCollectiveVariable", "DegreeWeightedOpinionShares", "Interfaces", "OpinionShares",
    "OpinionSharesByDegree", "Parameters", "Propensities", "sample_many_runs",
    "sample_moments", "sample_moments_sweep", "sample_histogram", "estimate_steady_state", "set_default_mode", "sample_states_uniform", "sample_states_uniform_shares",
    "sample_states_target_shares", "sample_stochastic_approximation", "NetworkGenerator", "ErdosRenyiGenerator",
    "RandomRegularGenerator", "BarabasiAlbertGenerator", "WattsStrogatzGenerator",
]
__version__ = "0.1.0"
