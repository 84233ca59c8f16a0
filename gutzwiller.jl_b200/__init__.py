"""gutzwiller.jl_b200 -- B200-native kinetic-VQMC hot path of mtsch/Gutzwiller.jl.

Host mirror (Python, ctypes) of the reference's exported API for the path
``HubbardReal1D`` + ``GutzwillerAnsatz``: ``kinetic_vqmc``, ``KineticVQMC``, ``LocalEnergyEvaluator``,
``val_and_grad``, ``val_err_and_grad``, ``amsgrad``, ``gradient_descent`` (src/Gutzwiller.jl:16-43).
All arithmetic runs in hand-written sm_100a CUDA kernels behind the C ABI declared in
``include/gutzwiller_b200.h``; importing this package loads ``lib/libgutzwiller_b200.so`` and fails
loudly if it has not been built.  The directory name contains a dot, so load it through the
``gutzwiller_b200`` shim module at the repository root (``import gutzwiller_b200 as gw``).
"""
from . import _capi
from . import _capi as capi
from ._capi import (KERNEL_BITPARALLEL, KERNEL_ENUMERATE, GutzwillerError, lib)
from .amsgrad import (GradientDescentResult, adaptive_gradient_descent, amsgrad, gradient_descent, val_and_grad,
                      val_err_and_grad)
from .ansatz import (CombinationAnsatz, DensityProfileAnsatz, ExtendedGutzwillerAnsatz, ExtendedHubbardReal1D, GenericGutzwillerAnsatz, GenWalkers,
                     JastrowAnsatz,
                     MultinomialAnsatz, RelativeJastrowAnsatz, VectorAnsatz, collect_to_vec)
from .hamiltonian import (AbstractAnsatz, BoseFS, GutzwillerAnsatz, HubbardReal1D, keytype, near_uniform,
                          num_modes, num_parameters, starting_address, valtype)
from .localenergy import LocalEnergyEvaluator
from .runtime import (Context, Model, comm_init_all, default_context, distributed_state, init_distributed,
                      launch_count, nccl_unique_id, shard_range)
from .sparse import (SparseGutzwillerAnsatz, SparseHamiltonian, SparseVectorAnsatz, SparseWalkers, TableAnsatz)
from .vqmc import (BlockingResult, CombinedBlockingResult, KineticVQMC, KineticVQMCResult, Walkers,
                   blocking_analysis, default_walkers, kinetic_vqmc, kinetic_vqmc_continue,
                   local_energy_estimator, mean, resample, run_group, sampled_vector, series_blocking, var)

lib()  # fail at import time if the CUDA library is missing

__all__ = [n for n in dir() if not n.startswith("_")]
