"""mcx_b200: B200-native lattice sampling hot path of MonteCarloX.jl / SpinSystems.

Python mirror of the reference's Julia API for this path, over the C ABI of libmcx_b200.so
(include/mcx_b200.h).  Names follow the reference; Julia's `f!` is spelled `f_`."""
from . import _lib
from ._lib import McxError, build, lib
from .algorithms import (Glauber, HeatBath, ImportanceSampling, Metropolis, Multicanonical, WangLandau,
                         acceptance_rate, logistic, reset_)
from .binned_object import BinnedObject, get_centers, get_values
from .ensembles import (BoltzmannEnsemble, FunctionEnsemble, MulticanonicalEnsemble, WangLandauEnsemble,
                        logweight)
from .flat import (DeviceFlat, PairBoltzmannSpin2Ensemble, ParallelMulticanonical, distribute_logweight_,
                   flat_for, merge_histograms_)
from .ising2d_exact import distribution_exact_ising2D, distribution_from_logdos, logdos_exact_ising2D
from .checkpointing import CheckpointSession, checkpoint_, finalize_, init_checkpoint, restore_checkpoint
from .measurements import (integrated_autocorrelation_time, integrated_autocorrelation_times,
                           optimize_exchange_interval_, series_tau_int_, sweep_series_, tau_int)
from .parallel import (GPUBackend, ParallelChains, ParallelTempering, ReplicaExchange, SlabIsing, ThreadsBackend,
                       attempt_exchange_pair_, exchange_log_ratio, partition_slots, philox_family, set_betas,
                       update_)
from .rng import PhiloxRNG, exchange_u, philox4x32_10
from .spin_systems import (BlumeCapel, Context, Ising, IsingLatticeOptim, default_context, energy, init_,
                           magnetization, sweep_)
from .graph_systems import IsingGraph, IsingMatrix, grid_graph
from .tables import build_table, table_len
from .windows import DeviceWindow, WangLandauWindows, join_logdos, partition_windows

__all__ = [n for n in dir() if not n.startswith("_")]
