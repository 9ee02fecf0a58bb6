"""ctypes binding of libmcx_b200.so (include/mcx_b200.h).  No CPU fallback: if the shared library
is missing or no CUDA device is present, every compute call raises."""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MCX_B200_LIB") or os.path.join(_HERE, "lib", "libmcx_b200.so")
CSRC = os.path.join(_HERE, "csrc")

OK, ERR_ARGUMENT, ERR_BOUNDS, ERR_CUDA, ERR_STATE, ERR_UNSUPPORTED = range(6)
ISING, BLUME_CAPEL = 0, 1
METROPOLIS, GLAUBER, HEATBATH = 0, 1, 2
STORAGE_INT8, STORAGE_BIT = 0, 1
INIT_UP, INIT_DOWN, INIT_ZERO, INIT_RANDOM = 0, 1, 2, 3
FLAT_MUCA, FLAT_WANG_LANDAU = 0, 1
OBS_ENERGY, OBS_SPIN2_WITH_PAIR_BOLTZMANN = 0, 1


class McxError(RuntimeError):
    """Non-argument failure inside libmcx_b200 (CUDA error, unsupported configuration)."""


def build(verbose=False):
    """Compile libmcx_b200.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    cmd = ["make", "-C", CSRC, "-j8"]
    if not verbose:
        cmd.append("-s")
    subprocess.check_call(cmd)
    return LIB_PATH


_vp, _i32, _i64, _u32, _u64, _dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_uint64, C.c_double
_P = C.POINTER

# name -> (restype, argtypes); every symbol include/mcx_b200.h declares
SIGNATURES = {
    "mcx_abi_version": (_i32, []),
    "mcx_last_error": (C.c_char_p, []),
    "mcx_ctx_create": (_i32, [_i32, _vp, _P(_vp)]),
    "mcx_ctx_destroy": (_i32, [_vp]),
    "mcx_ctx_set_stream": (_i32, [_vp, _vp]),
    "mcx_ctx_sync": (_i32, [_vp]),
    "mcx_ctx_info": (_i32, [_vp, _P(_i32), _P(_i32), _P(_i32), _P(_u64)]),
    "mcx_ctx_launch_count": (_i32, [_vp, _P(_u64)]),
    "mcx_ctx_async_error": (_i32, [_vp, _P(_i32)]),
    "mcx_ctx_clear_error": (_i32, [_vp]),
    "mcx_lattice_create": (_i32, [_vp, _i32, _i32, _P(_i32), _i32, _i32, _P(_vp)]),
    "mcx_lattice_destroy": (_i32, [_vp]),
    "mcx_lattice_set_couplings": (_i32, [_vp, _dbl, _dbl, _dbl]),
    "mcx_lattice_set_first_chain_id": (_i32, [_vp, _u32]),
    "mcx_lattice_upload": (_i32, [_vp, _vp]),
    "mcx_lattice_download": (_i32, [_vp, _vp]),
    "mcx_lattice_upload_begin": (_i32, [_vp, _vp]),
    "mcx_lattice_upload_commit": (_i32, [_vp]),
    "mcx_lattice_upload_bits": (_i32, [_vp, _vp]),
    "mcx_lattice_upload_bits_begin": (_i32, [_vp, _vp]),
    "mcx_lattice_download_bits": (_i32, [_vp, _vp]),
    "mcx_lattice_init": (_i32, [_vp, _i32, _u64]),
    "mcx_set_rule": (_i32, [_vp, _i32, _vp, _i32, _i32]),
    "mcx_set_labels": (_i32, [_vp, _vp]),
    "mcx_get_labels": (_i32, [_vp, _vp]),
    "mcx_set_rng": (_i32, [_vp, _u64, _u64]),
    "mcx_get_rng": (_i32, [_vp, _P(_u64), _P(_u64)]),
    "mcx_sweep": (_i32, [_vp, _i64]),
    "mcx_sweep_series": (_i32, [_vp, _i64, _i64, _vp]),
    "mcx_series_tau_int": (_i32, [_vp, _i64, _i32, _i64, _dbl, _vp]),
    "mcx_observables": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "mcx_energies": (_i32, [_vp, _vp]),
    "mcx_reset_counters": (_i32, [_vp]),
    "mcx_set_counters": (_i32, [_vp, _vp, _i64]),
    "mcx_recompute": (_i32, [_vp]),
    "mcx_set_tracking": (_i32, [_vp, _i32]),
    "mcx_pt_run": (_i32, [_vp, _i64, _i64]),
    "mcx_pt_run_info": (_i32, [_vp, _P(_i32), _P(_i32)]),
    "mcx_pt_export": (_i32, [_vp, _vp]),
    "mcx_pt_attach_peers": (_i32, [_vp, _i32, _i32, _vp]),
    "mcx_pt_peer_status": (_i32, [_vp, _P(_i32)]),
    "mcx_slab_configure": (_i32, [_vp, _i32, _i32]),
    "mcx_slab_export": (_i32, [_vp, _vp]),
    "mcx_slab_attach_ipc": (_i32, [_vp, _vp, _vp]),
    "mcx_slab_attach_local": (_i32, [_vp, _vp, _vp]),
    "mcx_slab_half_sweep": (_i32, [_vp]),
    "mcx_slab_status": (_i32, [_vp, _vp, _vp]),
    "mcx_lattice_device_sums": (_i32, [_vp, _P(_vp)]),
    "mcx_pt_create": (_i32, [_vp, _i32, _i32, _vp, _P(_vp)]),
    "mcx_pt_destroy": (_i32, [_vp]),
    "mcx_pt_energy_buffer": (_i32, [_vp, _P(_vp)]),
    "mcx_pt_publish": (_i32, [_vp]),
    "mcx_pt_exchange": (_i32, [_vp]),
    "mcx_pt_state": (_i32, [_vp, _vp, _vp, _vp, _P(_i64), _P(_i64)]),
    "mcx_pt_reset": (_i32, [_vp]),
    "mcx_pt_set_state": (_i32, [_vp, _vp, _vp, _vp, _i64, _i64]),
    "mcx_flat_create": (_i32, [_vp, _i32, _i32, _i64, _i64, _i64, _dbl, _i32, _P(_vp)]),
    "mcx_flat_destroy": (_i32, [_vp]),
    "mcx_flat_set_logweight": (_i32, [_vp, _vp]),
    "mcx_flat_get_logweight": (_i32, [_vp, _vp]),
    "mcx_flat_get_histogram": (_i32, [_vp, _vp]),
    "mcx_flat_reset_histogram": (_i32, [_vp]),
    "mcx_flat_set_logf": (_i32, [_vp, _dbl]),
    "mcx_flat_get_logf": (_i32, [_vp, _P(_dbl)]),
    "mcx_flat_sweep": (_i32, [_vp, _i64]),
    "mcx_flat_update": (_i32, [_vp]),
    "mcx_flat_device_histogram": (_i32, [_vp, _P(_vp), _P(_i64)]),
    "mcx_flat_device_logweight": (_i32, [_vp, _P(_vp), _P(_i64)]),
    "mcx_graph_create": (_i32, [_vp, _i64, _vp, _vp, _vp, _dbl, _i32, _dbl, _vp, _i32, _P(_vp)]),
    "mcx_graph_destroy": (_i32, [_vp]),
    "mcx_graph_colours": (_i32, [_vp, _P(_i32), _vp]),
    "mcx_graph_upload": (_i32, [_vp, _vp]),
    "mcx_graph_download": (_i32, [_vp, _vp]),
    "mcx_graph_init": (_i32, [_vp, _i32, _u64]),
    "mcx_graph_set_rule": (_i32, [_vp, _i32, _dbl]),
    "mcx_graph_set_rng": (_i32, [_vp, _u64, _u64, _u32]),
    "mcx_graph_get_rng": (_i32, [_vp, _P(_u64), _P(_u64)]),
    "mcx_graph_sweep": (_i32, [_vp, _i64]),
    "mcx_graph_observables": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "mcx_graph_energies": (_i32, [_vp, _vp]),
    "mcx_graph_reset_counters": (_i32, [_vp]),
}

_lib = None


def lib():
    """Load the shared library, binding every declared symbol.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise McxError("libmcx_b200.so not built (%s): run __graft_entry__.build(); there is no CPU fallback" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)   # AttributeError if the export is missing
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def check(status):
    """Map a status code to the exception the Julia shim would throw."""
    if status == OK:
        return
    msg = lib().mcx_last_error().decode("utf-8", "replace")
    if status == ERR_ARGUMENT:
        raise ValueError(msg)          # ArgumentError
    if status == ERR_BOUNDS:
        raise IndexError(msg)          # BoundsError
    if status == ERR_STATE:
        raise AssertionError(msg)      # AssertionError
    raise McxError("mcx status %d: %s" % (status, msg))
