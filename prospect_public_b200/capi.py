"""ctypes binding of libpb200.so (include/pb200.h).  This is the same stub INTEGRATION.md
shows for a PROSPECT maintainer: plain pointers and sizes, no torch types in any signature.

There is NO CPU fallback: if the shared object cannot be loaded, or a call fails, an exception
is raised.  Device memory is owned by the caller (torch tensors) and passed as raw addresses.
"""
import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('PB200_LIB') or os.path.join(_PKG, '_lib', 'libpb200.so')   # PB200_LIB: A/B builds

ACTIVE, CONVERGED, FAILED = 0, 1, 2
STEP_EXPONENTIAL, STEP_ADAPTIVE_FIXED, STEP_ADAPTIVE_SECANT = 0, 1, 2
RECORD_NONE, RECORD_ACCEPTED, RECORD_THINNED = 0, 1, 2
EXACT_BOX = 0x100
ABI_VERSION = 5

_vp = C.c_void_p


class Problem(C.Structure):
    _fields_ = [('d', C.c_int32), ('n', C.c_int32), ('npad', C.c_int32), ('n_points', C.c_int32),
                ('saa', _vp), ('sab', _vp), ('mu', _vp), ('lo', _vp), ('hi', _vp), ('f', _vp), ('c0', _vp),
                ('lt_t', _vp), ('lam', _vp), ('lt_norm', _vp), ('g_t', _vp), ('h', _vp), ('lt_hi', _vp), ('lt_lo', _vp)]


class Chains(C.Structure):
    _fields_ = [('n_chains', C.c_int32), ('x', _vp), ('loglkl', _vp), ('temperature', _vp), ('step_size', _vp),
                ('current_best', _vp), ('point', _vp), ('chain_id', _vp), ('steps_done', _vp), ('iteration', _vp),
                ('status', _vp), ('secant_stage', _vp), ('secant_prev_mult', _vp), ('secant_cur_mult', _vp),
                ('secant_prev_ar', _vp)]


class Schedule(C.Structure):
    _fields_ = [('steps_per_iteration', C.c_int32), ('step_mode', C.c_int32), ('temperature_change', C.c_double),
                ('step_size_change', C.c_double), ('adaptive_lo', C.c_double), ('adaptive_hi', C.c_double),
                ('adaptive_multiplier', C.c_double), ('chi2_tolerance', C.c_double)]


class Records(C.Structure):
    _fields_ = [('max_iterations', C.c_int32), ('iter_stride', C.c_int64), ('best_loglkl', _vp), ('last_loglkl', _vp), ('temperature', _vp),
                ('step_size', _vp), ('accepted', _vp), ('best_x', _vp), ('last_x', _vp)]


class Samples(C.Structure):
    _fields_ = [('mode', C.c_int32), ('thin', C.c_int32), ('capacity', C.c_int32), ('x', _vp), ('loglkl', _vp),
                ('mult', _vp), ('count', _vp)]


MAX_PEERS = 16


class Peers(C.Structure):
    _fields_ = [('n_peers', C.c_int32), ('rank', C.c_int32), ('n_alloc', C.c_int32), ('rows', C.c_int32),
                ('row_stride', C.c_int64), ('rank_stride', C.c_int64), ('buffer', _vp * MAX_PEERS),
                ('progress', _vp * MAX_PEERS)]


class Hosted(C.Structure):
    _fields_ = [('n', C.c_int32), ('npad', C.c_int32), ('n_points', C.c_int32), ('factor_t', _vp), ('lo', _vp),
                ('hi', _vp)]


# name -> (restype, argtypes); every name here must be declared in include/pb200.h
SIGNATURES = {
    'pb200_last_error': (C.c_char_p, []),
    'pb200_abi_version': (C.c_int, []),
    'pb200_device_info': (C.c_int, [C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    'pb200_loglkl_batch': (C.c_int, [C.POINTER(Problem), _vp, _vp, C.c_int32, _vp, _vp, _vp]),
    'pb200_anneal_run': (C.c_int, [C.POINTER(Problem), C.POINTER(Chains), C.POINTER(Schedule), C.POINTER(Records),
                                   C.c_int32, C.c_uint64, C.c_int32, _vp]),
    'pb200_anneal_run_signalled': (C.c_int, [C.POINTER(Problem), C.POINTER(Chains), C.POINTER(Schedule),
                                             C.POINTER(Records), C.c_int32, C.c_uint64, C.c_int32, _vp,
                                             C.POINTER(C.c_int32), _vp]),
    'pb200_stream_wait_geq': (C.c_int, [_vp, _vp, C.c_uint32]),
    'pb200_anneal_run_shared': (C.c_int, [C.POINTER(Problem), C.POINTER(Chains), C.POINTER(Schedule),
                                          C.POINTER(Records), C.c_int32, C.c_uint64, C.c_int32, C.POINTER(Peers), _vp,
                                          C.POINTER(C.c_int32), _vp]),
    'pb200_mh_run': (C.c_int, [C.POINTER(Problem), C.POINTER(Chains), C.c_int32, C.c_uint64, C.POINTER(Samples),
                               _vp, C.c_int32, _vp]),
    'pb200_choose_lanes': (C.c_int, [C.c_int32, C.c_int32]),
    'pb200_uses_producer_warps': (C.c_int, [C.c_int32, C.c_int32]),
    'pb200_exchange_arrivals': (C.c_int, [C.c_int32, C.c_int32, C.c_int32]),
    'pb200_launch_status': (C.c_int, [_vp]),
    'pb200_hosted_propose': (C.c_int, [C.POINTER(Hosted), C.POINTER(Chains), C.c_uint64, _vp, _vp, _vp]),
    'pb200_hosted_accept': (C.c_int, [C.POINTER(Hosted), C.POINTER(Chains), C.c_uint64, _vp, _vp, _vp, _vp, _vp, _vp,
                                      _vp, _vp]),
    'pb200_schedule_update': (C.c_int, [C.POINTER(Chains), C.POINTER(Schedule), _vp, _vp, _vp]),
    'pb200_profile_reduce': (C.c_int, [_vp, _vp, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                       _vp, _vp, _vp, _vp, _vp, _vp]),
    'pb200_weighted_moments': (C.c_int, [_vp, _vp, C.c_int64, C.c_int32, C.c_int32, _vp, _vp]),
    'pb200_chain_moments': (C.c_int, [_vp, _vp, _vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _vp, _vp, _vp, _vp]),
    'pb200_rng_variates': (C.c_int, [C.c_uint64, C.c_uint32, C.c_uint32, C.c_int32, C.c_int32, _vp, _vp, _vp]),
    'pb200_philox4x32_10': (None, [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
}


class Pb200Error(RuntimeError):
    pass


_lib = None


def lib():
    """Load libpb200.so once.  Raises if it was not built (python -m prospect_public_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise Pb200Error(f'{LIB_PATH} not found: the CUDA extension is mandatory (no CPU fallback). '
                             'Build it with `python -m prospect_public_b200.build`.')
        handle = C.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(handle, name)      # AttributeError if the symbol is missing
            fn.restype = restype
            fn.argtypes = argtypes
        if handle.pb200_abi_version() != ABI_VERSION:
            raise Pb200Error('libpb200.so ABI version mismatch; rebuild')
        _lib = handle
    return _lib


def check(status):
    if status != 0:
        raise Pb200Error(lib().pb200_last_error().decode())


def ptr(t):
    """Raw address of a torch tensor (or None -> NULL)."""
    return None if t is None else _vp(t.data_ptr())


def philox4x32_10(counter, key):
    """Host-side Philox block (known-answer tests; needs no GPU)."""
    ctr = (C.c_uint32 * 4)(*counter)
    k = (C.c_uint32 * 2)(*key)
    out = (C.c_uint32 * 4)()
    lib().pb200_philox4x32_10(ctr, k, out)
    return list(out)
