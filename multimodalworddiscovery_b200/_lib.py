"""ctypes binding of ``libmwd_b200.so`` (C ABI declared in ``include/mwd_b200.h``).

There is no CPU fallback: if the shared library is missing, loading raises and every product code
path that needs a kernel fails with it.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# MWD_B200_LIB: an alternative build of the same library (A/B timing of a kernel change on one box)
LIB_PATH = os.environ.get('MWD_B200_LIB') or os.path.join(_HERE, 'libmwd_b200.so')

NMAX = 16
KMAX = 128
INIT_STRIDE = NMAX
TRANS_STRIDE = NMAX * NMAX
EPS = 1e-50


MIXED_CONCEPT, MIXED_POSTERIOR, MIXED_GRAD, MIXED_RECURSION = 1, 2, 4, 8


class MwdError(RuntimeError):
    pass


def mixed_bits(spec):
    """modelConfigs['posterior_precision'] / engine ``mixed_precision`` -> MWD_MIXED_* bits.
    'float64' | None | 0: reference arithmetic everywhere (default).
    'mixed' (= 'all'): everything that holds the north-star tolerance (1e-5 on log-likelihood and every table over
             20 EM iterations, tests/test_gpu_mixed_precision.py) off the FP64 pipe: the two floor-free GEMMs
             (softmaxLayer, updateSoftmaxWeight) on the tcgen05 tensor cores, the forward / backward lattice in scaled
             float32, and the updateConceptCounts chains in float32 with the clamped emission carried as a float32
             (hi, lo) pair (4.4e-6 on obs after 20 iterations; a single rounded float32 table gave 1.4e-5 and was
             kept out of 'mixed' until the pair landed).
    Or an explicit int / '+'-joined subset of 'concept', 'posterior', 'grad', 'recursion'."""
    if spec is None or spec is False or spec == 0 or spec == 'float64':
        return 0
    if spec is True or spec in ('mixed', 'all'):
        return MIXED_CONCEPT | MIXED_POSTERIOR | MIXED_GRAD | MIXED_RECURSION
    if isinstance(spec, int):
        return spec & 15
    bits = 0
    for part in str(spec).split('+'):
        bits |= {'concept': MIXED_CONCEPT, 'posterior': MIXED_POSTERIOR, 'grad': MIXED_GRAD,
                 'recursion': MIXED_RECURSION}[part.strip()]
    return bits


class Geometry(C.Structure):
    _fields_ = [('sm_count', C.c_int32), ('estep_grid', C.c_int32), ('grad_splits', C.c_int32)]


class IkProblem(C.Structure):
    _fields_ = [
        ('n_pairs', C.c_int64), ('n_regions', C.c_int64), ('n_phones_total', C.c_int64),
        ('feat_dim', C.c_int32), ('feat_is_f64', C.c_int32), ('n_concepts', C.c_int32),
        ('n_phone_types', C.c_int32), ('t_max', C.c_int32), ('n_buckets', C.c_int32),
        ('bucket_n', C.c_void_p), ('bucket_lo', C.c_void_p), ('bucket_tmax', C.c_void_p),
        ('region_off', C.c_void_p), ('phone_off', C.c_void_p), ('feats', C.c_void_p),
        ('phones', C.c_void_p),
        ('init', C.c_void_p), ('trans', C.c_void_p), ('obsT', C.c_void_p),
        ('pz', C.c_void_p), ('concept_counts', C.c_void_p),
        ('pair_ll', C.c_void_p), ('concept_counts_a', C.c_void_p),
        ('part_phone', C.c_void_p), ('part_init', C.c_void_p), ('part_trans', C.c_void_p),
        ('scratch', C.c_void_p), ('scratch_bytes', C.c_int64),
        ('stats', C.c_void_p), ('slot_off', C.c_void_p), ('no_floor', C.c_int32), ('mixed_precision', C.c_int32),
        ('concept_alignment', C.c_void_p),
    ]


class PartialSizes(C.Structure):
    _fields_ = [('phone_elems', C.c_int64), ('init_elems', C.c_int64), ('trans_elems', C.c_int64)]


class IkMstepArgs(C.Structure):
    _fields_ = [
        ('gaussian', C.c_int32), ('n_concepts', C.c_int32), ('n_phone_types', C.c_int32),
        ('feat_dim', C.c_int32), ('n_lens', C.c_int32), ('flags', C.c_int32), ('lens', C.c_void_p),
        ('toeplitz', C.c_int32), ('n_pairs_global', C.c_int64),
        ('lr', C.c_double), ('momentum', C.c_double), ('width', C.c_double),
        ('counts', C.c_void_p), ('grad', C.c_void_p), ('init', C.c_void_p), ('trans', C.c_void_p),
        ('obsT', C.c_void_p), ('posterior_param', C.c_void_p),
    ]


class HmmProblem(C.Structure):
    _fields_ = [
        ('n_pairs', C.c_int64), ('n_slots', C.c_int64),
        ('n_tgt_types', C.c_int32), ('n_src_types', C.c_int32), ('t_max', C.c_int32),
        ('log_domain', C.c_int32), ('n_buckets', C.c_int32), ('reserved', C.c_int32),
        ('bucket_n', C.c_void_p), ('bucket_lo', C.c_void_p), ('bucket_tmax', C.c_void_p),
        ('tgt_off', C.c_void_p), ('tgt', C.c_void_p), ('src_off', C.c_void_p), ('src', C.c_void_p),
        ('slot_off', C.c_void_p),
        ('init', C.c_void_p), ('trans', C.c_void_p), ('obs', C.c_void_p),
        ('pair_ll', C.c_void_p), ('post', C.c_void_p),
        ('part_init', C.c_void_p), ('part_trans', C.c_void_p),
        ('alpha_out', C.c_void_p), ('beta_out', C.c_void_p), ('emis', C.c_void_p),
        ('n_src_rows', C.c_int64), ('row_pair', C.c_void_p), ('slot_row', C.c_void_p),
    ]


class HmmMstepArgs(C.Structure):
    _fields_ = [
        ('log_domain', C.c_int32), ('n_tgt_types', C.c_int32), ('n_src_types', C.c_int32),
        ('n_lens', C.c_int32), ('lens', C.c_void_p), ('counts', C.c_void_p), ('acc', C.c_void_p),
        ('init', C.c_void_p), ('trans', C.c_void_p), ('obs', C.c_void_p),
    ]


# every symbol include/mwd_b200.h declares: name -> (restype, argtypes)
_vp, _i, _i64, _d = C.c_void_p, C.c_int, C.c_int64, C.c_double
SYMBOLS = {
    'mwd_last_error': (C.c_char_p, []),
    'mwd_version': (_i, []),
    'mwd_abi_sizeof': (_i, [_i]),
    'mwd_get_geometry': (_i, [C.POINTER(Geometry)]),
    'mwd_ik_scratch_bytes': (_i64, [C.POINTER(IkProblem)]),
    'mwd_posterior_linear': (_i, [_vp, _i, _i64, _i, _vp, _i, _vp, _vp]),
    'mwd_posterior_tc_scratch_bytes': (_i64, [_i, _i]),
    'mwd_posterior_tc_supported': (_i, [_i, _i, _i]),
    'mwd_posterior_linear_tc': (_i, [_vp, _i64, _i, _vp, _i, _vp, _vp, _i, _vp]),
    'mwd_posterior_grad_tc_supported': (_i, [_i, _i, _i]),
    'mwd_posterior_grad_tc_partials_len': (_i64, [_i, _i]),
    'mwd_ik_posterior_grad_tc_partial': (_i, [C.POINTER(IkProblem), _vp, _i, _i, _vp]),
    'mwd_posterior_grad_tc_finish': (_i, [_i, _i, _vp, _vp, _vp]),
    'mwd_umma_probe': (_i, [_vp, _i, _vp, _i, C.c_uint64, C.c_uint64, C.c_uint32, _i, _i, C.c_uint32, C.c_uint32, _vp, _vp]),
    'mwd_posterior_gaussian': (_i, [_vp, _i, _i64, _i, _vp, _d, _i, _vp, _vp, _vp]),
    'mwd_hidden_relu': (_i, [_vp, _i, _i64, _i, _vp, _i, _vp, _vp]),
    'mwd_backprop_hidden': (_i, [_vp, _vp, _vp, _vp, _i64, _i, _i, _vp, _vp]),
    'mwd_outer_grad_partials_len': (_i64, [_i, _i]),
    'mwd_outer_grad': (_i, [_vp, _i, _i64, _i, _vp, _vp, _i, _vp, _vp, _vp]),
    'mwd_sgd_update': (_i, [_vp, _vp, _i64, _d, _d, _d, _vp]),
    'mwd_dense_emission': (_i, [_vp, _vp, _i64, _i, _i, _vp, _vp, _vp]),
    'mwd_concept_phone_partials_len': (_i64, [_i, _i]),
    'mwd_concept_phone_counts': (_i, [_vp, _vp, _i64, _i, _i, _vp, _vp, _vp]),
    'mwd_ik_estep': (_i, [C.POINTER(IkProblem), _vp]),
    'mwd_ik_loglik': (_i, [C.POINTER(IkProblem), _vp]),
    'mwd_ik_partial_sizes': (_i, [_i, _i, C.POINTER(PartialSizes)]),
    'mwd_ik_concept_counts': (_i, [C.POINTER(IkProblem), _vp]),
    'mwd_ik_counts_len': (_i64, [_i, _i]),
    'mwd_ik_reduce_counts': (_i, [C.POINTER(IkProblem), _vp, _vp]),
    'mwd_ik_posterior_grad': (_i, [C.POINTER(IkProblem), _vp, _vp, _vp]),
    'mwd_ik_posterior_grad_partial': (_i, [C.POINTER(IkProblem), _vp, _i, _vp]),
    'mwd_ik_posterior_grad_finish': (_i, [_i, _i, _vp, _vp, _vp]),
    'mwd_ik_mstep': (_i, [C.POINTER(IkMstepArgs), _vp]),
    'mwd_ik_decode': (_i, [C.POINTER(IkProblem), _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    'mwd_write_alignment_files': (_i, [C.c_char_p, C.c_char_p, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i]),
    'mwd_format_float_repr': (_i, [_d, C.c_char_p, _i]),
    'mwd_fill_f64': (_i, [_vp, _i64, _d, _vp]),
    'mwd_sum_f64': (_i, [_vp, _i64, _vp, _vp, _vp]),
    'mwd_rank_reduce': (_i, [_vp, _i, _i64, _i, _i64, _vp, _vp]),
    'mwd_argmax_rows': (_i, [_vp, _i64, _i, _vp, _vp]),
    'mwd_ik_forward_dense': (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    'mwd_ik_backward_dense': (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    'mwd_hmm_warps': (_i, []),
    'mwd_hmm_estep': (_i, [C.POINTER(HmmProblem), _vp]),
    'mwd_hmm_counts_len': (_i64, [_i, _i]),
    'mwd_hmm_reduce': (_i, [C.POINTER(HmmProblem), _vp, _vp, _vp, _vp]),
    'mwd_hmm_mstep': (_i, [C.POINTER(HmmMstepArgs), _vp]),
    'mwd_hmm_align': (_i, [C.POINTER(HmmProblem), _d, _vp, _vp, _vp, _vp]),
    'mwd_hmm_gauss_emission': (_i, [C.POINTER(HmmProblem), _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'mwd_hmm_gauss_stats': (_i, [C.POINTER(HmmProblem), _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    'mwd_hmm_gauss_update': (_i, [_i, _i, _i, _vp, _i, _vp, _vp, _vp, _vp]),
}

_lib = None


def load():
    """Load the library once; raise MwdError (never fall back) if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MwdError('%s not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                       'or `make -C multimodalworddiscovery_b200/csrc` (there is no CPU fallback)'
                       % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    for which, st in enumerate((Geometry, IkProblem, PartialSizes, IkMstepArgs, HmmProblem, HmmMstepArgs)):
        if lib.mwd_abi_sizeof(which) != C.sizeof(st):
            raise MwdError('ABI mismatch: %s is %d bytes in libmwd_b200.so, %d in the binding'
                           % (st.__name__, lib.mwd_abi_sizeof(which), C.sizeof(st)))
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise MwdError('mwd_b200: %s' % load().mwd_last_error().decode('utf-8', 'replace'))


def geometry():
    g = Geometry()
    check(load().mwd_get_geometry(C.byref(g)))
    return g
