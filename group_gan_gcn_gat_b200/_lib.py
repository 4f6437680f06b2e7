"""ctypes binding of libsgx_b200.so (the C ABI declared in include/sgx.h).

There is deliberately NO fallback: if the shared library is missing or was not built for
sm_100a the import of any op raises.  Build it with ``python -c "import __graft_entry__ as g; g.build()"``
or ``make -C group_gan_gcn_gat_b200/csrc``.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SGX_LIB: developer override to A/B a differently-built copy of the same library (tools/ only)
LIB_PATH = os.environ.get('SGX_LIB') or os.path.join(_HERE, 'lib', 'libsgx_b200.so')

SGX_OK = 0
SGX_ERR_INVALID = -1
SGX_ERR_UNSUPPORTED = -2
SGX_ERR_CUDA = -3

PRECISION_FP32 = 0
PRECISION_BF16 = 1
PRECISION_TC32 = 2

_P = ctypes.c_void_p
_I64 = ctypes.c_int64
_I32 = ctypes.c_int32
_F32 = ctypes.c_float

# name -> (restype, argtypes); kept in the order of include/sgx.h
SIGNATURES = {
    'sgx_last_error': (ctypes.c_char_p, []),
    'sgx_version': (ctypes.c_int, []),
    'sgx_has_tcgen05': (ctypes.c_int, []),
    'sgx_set_option': (ctypes.c_int, [ctypes.c_char_p, _I32]),
    'sgx_launch_count': (ctypes.c_longlong, []),
    'sgx_profile_events': (ctypes.c_int, [_P, _P]),
    'sgx_schedule_stats': (ctypes.c_int, [_P, _I64, _P]),
    'sgx_schedule_fill': (ctypes.c_int, [_P, _I64, _P, _P, _P, _P, _P]),
    'sgx_schedule_build': (ctypes.c_int, [_P, _I64, _P, _P, _P, _P, _P, _P, _I32, _P, _P]),
    'sgx_schedule_device_ws_bytes': (_I64, [_I64]),
    'sgx_schedule_build_device': (ctypes.c_int, [_P, _I64, _I64, _I64, _I64, _P, _P, _P, _P, _P, _P, _P, _I64, _P]),
    'sgx_fetch_pinned': (ctypes.c_int, [_P, _P, _I64, _P]),
    'sgx_schedule_partition': (ctypes.c_int, [_P, _I64, _I32, _P, _P]),
    'sgx_schedule_chunks': (ctypes.c_int, [_P, _I64, _I32, _P, _P]),
    'sgx_group_ids': (ctypes.c_int, [_P, _P, _P, _P, _I64, _I64, _P, _P, _P, _P, _P]),
    'sgx_group_dense': (ctypes.c_int, [_P, _P, _P, _P, _I64, _I64, _P, _P, _P, _P, _P]),
    'sgx_pool_ws_bytes': (_I64, [_I64, _I32, _I32, _I32, _I32]),
    'sgx_pool_fwd': (ctypes.c_int, [_P, _P, _P, _P, _P, _P, _I64, _I64, _P, _P, _P, _P, _P, _P, _I32, _I32, _I32,
                                    _I32, _P, _P, _P, _I64, _P]),
    'sgx_pool_prep_bytes': (_I64, [_I32, _I32, _I32, _I32]),
    'sgx_pool_prep': (ctypes.c_int, [_P, _P, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _P, _I64, _P]),
    'sgx_pool_fwd_prepped': (ctypes.c_int, [_P, _P, _P, _P, _P, _P, _I64, _I64, _P, _P, _P, _P, _P, _P, _I32, _I32, _I32,
                                            _I32, _P, _P, _P, _P, _I64, _P]),
    'sgx_pool_tc32_available': (ctypes.c_int, [_I32, _I32, _I32]),
    'sgx_pool_bwd_ws_bytes': (_I64, [_I64, _I32, _I32, _I32]),
    'sgx_pool_bwd': (ctypes.c_int, [_P, _P, _P, _P, _P, _I64, _P, _P, _P, _P, _P, _P, _I32, _I32, _I32, _P, _P, _P,
                                    _P, _P, _P, _P, _P, _P, _I64, _P]),
    'sgx_pool_bwd_scenes': (ctypes.c_int, [_P, _P, _P, _P, _P, _P, _P, _I64, _P, _P, _P, _P, _P, _P, _I32, _I32, _I32, _P, _P,
                                           _P, _P, _P, _P, _P, _P, _P, _I64, _P]),
    'sgx_mlp2_supported': (ctypes.c_int, [_I32, _I32, _I32]),
    'sgx_mlp2_fwd': (ctypes.c_int, [_P, _I32, _P, _I32, _I64, _P, _P, _P, _P, _I32, _I32, _P, _P]),
    'sgx_gcn_module_ws_bytes': (_I64, [_I64, _I64, _I32, _I32, _I32, _I32]),
    'sgx_gcn_module_fwd': (ctypes.c_int, [_P] * 7 + [_I64, _I64] + [_P] * 6 + [_I32] * 4 + [_P, _P, _I64, _P]),
    'sgx_gcn_module_fused_fwd': (ctypes.c_int, [_P] * 7 + [_I64] + [_P] * 6 + [_I32] * 4 + [_P, _P]),
    'sgx_gcn_module_fused_fwd_labels': (ctypes.c_int, [_P] * 6 + [_I64] + [_P] * 6 + [_I32] * 4 + [_P, _P, _P]),
    'sgx_gcn_module_tc_prep_bytes': (_I64, [_I32] * 4),
    'sgx_gcn_module_tc_prep': (ctypes.c_int, [_P] * 6 + [_I32] * 4 + [_P, _P]),
    'sgx_gcn_module_bwd': (ctypes.c_int, [_P] * 8 + [_I64, _I64] + [_P] * 6 + [_I32] * 4 + [_P] * 7 + [_P, _I64, _P]),
    'sgx_gcn_module_fused_bwd_ws_bytes': (_I64, []),
    'sgx_gcn_module_fused_bwd': (ctypes.c_int, [_P] * 8 + [_I64] + [_P] * 6 + [_I32] * 4 + [_P] * 7 + [_P, _I64, _P]),
    'sgx_gat_encoder_ws_bytes': (_I64, [_I64, _I64, _I32, _I32, _I32, _I32, _I32]),
    'sgx_gat_encoder_fwd': (ctypes.c_int, [_P] * 5 + [_I64, _I64] + [_P] * 10 + [_F32] + [_I32] * 5 +
                            [_P, _P, _I64, _P]),
    'sgx_gat_encoder_fused_fwd': (ctypes.c_int, [_P] * 7 + [_I64, _I32] + [_P] * 10 + [_F32] + [_I32] * 5 + [_P, _P]),
    'sgx_gat_encoder_fused_fwd_labels': (ctypes.c_int, [_P] * 6 + [_I64] + [_P] * 10 + [_F32] + [_I32] * 5 + [_P, _P, _P]),
    'sgx_gat_encoder_tc_prep_bytes': (_I64, []),
    'sgx_gat_encoder_tc_prep': (ctypes.c_int, [_P] * 10 + [_I32] * 5 + [_P, _P]),
    'sgx_gat_encoder_bwd': (ctypes.c_int, [_P] * 6 + [_I64, _I64] + [_P] * 10 + [_F32] + [_I32] * 5 + [_P] * 11 +
                            [_P, _I64, _P]),
    'sgx_gat_encoder_fwd_dense': (ctypes.c_int, [_P] * 6 + [_I64, _I64, _I32] + [_P] * 10 + [_F32] + [_I32] * 5 +
                                  [_P, _P, _I64, _P]),
    'sgx_gat_encoder_bwd_dense': (ctypes.c_int, [_P] * 7 + [_I64, _I64, _I32] + [_P] * 10 + [_F32] + [_I32] * 5 + [_P] * 11 +
                                  [_P, _I64, _P]),
    'sgx_gat_encoder_fused_bwd_ws_bytes': (_I64, []),
    'sgx_gat_encoder_fused_bwd': (ctypes.c_int, [_P] * 8 + [_I64] + [_P] * 10 + [_F32] + [_I32] * 5 + [_P] * 11 +
                                  [_P, _I64, _P]),
    'sgx_lstm_ws_bytes': (_I64, []),
    'sgx_lstm_encoder_fwd': (ctypes.c_int, [_P, _I32, _I64, _P, _P, _P, _P, _P, _P, _I32, _I32, _P, _P, _I64, _I32, _P]),
    'sgx_lstm_prep': (ctypes.c_int, [_P] * 6 + [_I32, _I32, _P, _I64, _P]),
    'sgx_lstm_tape_floats': (_I64, [_I32, _I64, _I32]),
    'sgx_lstm_bwd_ws_bytes': (_I64, [_I32, _I64, _I32]),
    'sgx_lstm_encoder_train_fwd': (ctypes.c_int, [_P, _I32, _I64] + [_P] * 6 + [_I32, _I32, _P, _P, _P]),
    'sgx_lstm_decoder_train_fwd': (ctypes.c_int, [_P, _P, _P, _I32, _I64] + [_P] * 8 + [_I32, _I32, _P, _P, _P, _P]),
    'sgx_lstm_bwd': (ctypes.c_int, [_I32, _P, _I32, _I64] + [_P] * 7 + [_I32, _I32] + [_P] * 9 + [_I64, _P]),
    'sgx_lstm_decoder_fwd': (ctypes.c_int, [_P, _P, _P, _P, _P, _I32, _I32, _I64] + [_P] * 8 + [_I32, _I32, _P, _P, _P, _P, _I64, _I32, _P]),
    'sgx_displacement_errors': (ctypes.c_int, [_P, _P, _P, _I32, _I64, _P, _P, _I32, _I32, _P]),
    'sgx_best_of_k': (ctypes.c_int, [_P, _P, _P, _I64, _I32, _P, _P]),
    'sgx_dense_att_fwd': (ctypes.c_int, [_P, _P, _I64, _F32, _P, _P]),
    'sgx_dense_att_bwd': (ctypes.c_int, [_P, _P, _P, _I64, _F32, _P, _P, _P]),
    'sgx_gemm': (ctypes.c_int, [_P, _I64, _I64, _P, _I64, _I64, _P, _I64, _I64, _I64, _I64, _I32, _I32, _P]),
}

_lib = None


class SgxError(RuntimeError):
    pass


def lib():
    """The loaded shared library; raises loudly when it is absent (no CPU / eager fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise SgxError(
            'libsgx_b200.so not found at %s -- the CUDA extension is mandatory (no fallback). '
            'Run `python -c "import __graft_entry__ as g; g.build()"` from the repo root.' % LIB_PATH)
    handle = ctypes.CDLL(LIB_PATH)
    missing = []
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(handle, name)
        except AttributeError:
            missing.append(name)
            continue
        fn.restype = res
        fn.argtypes = args
    if missing:
        raise SgxError('libsgx_b200.so lacks symbols declared in include/sgx.h: %s' % ', '.join(missing))
    _lib = handle
    # switches are resolved HERE, once, and handed to the library as options (no getenv on any call path)
    for env, opt in (('SGX_LSTM_TC', b'lstm_tc'), ('SGX_GRAPH_TC', b'graph_tc'), ('SGX_PDL', b'pdl'), ('SGX_GAT_MMA', b'gat_mma'), ('SGX_GCN_MMA', b'gcn_mma')):
        if env in os.environ:
            val = 0 if os.environ[env] == '0' else 1
            if handle.sgx_set_option(opt, val) == SGX_OK:                      # unknown in this build: ignored
                OPTIONS[opt.decode()] = val
    if 'SGX_SCHED_DEVICE' in os.environ:
        OPTIONS['sched_device'] = 0 if os.environ['SGX_SCHED_DEVICE'] == '0' else 1
    return _lib


HOST_OPTIONS = ('sched_device',)             # switches of the python host side only (the library has no use for them)
OPTIONS = {'lstm_tc': 1, 'graph_tc': 1, 'pdl': 1, 'sched_device': 1}      # host-side mirror of the library switches (defaults of the build)


def set_option(name, value):
    """sgx_set_option: 'lstm_tc', 'graph_tc', 'pdl', and the host-side 'sched_device' (schedule arrays derived on the GPU) (and 'gat_mma' / 'gcn_mma' in -DSGX_AB_VARIANTS builds)."""
    if name not in HOST_OPTIONS:
        check(lib().sgx_set_option(name.encode(), int(value)), 'sgx_set_option')
    OPTIONS[name] = int(value)


def option(name):
    lib()
    return OPTIONS.get(name, 1)


def last_error():
    return lib().sgx_last_error().decode('utf-8', 'replace')


def check(rc, what=''):
    """Maps the C status to the exceptions the reference's callers would see."""
    if rc == SGX_OK:
        return
    msg = '%s: %s' % (what, last_error()) if what else last_error()
    if rc == SGX_ERR_INVALID:
        raise ValueError(msg)
    if rc == SGX_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise SgxError(msg)
