"""ctypes binding of libppgs_b200.so (include/ppgs_b200.h).

The CUDA library IS the product: there is no Python / PyTorch fallback.  If the
shared object is missing or cannot be loaded, importing this module raises."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIBRARY_PATH = os.path.join(_HERE, 'lib', 'libppgs_b200.so')

PRECISION_FP32 = 0
PRECISION_F16X2 = 1
PRECISION_F16 = 2
PRECISIONS = {'fp32': PRECISION_FP32, 'f16x2': PRECISION_F16X2, 'f16': PRECISION_F16}

E_INVALID, E_CUDA, E_STATE, E_TOO_LARGE, E_UNSUPPORTED = -1, -2, -3, -4, -5


class ModelConfig(ctypes.Structure):
    _fields_ = [
        ('input_channels', ctypes.c_int32),
        ('hidden_channels', ctypes.c_int32),
        ('num_layers', ctypes.c_int32),
        ('num_heads', ctypes.c_int32),
        ('ffn_channels', ctypes.c_int32),
        ('output_channels', ctypes.c_int32),
        ('kernel_size', ctypes.c_int32),
        ('is_causal', ctypes.c_int32),
        ('chunk_length', ctypes.c_int32),
        ('chunk_overlap', ctypes.c_int32),
        ('max_len', ctypes.c_int32),
        ('layer_norm_eps', ctypes.c_float),
    ]


# name -> (restype, argtypes); every symbol declared in include/ppgs_b200.h
_c = ctypes
_vp, _i, _i64, _sz = _c.c_void_p, _c.c_int, _c.c_int64, _c.c_size_t
SYMBOLS = {
    'ppgs_abi_version': (_i, []),
    'ppgs_last_error': (_c.c_char_p, []),
    'ppgs_default_config': (None, [_c.POINTER(ModelConfig)]),
    'ppgs_engine_create': (_i, [_c.POINTER(ModelConfig), _i, _c.POINTER(_vp)]),
    'ppgs_engine_destroy': (None, [_vp]),
    'ppgs_engine_set_weight': (_i, [_vp, _c.c_char_p, _vp, _c.POINTER(_i64), _i]),
    'ppgs_engine_finalize': (_i, [_vp]),
    'ppgs_engine_set_precision': (_i, [_vp, _i]),
    'ppgs_engine_get_precision': (_i, [_vp]),
    'ppgs_engine_blob_bytes': (_sz, [_vp]),
    'ppgs_engine_blob_dev': (_vp, [_vp]),
    'ppgs_engine_adopt_blob': (_i, [_vp]),
    'ppgs_mel_forward': (_i, [_vp, _vp, _i, _i64, _i64, _vp, _vp]),
    'ppgs_w2v2_finalize': (_i, [_vp]),
    'ppgs_w2v2fb_forward': (_i, [_vp, _vp, _i, _i64, _i64, _c.POINTER(_i64), _vp, _vp]),
    'ppgs_transformer_forward': (_i, [_vp, _vp, _i, _i, _c.POINTER(_i64), _i, _i, _vp, _vp]),
    'ppgs_from_audio': (_i, [_vp, _vp, _i, _i64, _i64, _c.POINTER(_i64), _i, _i, _vp, _vp]),
    'ppgs_from_audio_host': (_i, [_vp, _vp, _i, _i64, _c.POINTER(_i64), _i, _i, _vp, _vp]),
    'ppgs_from_audio_host_submit': (_i, [_vp, _vp, _i, _i64, _c.POINTER(_i64), _i, _i, _vp, _vp]),
    'ppgs_engine_wait': (_i, [_vp]),
    'ppgs_engine_launch_count': (_i64, [_vp]),
    'ppgs_engine_graph_replays': (_i64, [_vp]),
    'ppgs_engine_set_graphs': (_i, [_vp, _i]),
    'ppgs_engine_workspace_bytes': (_sz, [_vp]),
    'ppgs_engine_set_profiling': (_i, [_vp, _i]),
    'ppgs_engine_kernel_stat': (_i, [_vp, _i, _c.c_char_p, _sz, _c.POINTER(_c.c_double),
                                     _c.POINTER(_i64)]),
    'ppgs_wav_info': (_i, [_c.c_char_p, _c.POINTER(_i64), _c.POINTER(_i), _c.POINTER(_i),
                           _c.POINTER(_i), _c.POINTER(_i)]),
    'ppgs_wav_info_many': (_i, [_c.POINTER(_c.c_char_p), _i64, _i, _c.POINTER(_i64),
                                _c.POINTER(_c.c_int32), _c.POINTER(_c.c_int32), _c.POINTER(_c.c_int32),
                                _c.POINTER(_c.c_int32), _c.POINTER(_c.c_int32)]),
    'ppgs_wav_read_f32': (_i, [_c.c_char_p, _vp, _i64, _c.POINTER(_i64), _c.POINTER(_i)]),
    'ppgs_flac_info': (_i, [_c.c_char_p, _c.POINTER(_i64), _c.POINTER(_i), _c.POINTER(_i), _c.POINTER(_i)]),
    'ppgs_flac_read_f32': (_i, [_c.c_char_p, _vp, _i64, _c.POINTER(_i64), _c.POINTER(_i), _c.POINTER(_i)]),
    'ppgs_pcm16_to_f32': (_i, [_vp, _vp, _i64, _vp, _vp]),
    'ppgs_resample_length': (_i64, [_i64, _i, _i]),
    'ppgs_resample': (_i, [_vp, _vp, _i, _i64, _i64, _i, _i, _vp, _i64, _vp]),
    'ppgs_resample_taps': (_i, [_i, _i, _vp, _i64, _c.POINTER(_i), _c.POINTER(_i), _c.POINTER(_i)]),
    'ppgs_pt_write_f32': (_i, [_c.c_char_p, _vp, _i64, _i64, _i64]),
    'ppgs_pt_write_f16': (_i, [_c.c_char_p, _vp, _i64, _i64, _i64]),
    'ppgs_pt_info': (_i, [_c.c_char_p, _c.POINTER(_i), _c.POINTER(_i64), _c.POINTER(_i)]),
    'ppgs_pt_read': (_i, [_c.c_char_p, _vp, _i64, _i64, _i, _i64]),
    'ppgs_files_to_files': (_i, [_vp, _i, _c.POINTER(_c.c_int32), _c.POINTER(_c.c_char_p),
                                 _c.POINTER(_c.c_char_p), _c.POINTER(_i64), _i, _i, _i, _vp,
                                 _c.POINTER(_i64)]),
    'ppgs_stream_capacity': (_i, []),
    'ppgs_stream_create': (_i, [_vp, _i, _c.POINTER(_vp)]),
    'ppgs_stream_destroy': (None, [_vp]),
    'ppgs_stream_reset': (_i, [_vp, _vp]),
    'ppgs_stream_length': (_i, [_vp]),
    'ppgs_stream_emitted': (_i, [_vp]),
    'ppgs_stream_push': (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _c.POINTER(_i), _vp]),
    'ppgs_stream_push_ragged': (_i, [_vp, _vp, _i, _c.POINTER(_c.c_int32), _c.POINTER(_c.c_int32), _i,
                                     _vp, _i, _c.POINTER(_c.c_int32), _vp]),
    'ppgs_stream_reset_streams': (_i, [_vp, _c.POINTER(_c.c_int32), _vp]),
    'ppgs_stream_state': (_i, [_vp, _c.POINTER(_c.c_int32), _c.POINTER(_c.c_int32)]),
    'ppgs_ppg_distance': (_i, [_vp, _vp, _vp, _i, _i64, _i64, _i64, _vp, _c.c_float, _i, _vp, _vp]),
    'ppgs_ppg_interpolate': (_i, [_vp, _vp, _vp, _vp, _c.c_float, _i64, _i64, _vp, _vp]),
    'ppgs_ppg_grid_sample': (_i, [_vp, _vp, _i, _i64, _vp, _i64, _vp, _vp]),
    'ppgs_ppg_sparsify': (_i, [_vp, _vp, _i, _i, _i64, _i, _c.c_float, _vp, _vp]),
}


def _load():
    if not os.path.exists(LIBRARY_PATH):
        raise RuntimeError(
            f'ppgs_b200: CUDA library not built ({LIBRARY_PATH} is missing). '
            'Run `python -c "import __graft_entry__ as g; g.build()"` or '
            '`make -C ppgs_b200/csrc`. There is no CPU / PyTorch fallback.')
    lib = ctypes.CDLL(LIBRARY_PATH)
    for name, (restype, argtypes) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.ppgs_abi_version() != 1:
        raise RuntimeError('ppgs_b200: ABI version mismatch, rebuild the library')
    return lib


lib = _load()


def last_error():
    return lib.ppgs_last_error().decode('utf-8', 'replace')


def check(code):
    """Map a PPGS_E_* status to the exception type the reference raises."""
    if code == 0:
        return
    message = last_error()
    if code in (E_INVALID, E_TOO_LARGE):
        raise ValueError(message)
    raise RuntimeError(message)
