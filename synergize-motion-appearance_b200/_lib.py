"""ctypes binding of the C ABI in include/sma_b200.h (the reference-side stub a maintainer would add;
see INTEGRATION.md).  Fails loudly when the CUDA library is missing: there is no CPU fallback."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('SMA_B200_LIB') or os.path.join(HERE, 'csrc', 'libsma_b200.so')      # (the override is for kernel A/B experiments)

c_f32p = C.c_void_p
c_i64 = C.c_int64


class ConvDesc(C.Structure):
    """Mirror of `struct sma_conv_desc`."""
    _fields_ = [
        ('x', C.c_void_p), ('B', C.c_int), ('Hi', C.c_int), ('Wi', C.c_int), ('Cin', C.c_int),
        ('in_bstride', c_i64), ('in_ld', C.c_int),
        ('w', C.c_void_p), ('ldw', C.c_int), ('bias', C.c_void_p),
        ('Cout', C.c_int), ('kh', C.c_int), ('kw', C.c_int), ('stride', C.c_int), ('pad_t', C.c_int), ('pad_l', C.c_int),
        ('upsample2', C.c_int),
        ('pre_scale', C.c_void_p), ('pre_shift', C.c_void_p), ('pre_act', C.c_int),
        ('y', C.c_void_p), ('Ho', C.c_int), ('Wo', C.c_int), ('out_bstride', c_i64), ('out_ld', C.c_int),
        ('act', C.c_int),
        ('res', C.c_void_p), ('res_bstride', c_i64), ('res_ld', C.c_int),
        ('d2s', C.c_int), ('out_nchw', C.c_int), ('precision', C.c_int), ('w_tc', C.c_void_p), ('w_tc16', C.c_void_p), ('w_ts', C.c_void_p), ('tc_variant', C.c_int), ('kernel_used', C.c_int),
        ('w_tc_nt', C.c_int), ('plan_only', C.c_int), ('aux', C.c_void_p), ('aux_bstride', c_i64), ('aux_ld', C.c_int), ('sft_w', C.c_float),
        ('gn_want', C.c_int), ('gn_partial', C.c_void_p), ('gn_chunks', C.c_int),
        ('x2', C.c_void_p), ('in2_bstride', c_i64), ('in2_ld', C.c_int), ('Cin1', C.c_int), ('split_ws', C.c_void_p), ('split_qscale', C.c_float), ('x2_k1', C.c_int),
    ]


_V, _I, _F, _L = C.c_void_p, C.c_int, C.c_float, C.c_int64

SIGNATURES = {
    'sma_abi_version': ([], C.c_int),
    'sma_status_string': ([_I], C.c_char_p),
    'sma_device_check': ([_I], C.c_int),
    'sma_kernel_launch_count': ([], C.c_int),
    'sma_conv2d_fwd': ([C.POINTER(ConvDesc), _V], C.c_int),
    'sma_sizeof_conv_desc': ([], C.c_int),
    'sma_pack_conv_weight': ([_V, _V, _I, _I, _I, _I, _V, _V, _V, _V, _F, _V, _I, _V, _V], C.c_int),
    'sma_conv_weight_tc_floats': ([_I, _I, _I, _I, _I], C.c_int64),
    'sma_pack_conv_weight_tc': ([_V, _I, _I, _I, _I, _I, _I, _V, _V], C.c_int),
    'sma_conv_weight_tc16_floats': ([_I, _I, _I, _I], C.c_int64),
    'sma_pack_conv_weight_tc16': ([_V, _I, _I, _I, _I, _I, _V, _V], C.c_int),
    'sma_conv_weight_ts_floats': ([_I, _I, _I, _I], C.c_int64),
    'sma_pack_conv_weight_ts': ([_V, _I, _I, _I, _I, _I, _V, _V], C.c_int),
    'sma_debug_conv_ts_prof': ([_V], C.c_int),
    'sma_groupnorm_stats': ([_V, _I, _I, _I, _L, _I, _I, _F, _V, _V, _V, _V, _V, _V], C.c_int),
    'sma_groupnorm_finalize_pairs': ([_V, _I, _I, _I, _I, _I, _I, _F, _V, _V, _V, _V, _I, _V], C.c_int),
    'sma_affine_act': ([_V, _I, _I, _I, _L, _I, _V, _V, _I, _V, _L, _I, _V], C.c_int),
    'sma_layernorm': ([_V, _I, _I, _V, _V, _F, _V, _I, _V, _V, _V], C.c_int),
    'sma_warp_occlude_fwd': ([_V, _L, _I, _I, _I, _I, _V, _V, _I, _I, _V, _V], C.c_int),
    'sma_warp_occlude_gather_fwd': ([_V, _L, _I, _I, _I, _I, _V, _V, _I, _I, _I, _I, _V, _V], C.c_int),
    'sma_resize_bilinear_ac': ([_V, _I, _I, _I, _I, _L, _I, _V, _I, _I, _L, _I, _V], C.c_int),
    'sma_gather_bilinear4': ([_V, _I, _I, _I, _I, _L, _I, _V, _I, _I, _V], C.c_int),
    'sma_blend_bilinear4': ([_V, _I, _I, _I, _I, _V, _I, _I, _L, _I, _V], C.c_int),
    'sma_mha_fwd': ([_V, _I, _V, _I, _V, _I, _L, _I, _I, _I, _I, _I, _F, _V, _V, _I, _I, _V], C.c_int),
    'sma_attn_split_kv': ([_V, _I, _V, _I, _I, _V, _V], C.c_int),
    'sma_attn256_workspace_bytes': ([_I, _I, _I], C.c_int64),
    'sma_attn256_fwd': ([_V, _I, _V, _I, _V, _I, _L, _L, _I, _I, _I, _F, _V, _V, _I, _I, _V], C.c_int),
    'sma_mha_e256_workspace_bytes': ([_I, _I, _I, _I], C.c_int64),
    'sma_mha_e256_fwd': ([_V, _I, _V, _I, _V, _I, _L, _L, _I, _I, _I, _F, _V, _V, _V, _I, _I, _V], C.c_int),
    'sma_vq_lookup_fwd': ([_V, _I, _I, _V, _I, _V, _V, _V, _V, _V], C.c_int),
    'sma_vq_workspace_floats': ([], C.c_int),
    'sma_vq_commit_fwd': ([_V, _V, _L, _F, _V, _V, _V, _V], C.c_int),
    'sma_antialias_down4': ([_V, _I, _I, _I, _I, _V, _V, _I, _V], C.c_int),
    'sma_avgpool2': ([_V, _I, _I, _I, _I, _V, _I, _V], C.c_int),
    'sma_kp_head_fwd': ([_V, _I, _I, _I, _I, _I, _F, _V, _V, _V], C.c_int),
    'sma_normalize_kp': ([_V, _V, _V, _V, _V, _V, _I, _I, _F, _V, _I, _V, _V, _V], C.c_int),
    'sma_hull_scale': ([_V, _V, _I, _V, _V], C.c_int),
    'sma_u8hwc_to_f32nchw': ([_V, _I, _I, _I, _I, _I, _V, _V], C.c_int),
    'sma_dense_motion_prep': ([_V, _I, _I, _V, _V, _V, _V, _I, _I, _F, _V, _I, _V, _V], C.c_int),
    'sma_dense_motion_head': ([_V, _I, _I, _I, _V, _V, _V, _V, _I, _I, _V, _V, _V, _V], C.c_int),
    'sma_im2col_small': ([_V, _I, _I, _I, _I, _I, _I, _I, _V, _I, _V], C.c_int),
    'sma_flow_to_px': ([_V, _I, _I, _I, _V, _I, _V], C.c_int),
    'sma_conv_tapsum': ([_V, _I, _I, _I, _I, _I, _I, _I, _V, _V, _I, _V], C.c_int),
    'sma_flow_update': ([_V, _V, _V, _I, _I, _I, _I, _V, _V, _V], C.c_int),
    'sma_motion_ignore_mask': ([_V, _I, _I, _I, _I, _I, _V, _V], C.c_int),
    'sma_sft_combine': ([_V, _V, _V, _F, _L, _V, _V], C.c_int),
    'sma_to_uint8': ([_V, _I, _I, _I, _I, _I, _I, _V, _V], C.c_int),
    'sma_nchw_to_nhwc': ([_V, _I, _I, _I, _I, _V, _I, _V], C.c_int),
    'sma_nhwc_to_nchw': ([_V, _I, _I, _I, _I, _I, _V, _V], C.c_int),
}

_lib = None


class SmaError(RuntimeError):
    pass


def load():
    """Load libsma_b200.so (built in-tree by build.py / __graft_entry__.build()).  Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SmaError(f'{LIB_PATH} is missing: run `python __graft_entry__.py` (or build.py) to compile the '
                       f'sm_100a kernels; there is no CPU fallback')
    lib = C.CDLL(LIB_PATH)
    for name, (args, res) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.argtypes = args
        fn.restype = res
    if lib.sma_sizeof_conv_desc() != C.sizeof(ConvDesc):
        raise SmaError(f'struct sma_conv_desc is {lib.sma_sizeof_conv_desc()} bytes in {LIB_PATH} but {C.sizeof(ConvDesc)} in _lib.ConvDesc: '
                       f'rebuild the library (python __graft_entry__.py)')
    _lib = lib
    return lib


def check(status: int, what: str = ''):
    if status != 0:
        msg = load().sma_status_string(status).decode()
        raise SmaError(f'{what}: {msg} (status {status})')
