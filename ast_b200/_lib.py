"""ctypes binding of libast_b200.so (the C ABI declared in include/ast_b200.h).

There is no CPU fallback: if the CUDA library is missing or a call fails, this raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AST_B200_LIB") or os.path.join(_HERE, "libast_b200.so")     # override: A/B builds of the kernels (tools/)


class AstConfig(C.Structure):
    _fields_ = [("feat_dim", C.c_int),
                ("cnn_cout", C.c_int * 2), ("cnn_kh", C.c_int * 2), ("cnn_kw", C.c_int * 2),
                ("cnn_sh", C.c_int * 2), ("cnn_sw", C.c_int * 2), ("cnn_ph", C.c_int * 2), ("cnn_pw", C.c_int * 2),
                ("enc_layers", C.c_int), ("dec_layers", C.c_int),
                ("hidden_units", C.c_int), ("embedding_units", C.c_int), ("attn_units", C.c_int), ("vocab", C.c_int),
                ("drop_embed", C.c_float), ("drop_rnn", C.c_float), ("drop_out", C.c_float)]


class AstError(RuntimeError):
    pass


_P = C.c_void_p
_I = C.c_int
_F = C.c_float
_LL = C.c_longlong
_ULL = C.c_ulonglong

# name -> (restype, argtypes); every symbol declared in include/ast_b200.h
SIGNATURES = {
    "ast_last_error": (C.c_char_p, []),
    "ast_abi_version": (_I, []),
    "ast_launch_count": (_ULL, [_I]),
    "ast_create": (_I, [C.POINTER(AstConfig), _I, C.POINTER(_P)]),
    "ast_destroy": (_I, [_P]),
    "ast_param_floats": (_LL, [_P]),
    "ast_param_count": (_I, [_P]),
    "ast_param_info": (_I, [_P, _I, C.c_char_p, _I, C.POINTER(_LL), C.POINTER(_I), C.POINTER(_I)]),
    "ast_bn_state_floats": (_I, [_P]),
    "ast_bind_params": (_I, [_P, _P, _P, _P]),
    "ast_weights_changed": (_I, [_P]),
    "ast_workspace_bytes": (_LL, [_P, _I, _I, _I, _I, _I]),
    "ast_bind_workspace": (_I, [_P, _P, _LL, _I, _I, _I, _I, _I]),
    "ast_set_option": (_I, [_P, C.c_char_p, C.c_double]),
    "ast_get_option": (C.c_double, [_P, C.c_char_p]),
    "ast_encode": (_I, [_P, _P, _I, _I, _I, _P, _F, _P]),
    "ast_enc_len": (_I, [_P, _I]),
    "ast_get_enc_states": (_I, [_P, _P, _P]),
    "ast_forward_loss": (_I, [_P, _P, _P, _I, _I, _I, _P, _P, _F, _P, _P]),
    "ast_backward": (_I, [_P, _P]),
    "ast_grad_bucket_count": (_I, [_P]),
    "ast_grad_bucket_range": (_I, [_P, _I, C.POINTER(_LL), C.POINTER(_LL)]),
    "ast_grad_bucket_wait": (_I, [_P, _I, _P]),
    "ast_get_step_argmax": (_I, [_P, _P, _P]),
    "ast_opt_step": (_I, [_P, _P, _P, _P, _I, _F, _F, _F, _F, _F, _F, _F, C.POINTER(_I), _I, _P]),
    "ast_opt_step_sgd": (_I, [_P, _F, _F, _F, _F, C.POINTER(_I), _I, _P]),
    "ast_grad_buckets_mark": (_I, [_P, _P]),
    "ast_scale_grads": (_I, [_P, _F, _P]),
    "ast_last_grad_norm": (C.c_double, [_P, _P]),
    "ast_init_decoder_state": (_I, [_P, _I, _P]),
    "ast_get_encoder_states": (_I, [_P, _P, _P]),
    "ast_get_decoder_states": (_I, [_P, _P, _I, _P]),
    "ast_set_decoder_states": (_I, [_P, _P, _I, _P]),
    "ast_decode_step": (_I, [_P, _P, _P, _I, _P, _P, _P, _P]),
    "ast_attention": (_I, [_P, _P, _I, _P, _P, _P, _P, _P]),
    "ast_predict": (_I, [_P, _P, _I, _I, _I, _I, _I, _P, C.POINTER(_I), _P]),
    "ast_beam_search": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, C.POINTER(_I), C.POINTER(_I), _P, _P, _P, _P, _P, _P, _P]),
    "ast_beam_search_batch": (_I, [_P, _P, C.POINTER(_I), _I, _I, _I, _I, _I, _I, C.POINTER(_I), C.POINTER(_I), C.POINTER(_I), _P, _P, _P, _P,
                                   _I, _P, _P, _P]),
    "ast_pack_cmvn": (_I, [_P, _P, _P, _P, _P, _P, _P, _F, _ULL, _P, _I, _I, _I, _P]),
    "ast_softmax_ce": (_I, [_P, _I, _P, _I, _I, _P, _P, _P]),
    "ast_gemm": (_I, [_I, _I, _I, _I, _I, _I, _F, _P, _I, _P, _I, _F, _P, _I, _P, _P]),
    "ast_gemm_grouped": (_I, [_I, _I, _I, _I, _I, _I, _P, _LL, _I, _P, _LL, _I, _P, _LL, _I, _I, _P]),
    "ast_lstm_probe": (_I, [_P]),
    "ast_stage_times": (_I, [_P, _P, _P, _I, _I]),
    "ast_split_tf32": (_I, [_P, _P, _P, _LL, _P]),
    "ast_gemm3_nt": (_I, [_I, _I, _I, _P, _P, _P, _LL, _I, _P, _P, _P, _LL, _I, _P, _I, _P, _P]),
    "ast_lstm_seq": (_I, [_I, _P, _P, _P, _P, _P, _I, _I, _I, _P, _P, _I, _P]),
    "ast_debug_fetch": (_I, [_P, C.c_char_p, _P, _LL, C.POINTER(_LL), _P]),
}

_lib = None


def load():
    """Load the shared library (once).  Raises AstError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AstError(f"{LIB_PATH} not found: build it with `python -m ast_b200.build` "
                       "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().ast_last_error()
        raise AstError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")
    return rc


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())
