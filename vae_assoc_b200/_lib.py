"""ctypes binding of libvaeassoc.so (include/vaeassoc.h).  No fallback: if the shared library is missing or
does not load, importing the model raises."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# VAEASSOC_LIB selects a build-time variant of the library (vae_assoc_b200/build.py: build_variant); default = the product
LIB_PATH = os.environ.get("VAEASSOC_LIB") or os.path.join(HERE, "libvaeassoc.so")
MAX_MODALITIES = 4
PEER_BLOB_BYTES = 128
ABI_VERSION = 2

RELU, SOFTPLUS = 0, 1
FP32, TF32, BF16 = 0, 1, 2
PARAMS, GRADS, ADAM_M, ADAM_V = 0, 1, 2, 3
(PROBE_Z_MEAN, PROBE_Z_LOG_SIGMA_SQ, PROBE_Z, PROBE_X_RECONSTR_MEAN, PROBE_RECONSTR_LOSS, PROBE_LATENT_LOSS,
 PROBE_VAE_COST, PROBE_ASSOC_COST, PROBE_D_Z_MEAN, PROBE_D_Z_LOG_SIGMA_SQ, PROBE_EPS) = range(11)


class Modality(C.Structure):
    _fields_ = [("n_input", C.c_int32), ("n_hidden_recog_1", C.c_int32), ("n_hidden_recog_2", C.c_int32),
                ("n_hidden_gener_1", C.c_int32), ("n_hidden_gener_2", C.c_int32), ("hidden_conv", C.c_int32),
                ("binary", C.c_int32), ("weight", C.c_float), ("scope", C.c_char * 32)]


class Config(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("n_modalities", C.c_int32), ("batch_size", C.c_int32),
                ("n_z", C.c_int32), ("transfer_fct", C.c_int32), ("precision", C.c_int32), ("device", C.c_int32),
                ("use_graph", C.c_int32), ("assoc_lambda", C.c_float), ("learning_rate", C.c_float),
                ("beta1", C.c_float), ("beta2", C.c_float), ("adam_epsilon", C.c_float),
                ("global_batch", C.c_int64), ("global_row0", C.c_int64), ("eps_seed", C.c_uint32),
                ("reserved", C.c_uint32), ("mod", Modality * MAX_MODALITIES)]


class TensorInfo(C.Structure):
    _fields_ = [("name", C.c_char * 48), ("role", C.c_char * 16), ("modality", C.c_int32), ("ndim", C.c_int32),
                ("shape", C.c_int32 * 4), ("offset", C.c_int64), ("rows", C.c_int64), ("cols", C.c_int64),
                ("ld", C.c_int64)]


Handle = C.c_void_p
FloatPP = C.POINTER(C.c_void_p)
I64P = C.POINTER(C.c_int64)

# name -> (restype, argtypes); every entry is declared in include/vaeassoc.h (tests/test_abi.py checks both ways)
SIGNATURES = {
    "vaeassoc_create": (C.c_int, [C.POINTER(Config), C.POINTER(Handle)]),
    "vaeassoc_destroy": (C.c_int, [Handle]),
    "vaeassoc_last_error": (C.c_char_p, [Handle]),
    "vaeassoc_abi_version": (C.c_int, []),
    "vaeassoc_set_stream": (C.c_int, [Handle, C.c_void_p]),
    "vaeassoc_stream_sync": (C.c_int, [Handle]),
    "vaeassoc_set_precision": (C.c_int, [Handle, C.c_int]),
    "vaeassoc_set_learning_rate": (C.c_int, [Handle, C.c_float]),
    "vaeassoc_num_tensors": (C.c_int, [Handle]),
    "vaeassoc_layout_query": (C.c_int, [Handle, C.c_int, C.POINTER(TensorInfo)]),
    "vaeassoc_flat_size": (C.c_int64, [Handle]),
    "vaeassoc_flat_ptr": (C.c_void_p, [Handle, C.c_int]),
    "vaeassoc_tensor_set": (C.c_int, [Handle, C.c_int, C.c_int, C.c_void_p]),
    "vaeassoc_tensor_get": (C.c_int, [Handle, C.c_int, C.c_int, C.c_void_p]),
    "vaeassoc_step_get": (C.c_int, [Handle, I64P]),
    "vaeassoc_step_set": (C.c_int, [Handle, C.c_int64]),
    "vaeassoc_train_step": (C.c_int, [Handle, FloatPP, I64P, C.c_void_p]),
    "vaeassoc_grad_step": (C.c_int, [Handle, FloatPP, I64P, C.c_void_p]),
    "vaeassoc_adam_step": (C.c_int, [Handle]),
    "vaeassoc_cost_read": (C.c_int, [Handle, C.POINTER(C.c_float)]),
    "vaeassoc_cost_history": (C.c_int, [Handle, C.c_int64, C.c_int64, C.c_void_p]),
    "vaeassoc_partial_fit_host": (C.c_int, [Handle, FloatPP, C.c_void_p, C.POINTER(C.c_float)]),
    "vaeassoc_submit_host": (C.c_int, [Handle, FloatPP, C.c_void_p]),
    "vaeassoc_submit_indexed": (C.c_int, [Handle, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "vaeassoc_submit_count": (C.c_int64, [Handle]),
    "vaeassoc_upload_wait": (C.c_int, [Handle, C.c_int64]),
    "vaeassoc_submit_costs": (C.c_int, [Handle, C.c_int64, C.c_int64, C.c_void_p]),
    "vaeassoc_eval_cost": (C.c_int, [Handle, FloatPP, I64P, C.c_void_p, C.POINTER(C.c_float)]),
    "vaeassoc_encode": (C.c_int, [Handle, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "vaeassoc_decode": (C.c_int, [Handle, C.c_int, C.c_void_p, C.c_void_p]),
    "vaeassoc_reconstruct": (C.c_int, [Handle, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "vaeassoc_infer_host": (C.c_int, [Handle, C.c_int, C.c_int, FloatPP, C.c_void_p, FloatPP]),
    "vaeassoc_probe_get": (C.c_int, [Handle, C.c_int, C.c_int, C.c_void_p, C.c_int64, I64P]),
    "vaeassoc_probe_mask": (C.c_int, [Handle, C.c_int, C.c_int, C.c_void_p, C.c_int64, I64P, I64P]),
    "vaeassoc_synth_batch": (C.c_int, [Handle, C.c_uint32, C.c_uint32, C.c_int64, C.c_int64, FloatPP]),
    "vaeassoc_philox_normal": (C.c_int, [Handle, C.c_uint32, C.c_uint32, C.c_int64, C.c_int64, C.c_int32,
                                         C.c_uint32, C.c_void_p]),
    "vaeassoc_comm_unique_id": (C.c_int, [C.c_char_p, C.c_void_p]),
    "vaeassoc_comm_init": (C.c_int, [Handle, C.c_char_p, C.c_void_p, C.c_int, C.c_int]),
    "vaeassoc_comm_destroy": (C.c_int, [Handle]),
    "vaeassoc_comm_sync_state": (C.c_int, [Handle]),
    "vaeassoc_comm_check": (C.c_int, [Handle]),
    "vaeassoc_peer_export": (C.c_int, [Handle, C.c_void_p]),
    "vaeassoc_peer_attach": (C.c_int, [Handle, C.c_void_p]),
    "vaeassoc_arena_floats": (C.c_int64, [Handle]),
    "vaeassoc_peer_attach_symmetric": (C.c_int, [Handle, C.c_void_p, FloatPP, C.c_void_p]),
    "vaeassoc_peer_multicast": (C.c_int, [Handle]),
    "vaeassoc_peer_detach": (C.c_int, [Handle]),
    "vaeassoc_peer_active": (C.c_int, [Handle]),
    "vaeassoc_debug_guard_check": (C.c_int, [Handle, I64P, I64P]),
    "vaeassoc_save": (C.c_int, [Handle, C.c_char_p]),
    "vaeassoc_load": (C.c_int, [Handle, C.c_char_p]),
    "vaeassoc_launch_count": (C.c_int64, [Handle]),
    "vaeassoc_debug_gemm": (C.c_int, [Handle, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64,
                                      C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_int64, C.c_int, C.c_int]),
    "vaeassoc_profile_step": (C.c_int, [Handle, FloatPP, I64P, C.c_void_p, C.c_char_p, C.POINTER(C.c_float),
                                        C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int]),
}

_lib = None


def load():
    """dlopen libvaeassoc.so and declare every prototype.  Raises if the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libvaeassoc.so is not built (%s). Run `python -m vae_assoc_b200.build`; there is no "
                          "CPU / PyTorch fallback for the associated-VAE train step." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header / library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    if lib.vaeassoc_abi_version() != ABI_VERSION:
        raise ImportError("libvaeassoc ABI %d != binding ABI %d" % (lib.vaeassoc_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


def nccl_library_path():
    """torch's bundled libnccl.so.2 (the library dlopens it; nothing is linked at build time)."""
    try:
        import nvidia.nccl as _n
        for root in list(getattr(_n, "__path__", [])):
            p = os.path.join(root, "lib", "libnccl.so.2")
            if os.path.exists(p):
                return p
    except ImportError:
        pass
    return "libnccl.so.2"
