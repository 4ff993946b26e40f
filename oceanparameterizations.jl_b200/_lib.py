"""ctypes binding of libopz.so (include/opz.h).  No compute happens in Python: every call below
forwards plain pointers and sizes to the C ABI.  There is no fallback — if the shared library is
missing the import fails loudly, and if no sm_100 GPU is present `opz_ctx_create` fails."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libopz.so")


def _prefer_bundled_nccl() -> None:
    """libopz binds NCCL at run time (dlopen on first use of a communicator).  The dynamic loader keeps ONE library per SONAME
    in a process, so if libopz loaded the system libnccl.so.2 first, a later `import torch` would be bound to it as well —
    and fails when PyTorch's own wheel is newer.  Point libopz at the wheel PyTorch uses (nvidia-nccl), found without
    importing torch; an explicit OPZ_NCCL_LIB wins."""
    if os.environ.get("OPZ_NCCL_LIB"):
        return
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for loc in (spec.submodule_search_locations or []) if spec else []:
            cand = os.path.join(loc, "lib", "libnccl.so.2")
            if os.path.exists(cand):
                os.environ["OPZ_NCCL_LIB"] = cand
                return
    except Exception:
        pass


_prefer_bundled_nccl()

# ---- constants mirrored from include/opz.h -------------------------------------------------------
OPZ_OK, OPZ_EINVAL, OPZ_ENODEV, OPZ_ECUDA, OPZ_ENOMEM, OPZ_EUNSUP, OPZ_ENCCL = 0, -1, -2, -3, -4, -5, -6
COMM_ID_BYTES = 128
FLAG_MPP = 1 << 0
FLAG_CA = 1 << 1
FLAG_CA_SELECT_DUDZ = 1 << 2
FLAG_BC_MINUS_SCALE0 = 1 << 3
FLAG_DIURNAL = 1 << 4
FLAG_DIURNAL_MINUS_SCALE0 = 1 << 5
FLAG_RI_EPS = 1 << 6
FLAG_SMOOTH_NN = 1 << 7
FLAG_SMOOTH_RI = 1 << 8
SCHEME_EULER, SCHEME_RK2, SCHEME_RK4, SCHEME_IMEX = 0, 1, 2, 3
ACT_IDENTITY, ACT_RELU, ACT_TANH, ACT_MISH, ACT_SWISH, ACT_LEAKYRELU = range(6)
PREC_FP32, PREC_BF16, PREC_BF16X2, PREC_BF16X3, PREC_FP16, PREC_FP16X2, PREC_FP16X3 = 0, 1, 2, 3, 4, 5, 6
BC_STRIDE = 8

EXPORTED_SYMBOLS = (
    "opz_version", "opz_last_error", "opz_ctx_create", "opz_ctx_destroy", "opz_ctx_set_stream",
    "opz_ctx_synchronize", "opz_ctx_launch_count", "opz_model_create", "opz_model_set_weights",
    "opz_model_get_weights", "opz_model_set_weights_dev", "opz_model_n_params", "opz_model_destroy",
    "opz_rhs", "opz_rollout_forward", "opz_rollout_forward_dev", "opz_loss_grad", "opz_loss_grad_dev", "opz_loss_grad_mpp", "opz_loss_grad_mpp_dev", "opz_profile_diagnostics", "opz_flux_loss_grad",
    "opz_mlp_forward",
    "opz_comm_get_unique_id", "opz_ctx_comm_init", "opz_ctx_comm_adopt", "opz_ctx_comm_destroy", "opz_ctx_comm_info",
    "opz_comm_nccl_version", "opz_shard_range",
)


class OpzParams(C.Structure):
    """struct opz_params (include/opz.h): 37 x 4 bytes."""
    _fields_ = [
        ("nz", C.c_int32),
        ("H", C.c_float), ("tau", C.c_float), ("f", C.c_float), ("g", C.c_float), ("alpha", C.c_float),
        ("nu0", C.c_float), ("nu_m", C.c_float), ("Ric", C.c_float), ("dRi", C.c_float), ("Pr", C.c_float),
        ("kappa", C.c_float), ("eps", C.c_float),
        ("mu", C.c_float * 6), ("sigma", C.c_float * 6),
        ("flags", C.c_uint32),
        ("diurnal_period", C.c_float),
        ("reserved", C.c_float * 10),
    ]


assert C.sizeof(OpzParams) == 37 * 4


class OpzError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libopz error {code}: {msg}")
        self.code = code


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `make` (or __graft_entry__.build()). "
            "This package has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, fp, i32, i64, f32 = C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_float
    P = C.POINTER(OpzParams)
    lib.opz_version.restype = C.c_int
    lib.opz_last_error.restype = C.c_char_p
    lib.opz_ctx_create.argtypes = [i32, C.POINTER(vp)]
    lib.opz_ctx_destroy.argtypes = [vp]
    lib.opz_ctx_set_stream.argtypes = [vp, vp]
    lib.opz_ctx_synchronize.argtypes = [vp]
    lib.opz_ctx_launch_count.argtypes = [vp]
    lib.opz_ctx_launch_count.restype = i64
    lib.opz_model_create.argtypes = [vp, i32, i32, C.POINTER(C.c_int), i32, fp, fp, fp, C.POINTER(vp)]
    lib.opz_model_set_weights.argtypes = [vp, fp]
    lib.opz_model_get_weights.argtypes = [vp, fp]
    lib.opz_model_set_weights_dev.argtypes = [vp, fp]
    lib.opz_model_n_params.argtypes = [vp]
    lib.opz_model_n_params.restype = i64
    lib.opz_model_destroy.argtypes = [vp]
    lib.opz_rhs.argtypes = [vp, vp, P, fp, fp, f32, i64, i32, fp, fp]
    roll = [vp, vp, P, fp, fp, i64, i32, f32, f32, i32, i32, i32, fp]
    lib.opz_rollout_forward.argtypes = roll
    lib.opz_rollout_forward_dev.argtypes = roll
    lg = [vp, vp, P, fp, fp, i64, i64, i32, f32, f32, i32, i32, fp, fp, i32, fp, fp]
    lib.opz_loss_grad.argtypes = lg
    lib.opz_loss_grad_dev.argtypes = lg
    lib.opz_loss_grad_mpp.argtypes = lg + [fp]
    lib.opz_loss_grad_mpp_dev.argtypes = lg + [fp]
    lib.opz_profile_diagnostics.argtypes = [vp, vp, P, fp, fp, i64, i32, f32, f32, fp, fp, fp, fp, fp]
    lib.opz_flux_loss_grad.argtypes = [vp, vp, P, i32, fp, fp, fp, i64, f32, fp, fp]
    lib.opz_mlp_forward.argtypes = [vp, vp, i32, fp, i64, i32, fp]
    lib.opz_comm_get_unique_id.argtypes = [vp]
    lib.opz_ctx_comm_init.argtypes = [vp, vp, i32, i32]
    lib.opz_ctx_comm_adopt.argtypes = [vp, vp, i32, i32]
    lib.opz_ctx_comm_destroy.argtypes = [vp]
    lib.opz_ctx_comm_info.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int64)]
    lib.opz_comm_nccl_version.argtypes = [C.POINTER(C.c_int)]
    lib.opz_shard_range.argtypes = [i64, i32, i32, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    for name in EXPORTED_SYMBOLS:
        fn = getattr(lib, name)  # raises AttributeError if the symbol is not exported
        if name not in ("opz_last_error", "opz_ctx_launch_count", "opz_model_n_params"):
            fn.restype = C.c_int
    return lib


lib = _load()


def check(rc: int) -> None:
    if rc != OPZ_OK:
        raise OpzError(rc, (lib.opz_last_error() or b"").decode())
