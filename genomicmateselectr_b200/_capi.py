"""ctypes binding of include/gms_b200.h (the C-ABI drop-in boundary).

There is no CPU fallback: if libgms_b200.so is missing this module raises, and if no CUDA
device is present `gms_create` fails with GMS_ECUDA."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# GMS_B200_LIB: profiling hook (tools/): load another build of the same library, e.g. one compiled with -DGMS_PROFILE
LIB_PATH = os.environ.get("GMS_B200_LIB") or os.path.join(_HERE, "lib", "libgms_b200.so")

GMS_OK, GMS_EINVAL, GMS_ECUDA, GMS_ENOMEM, GMS_ESTATE = 0, -1, -2, -3, -4
MODEL = {"A": 0, "AD": 1, "DirDom": 2}
ERRNAME = {-1: "GMS_EINVAL", -2: "GMS_ECUDA", -3: "GMS_ENOMEM", -4: "GMS_ESTATE"}


class GmsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{ERRNAME.get(code, code)}: {msg}")
        self.code = code


class Stats(C.Structure):
    _fields_ = [("m", C.c_int64), ("m_padded", C.c_int64), ("S2", C.c_int64), ("tile_bytes", C.c_int64),
                ("C", C.c_int32), ("P", C.c_int32), ("last_T", C.c_int32), ("last_model", C.c_int32),
                ("last_N", C.c_int64), ("last_parents_used", C.c_int64), ("last_launches", C.c_int32), ("last_mode", C.c_int32),
                ("last_ms_parent_stage", C.c_float), ("last_ms_cross_stage", C.c_float),
                ("last_ms_cross_kernel", C.c_float), ("last_ms_assemble", C.c_float),
                ("last_ms_total", C.c_float), ("last_flops_cross", C.c_double),
                ("last_flops_parent", C.c_double), ("last_fallback_units", C.c_int64),
                ("last_parents_computed", C.c_int64), ("last_fallback_overflow", C.c_int32), ("last_cache_hit", C.c_int32)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_P = C.c_void_p
_SIGS = {
    "gms_version": (C.c_char_p, []),
    "gms_last_error": (C.c_char_p, [_P]),
    "gms_create": (C.c_int, [C.POINTER(_P), C.c_int]),
    "gms_destroy": (None, [_P]),
    "gms_set_stream": (C.c_int, [_P, _P]),
    "gms_set_genmap": (C.c_int, [_P, C.c_int, _P, _P]),
    "gms_set_recomb_blocks": (C.c_int, [_P, C.c_int, _P, _P]),
    "gms_set_haplotypes_i32cm": (C.c_int, [_P, C.c_int64, C.c_int64, _P]),
    "gms_set_haplotypes_u8rm": (C.c_int, [_P, C.c_int64, C.c_int64, _P]),
    "gms_set_chromosome_shard": (C.c_int, [_P, C.c_int, C.c_int]),
    "gms_set_mode": (C.c_int, [_P, C.c_int]),
    "gms_set_int8_tolerance": (C.c_int, [_P, C.c_double]),
    "gms_set_cache": (C.c_int, [_P, C.c_int]),
    "gms_set_parent_shard": (C.c_int, [_P, C.c_int, C.c_int]),
    "gms_stage_inputs_range": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, _P, C.c_int64, _P, _P, C.c_int64, C.c_int64]),
    "gms_compute_parents": (C.c_int, [_P]),
    "gms_parent_exchange": (C.c_int, [_P, C.POINTER(_P), C.POINTER(C.c_int64), C.POINTER(C.c_int32)]),
    "gms_compute_crosses": (C.c_int, [_P, C.c_int]),
    "gms_device_outputs": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_P), C.POINTER(C.c_int64)]),
    "gms_predict_batched": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, _P, _P, C.c_int64, _P, _P, _P, _P]),
    "gms_trait_pairs": (C.c_int, [C.c_int, _P, _P]),
    "gms_predict": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, _P, C.c_int64, _P, _P, _P, _P]),
    "gms_predict_sharded": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, _P, _P, C.c_int64, _P, _P, _P, _P]),
    "gms_stage_inputs": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, _P, C.c_int64, _P, _P]),
    "gms_compute": (C.c_int, [_P, C.c_int]),
    "gms_fetch_outputs": (C.c_int, [_P, _P, _P]),
    "gms_selind": (C.c_int, [C.c_int, C.c_int, C.c_int64, _P, _P, _P]),
    "gms_tidy": (C.c_int, [_P, _P, C.c_double, _P]),
    "gms_cross_means": (C.c_int, [_P, C.c_int, _P, _P, C.c_int64, _P, _P, _P, _P, _P]),
    "gms_get_stats": (C.c_int, [_P, C.POINTER(Stats)]),
}
EXPORTS = sorted(_SIGS)

_lib = None


def lib():
    """Load libgms_b200.so (built in-tree by `python -m genomicmateselectr_b200.build`)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m genomicmateselectr_b200.build` "
                "(nvcc, sm_100a). There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib
