"""ctypes binding of libsqcc_b200.so (include/sqcc_b200.h).  No CPU fallback: a missing library is an error."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SQCC_LIB") or os.path.join(_HERE, "lib", "libsqcc_b200.so")   # SQCC_LIB: A/B builds (tuning)

c_int, c_i64, c_u64, c_void_p = ctypes.c_int, ctypes.c_int64, ctypes.c_uint64, ctypes.c_void_p
c_double_p = ctypes.POINTER(ctypes.c_double)
c_i64_p = ctypes.POINTER(ctypes.c_int64)

# name -> (restype, argtypes); mirrors include/sqcc_b200.h declaration for declaration
SIGNATURES = {
    "sqcc_version": (c_int, []),
    "sqcc_last_error": (ctypes.c_char_p, []),
    "sqcc_device_count": (c_int, [ctypes.POINTER(c_int)]),
    "sqcc_ctx_create": (c_int, [c_int, ctypes.POINTER(c_void_p)]),
    "sqcc_ctx_destroy": (c_int, [c_void_p]),
    "sqcc_ctx_set_stream": (c_int, [c_void_p, c_void_p]),
    "sqcc_ctx_synchronize": (c_int, [c_void_p]),
    "sqcc_npair": (c_i64, [c_int]),
    "sqcc_nquad": (c_i64, [c_int]),
    "sqcc_tile_count": (c_i64, [c_int]),
    "sqcc_shard_tiles": (c_int, [c_int, c_int, c_int, c_i64_p, c_i64_p]),
    "sqcc_shard_tiles_weighted": (c_int, [c_int, c_int, c_int, c_double_p, c_i64_p, c_i64_p]),
    "sqcc_tile_rows": (c_int, [c_int, c_i64, c_i64, ctypes.POINTER(c_int), ctypes.POINTER(c_int)]),
    "sqcc_native_len_tiles": (c_i64, [c_int, c_i64, c_i64]),
    "sqcc_native_offset": (c_i64, [c_int, c_i64, c_int, c_int, c_int, c_int]),
    "sqcc_eri_attach_tiles": (c_int, [c_void_p, c_void_p, c_int, c_i64, c_i64]),
    "sqcc_eri_synth": (c_int, [c_void_p, c_u64]),
    "sqcc_eri_pack_full_slab": (c_int, [c_void_p, c_void_p, c_int, c_int]),
    "sqcc_eri_pack_full_host": (c_int, [c_void_p, c_void_p]),
    "sqcc_eri_pack_canonical": (c_int, [c_void_p, c_void_p, c_i64, c_i64]),
    "sqcc_eri_clear": (c_int, [c_void_p]),
    "sqcc_eri_pack_blocks": (c_int, [c_void_p, c_void_p, c_void_p, c_int]),
    "sqcc_eri_unpack_full": (c_int, [c_void_p, c_void_p]),
    "sqcc_jk_build": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "sqcc_jk_build_host": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "sqcc_jk_set_variant": (c_int, [c_void_p, c_int]),
    "sqcc_jk_set_deterministic": (c_int, [c_void_p, c_int]),
    "sqcc_eri_absmax": (c_int, [c_void_p, c_double_p, c_int]),
    "sqcc_comm_export": (c_int, [c_void_p, c_void_p]),
    "sqcc_comm_attach": (c_int, [c_void_p, c_int, c_int, c_void_p]),
    "sqcc_comm_attach_local": (c_int, [ctypes.POINTER(c_void_p), c_int]),
    "sqcc_comm_detach": (c_int, [c_void_p]),
    "sqcc_comm_status": (c_int, [c_void_p]),
    "sqcc_comm_reset_status": (c_int, [c_void_p]),
    "sqcc_comm_set_enabled": (c_int, [c_void_p, c_int]),
    "sqcc_comm_stamps": (c_int, [c_void_p, c_void_p]),
    "sqcc_jk_partial": (c_int, [c_void_p, c_void_p, c_int, c_int]),
    "sqcc_jk_reduce_local": (c_int, [ctypes.POINTER(c_void_p), c_int, c_int, c_int, ctypes.POINTER(c_void_p), ctypes.POINTER(c_void_p)]),
    "sqcc_grid_attach": (c_int, [c_void_p, c_void_p, c_void_p, c_i64, c_int]),
    "sqcc_lda_vxc": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "sqcc_lda_vxc_host": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "sqcc_grid_quadrature": (c_int, [c_void_p, c_void_p, c_void_p]),
    "sqcc_grid_rho": (c_int, [c_void_p, c_void_p, c_void_p]),
    "sqcc_mp2": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "sqcc_mp2_host": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "sqcc_ao2mo": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p]),
    "sqcc_scf_setup": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int]),
    "sqcc_scf_set_density": (c_int, [c_void_p, c_void_p]),
    "sqcc_scf_fock_energy": (c_int, [c_void_p, c_double_p]),
    "sqcc_scf_jk": (c_int, [c_void_p, ctypes.POINTER(c_void_p), ctypes.POINTER(ctypes.c_int64)]),
    "sqcc_scf_finish": (c_int, [c_void_p, c_double_p]),
    "sqcc_scf_update": (c_int, [c_void_p]),
    "sqcc_scf_get": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "sqcc_get_timings": (c_int, [c_void_p, c_double_p, c_int]),
    "sqcc_launch_count": (c_i64, [c_void_p]),
    "sqcc_eri_bytes": (c_int, [c_void_p, c_i64_p, c_i64_p]),
}

_lib = None


class SqccError(RuntimeError):
    pass


def load():
    """Load the shared library (built by sqcc_b200/build.py).  Raises if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SqccError(f"{LIB_PATH} not found: run `python -m sqcc_b200.build` (or __graft_entry__.build()). "
                            "There is no CPU fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(status):
    if status < 0:
        raise SqccError(f"libsqcc_b200 error {status}: {load().sqcc_last_error().decode()}")
    return status
