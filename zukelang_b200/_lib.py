"""ctypes binding of libzkb200.so — the same symbols the OCaml foreign_stubs bind
(INTEGRATION.md).  Fails loudly when the library is missing: there is no
fallback implementation."""

from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_size_t, c_uint8, c_uint32, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# ZKB200_LIB selects another build of the same library (libzkb200_checked.so: device-side asserts on)
LIB_PATH = os.environ.get("ZKB200_LIB") or os.path.join(_HERE, "libzkb200.so")

ZK_OK, ZK_EARG, ZK_EPOINT, ZK_ECUDA, ZK_EREMAINDER = 0, -1, -2, -3, -4
FR_BYTES, G1_RAW, G1_COMP, G1_OUT, G2_RAW, G2_COMP, G2_OUT = 32, 96, 48, 144, 192, 96, 288
GROTH16_PROOF_OUT = G1_OUT + G2_OUT + G1_OUT
PINOCCHIO_PROOF_OUT = 6 * G1_OUT + 2 * G2_OUT
GT_BYTES = 576


class ZkError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("libzkb200 error %d: %s" % (code, msg))
        self.code = code


class InvalidArgument(ZkError, ValueError):
    """OCaml's Invalid_argument (ZK_EARG)."""


_u8p = POINTER(c_uint8)
_u32p = POINTER(c_uint32)

# name -> (restype, argtypes); mirrors include/zkb200.h one to one
SIGNATURES = {
    "zk_init": (c_int, [c_int]),
    "zk_init_devices": (c_int, [POINTER(c_int), c_int]),
    "zk_device_count": (c_int, []),
    "zk_shutdown": (c_int, []),
    "zk_last_error": (c_char_p, []),
    "zk_device_info": (c_int, [c_char_p, c_size_t]),
    "zk_g1_msm": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "zk_g2_msm": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "zk_g1_table_load": (c_int, [c_void_p, c_void_p, c_size_t, c_int, c_int, POINTER(c_uint64)]),
    "zk_g2_table_load": (c_int, [c_void_p, c_void_p, c_size_t, c_int, c_int, POINTER(c_uint64)]),
    "zk_g1_table_msm": (c_int, [c_uint64, c_void_p, c_size_t, c_void_p]),
    "zk_g2_table_msm": (c_int, [c_uint64, c_void_p, c_size_t, c_void_p]),
    "zk_g1_table_msm_batch": (c_int, [c_uint64, c_void_p, c_size_t, c_size_t, c_void_p]),
    "zk_g2_table_msm_batch": (c_int, [c_uint64, c_void_p, c_size_t, c_size_t, c_void_p]),
    "zk_g1_table_msm_dev": (c_int, [c_uint64, c_void_p, c_size_t, c_void_p, c_void_p]),
    "zk_g2_table_msm_dev": (c_int, [c_uint64, c_void_p, c_size_t, c_void_p, c_void_p]),
    "zk_table_info": (c_int, [c_uint64, POINTER(c_uint64)]),
    "zk_table_pipeline": (c_int, [c_uint64, c_int]),
    "zk_table_profile_totals": (c_int, [c_uint64, POINTER(ctypes.c_float), POINTER(ctypes.c_uint64)]),
    "zk_table_join": (c_int, [c_uint64, c_void_p]),
    "zk_table_profile": (c_int, [c_uint64, c_int, c_void_p]),
    "zk_table_batch_timing": (c_int, [c_uint64, c_int, c_void_p, c_size_t, POINTER(c_size_t)]),
    "zk_table_free": (c_int, [c_uint64]),
    "zk_g1_sum": (c_int, [c_void_p, c_size_t, c_void_p]),
    "zk_g2_sum": (c_int, [c_void_p, c_size_t, c_void_p]),
    "zk_g1_sum_dev": (c_int, [c_void_p, c_size_t, c_void_p, c_void_p]),
    "zk_g2_sum_dev": (c_int, [c_void_p, c_size_t, c_void_p, c_void_p]),
    "zk_g1_sum_strided_dev": (c_int, [c_void_p, c_size_t, c_size_t, c_void_p, c_void_p]),
    "zk_g2_sum_strided_dev": (c_int, [c_void_p, c_size_t, c_size_t, c_void_p, c_void_p]),
    "zk_g1_fixed_base_mul": (c_int, [c_void_p, c_size_t, c_void_p]),
    "zk_g2_fixed_base_mul": (c_int, [c_void_p, c_size_t, c_void_p]),
    "zk_qap_load": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_size_t, POINTER(c_uint64)]),
    "zk_quotient_domain_load": (c_int, [c_void_p, c_size_t, POINTER(c_uint64)]),
    "zk_qap_eval": (c_int, [c_uint64, c_void_p, c_void_p, c_void_p]),
    "zk_qap_free": (c_int, [c_uint64]),
    "zk_fr_quotient": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "zk_groth16_pk_load": (c_int, [c_void_p, c_int, c_int, POINTER(c_uint64)]),
    "zk_groth16_prove": (c_int, [c_uint64, c_uint64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "zk_groth16_prove_coeffs": (c_int, [c_uint64, c_uint64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "zk_groth16_combine": (c_int, [c_void_p, c_size_t, c_void_p]),
    "zk_groth16_last_device_ms": (c_int, [c_uint64, POINTER(ctypes.c_float)]),
    "zk_groth16_last_stage_ms": (c_int, [c_uint64, POINTER(ctypes.c_float)]),
    "zk_eval_domain_load": (c_int, [c_size_t, c_void_p, c_void_p, POINTER(c_uint64)]),
    "zk_r1cs_load": (c_int, [c_uint64, c_int, c_size_t, c_void_p, c_void_p, c_void_p]),
    "zk_groth16_prove_r1cs": (c_int, [c_uint64, c_uint64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "zk_pinocchio_pk_load": (c_int, [c_void_p, c_int, c_int, POINTER(c_uint64)]),
    "zk_pinocchio_prove": (c_int, [c_uint64, c_uint64, c_void_p, c_void_p, c_void_p]),
    "zk_key_free": (c_int, [c_uint64]),
    "zk_pairing_product": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "zk_pairing_product_batch": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "zk_gt_mul": (c_int, [c_void_p, c_void_p, c_void_p]),
    "zk_g1_decompress": (c_int, [c_void_p, c_size_t, c_void_p]),
    "zk_g2_decompress": (c_int, [c_void_p, c_size_t, c_void_p]),
    "zk_bench_intpipe": (c_int, [c_int, c_int, POINTER(c_double), POINTER(c_double)]),
    "zk_test_field_op": (c_int, [c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t]),
    "zk_test_g1_madd": (c_int, [c_void_p, c_void_p, c_int, c_size_t, c_void_p]),
}

class Groth16PKeyStruct(ctypes.Structure):
    """zk_groth16_pkey of include/zkb200.h."""
    _fields_ = [("n", c_size_t), ("m", c_size_t), ("n_mid", c_size_t), ("n_h", c_size_t), ("mid_index", c_void_p),
                ("a", c_void_p), ("b1", c_void_p), ("d1", c_void_p), ("b2", c_void_p), ("d2", c_void_p),
                ("ti1", c_void_p), ("ti2", c_void_p), ("tiztd", c_void_p), ("ltd_mid", c_void_p)]


class PinocchioPKeyStruct(ctypes.Structure):
    """zk_pinocchio_pkey of include/zkb200.h."""
    _fields_ = [("n", c_size_t), ("m", c_size_t), ("n_mid", c_size_t), ("mid_index", c_void_p),
                ("vv", c_void_p), ("yy", c_void_p), ("vav", c_void_p), ("yay", c_void_p), ("bvwy", c_void_p),
                ("ww", c_void_p), ("waw", c_void_p), ("si", c_void_p), ("v_all", c_void_p), ("w_all", c_void_p),
                ("one", c_void_p), ("vt", c_void_p), ("yt", c_void_p), ("vavt", c_void_p), ("yayt", c_void_p),
                ("vbt", c_void_p), ("wbt", c_void_p), ("ybt", c_void_p), ("wt", c_void_p), ("wawt", c_void_p)]


_lib = None
_initialised = False


def load(path: str | None = None) -> ctypes.CDLL:
    """dlopen libzkb200.so and declare every prototype.  No device needed."""
    global _lib
    if _lib is not None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise ImportError(
            "libzkb200.so not found at %s — build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C zukelang_b200/csrc`; there is no CPU fallback" % p)
    lib = ctypes.CDLL(p)
    for name, (res, args) in SIGNATURES.items():
        if not hasattr(lib, name):
            continue                      # reported by tests/test_abi.py
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc == ZK_OK:
        return
    msg = (load().zk_last_error() or b"").decode("utf-8", "replace")
    if rc == ZK_EARG:
        raise InvalidArgument(rc, msg)
    raise ZkError(rc, msg)


def lib() -> ctypes.CDLL:
    """The initialised library.  ZKB200_DEVICES="0,1,2,3" drives several devices from this one
    process (zk_init_devices); otherwise the device of ZKB200_DEVICE / LOCAL_RANK, else device 0."""
    global _initialised
    l = load()
    if not _initialised:
        many = os.environ.get("ZKB200_DEVICES", "")
        if many:
            init_devices([int(x) for x in many.split(",")])
        else:
            dev = int(os.environ.get("ZKB200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
            check(l.zk_init(dev))
        _initialised = True
    return l


def init_devices(devs) -> ctypes.CDLL:
    """zk_init_devices: drive these CUDA devices from this process (devs[0] = primary)."""
    global _initialised
    l = load()
    arr = (c_int * len(devs))(*devs)
    check(l.zk_init_devices(arr, len(devs)))
    _initialised = True
    return l


def buf(data) -> ctypes.Array:
    """bytes-like -> ctypes array usable as a void* argument (zero copy for bytearray)."""
    if isinstance(data, (bytes, bytearray, memoryview)):
        b = bytes(data) if not isinstance(data, bytes) else data
        return ctypes.create_string_buffer(b, len(b))
    return data
