"""
ctypes binding of libfatiando_b200.so (C ABI: include/fatiando_b200.h).

The library is built in-tree by :mod:`fatiando_b200.build` (nvcc, sm_100a).
There is no fallback of any kind: if the shared object is missing the import of
the compute API raises, and if no CUDA device is usable every compute call
raises ``RuntimeError`` with the library's own message.
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# FB200_LIBRARY selects another build of the same ABI (kernel tuning experiments)
LIB_PATH = os.environ.get("FB200_LIBRARY") or os.path.join(HERE, "libfatiando_b200.so")

OK, EINVAL, ECUDA, EPROP, EUNSUP = 0, -1, -2, -3, -4
DEVICE_POINTERS, REFERENCE_ORDER, TRANSPOSED, NO_PRISM_SPLIT, PER_CELL = 1, 2, 4, 8, 16

FIELDS = ("potential", "gx", "gy", "gz", "gxx", "gxy", "gxz", "gyy", "gyz",
          "gzz", "tf", "bx", "by", "bz")
FIELD_ID = {name: i for i, name in enumerate(FIELDS)}
GRAVITY_MASK, MAGNETIC_MASK = 0x03FF, 0x3C00

_dp = ctypes.POINTER(ctypes.c_double)
_i64 = ctypes.c_int64
_d = ctypes.c_double

_PRISM_GRAVITY = [_dp, _dp, _dp, _i64] + [_d] * 7 + [_dp]
_SIGNATURES = {
    "fb200_abi_version": ([], ctypes.c_int),
    "fb200_last_error": ([], ctypes.c_char_p),
    "fb200_device_count": ([ctypes.POINTER(ctypes.c_int)], ctypes.c_int),
    "fb200_device_info": ([ctypes.c_int, ctypes.c_char_p, ctypes.c_int,
                           ctypes.POINTER(ctypes.c_int), ctypes.POINTER(_i64)], ctypes.c_int),
    "fb200_model_create": ([ctypes.c_int, _i64] + [_dp] * 10 + [ctypes.POINTER(ctypes.c_void_p)],
                           ctypes.c_int),
    "fb200_model_create_spheres": ([ctypes.c_int, _i64] + [_dp] * 8 + [ctypes.POINTER(ctypes.c_void_p)],
                                   ctypes.c_int),
    "fb200_model_create_polyprisms": ([ctypes.c_int, _i64, ctypes.POINTER(_i64)] + [_dp] * 8
                                      + [ctypes.POINTER(ctypes.c_void_p)], ctypes.c_int),
    "fb200_model_attach_mesh": ([ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
                                + [_dp] * 8, ctypes.c_int),
    "fb200_model_attach_surface": ([ctypes.c_void_p, _i64] + [_dp] * 7 + [_i64] + [_dp] * 5,
                                   ctypes.c_int),
    "fb200_model_set_hazard_planes": ([ctypes.c_void_p, _i64, _dp, _i64, _dp, _i64, _dp], ctypes.c_int),
    "fb200_model_destroy": ([ctypes.c_void_p], ctypes.c_int),
    "fb200_model_size": ([ctypes.c_void_p], _i64),
    "fb200_model_device": ([ctypes.c_void_p], ctypes.c_int),
    "fb200_forward": ([ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, _i64,
                       ctypes.c_uint32, _dp, _dp, _dp, _dp, ctypes.c_void_p, _i64,
                       ctypes.c_uint32], ctypes.c_int),
    "fb200_jacobian": ([ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, _i64,
                        ctypes.c_int, _dp, _dp, _d, ctypes.c_void_p, _i64, ctypes.c_uint32],
                       ctypes.c_int),
    "fb200_adjoint": ([ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, _i64,
                       ctypes.c_int, _dp, _dp, _d, ctypes.c_void_p, ctypes.c_void_p,
                       ctypes.c_uint32], ctypes.c_int),
    "fb200_set_default_device": ([ctypes.c_int], ctypes.c_int),
    "fb200_prism_tf": ([_dp, _dp, _dp, _i64] + [_d] * 12 + [_dp], ctypes.c_int),
    "fb200_prism_bx": ([_dp, _dp, _dp, _i64] + [_d] * 9 + [_dp], ctypes.c_int),
    "fb200_prism_by": ([_dp, _dp, _dp, _i64] + [_d] * 9 + [_dp], ctypes.c_int),
    "fb200_prism_bz": ([_dp, _dp, _dp, _i64] + [_d] * 9 + [_dp], ctypes.c_int),
    "fb200_last_kernel_ms": ([ctypes.c_void_p, ctypes.POINTER(_d), ctypes.POINTER(ctypes.c_int)],
                             ctypes.c_int),
    "fb200_last_fallback_points": ([ctypes.c_void_p, ctypes.POINTER(ctypes.c_int)], ctypes.c_int),
    "fb200_jacobian_grid": ([ctypes.c_void_p, _i64, _i64, ctypes.c_int, ctypes.c_int, _dp, _dp, _dp,
                             ctypes.c_int, _dp, _dp, _d, ctypes.c_void_p, _i64, ctypes.c_uint32], ctypes.c_int),
    "fb200_gram": ([ctypes.c_int, ctypes.c_void_p, _i64, _i64, _i64, ctypes.c_void_p, _d, ctypes.c_void_p,
                    _i64, ctypes.POINTER(_d)], ctypes.c_int),
    "fb200_matvec": ([ctypes.c_int, ctypes.c_void_p, _i64, _i64, _i64, ctypes.c_void_p, ctypes.c_void_p],
                     ctypes.c_int),
    "fb200_matvec_t": ([ctypes.c_int, ctypes.c_void_p, _i64, _i64, _i64, ctypes.c_void_p, ctypes.c_void_p,
                        _d, ctypes.c_void_p], ctypes.c_int),
    "fb200_device_alloc": ([ctypes.c_int, ctypes.c_size_t, ctypes.POINTER(ctypes.c_void_p)], ctypes.c_int),
    "fb200_device_free": ([ctypes.c_int, ctypes.c_void_p], ctypes.c_int),
    "fb200_copy": ([ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int],
                   ctypes.c_int),
    "fb200_host_alloc": ([ctypes.c_size_t, ctypes.POINTER(ctypes.c_void_p)], ctypes.c_int),
    "fb200_host_free": ([ctypes.c_void_p], ctypes.c_int),
    "fb200_fp64_peak": ([ctypes.c_int, _d, ctypes.POINTER(_d), ctypes.POINTER(_d)], ctypes.c_int),
    "fb200_flush_l2": ([ctypes.c_int], ctypes.c_int),
    "fb200_debug_math": ([ctypes.c_int, ctypes.c_int, _dp, _dp, _i64, _dp], ctypes.c_int),
    "fb200_harvest_create": ([ctypes.c_int, ctypes.c_int, ctypes.POINTER(_i64),
                              ctypes.POINTER(ctypes.c_int)] + [_dp] * 7
                             + [ctypes.POINTER(ctypes.c_void_p)], ctypes.c_int),
    "fb200_harvest_destroy": ([ctypes.c_void_p], ctypes.c_int),
    "fb200_harvest_effects": ([ctypes.c_void_p, ctypes.c_int, _dp, _dp, _dp,
                               ctypes.POINTER(ctypes.c_int), ctypes.POINTER(_i64), ctypes.c_uint32],
                              ctypes.c_int),
    "fb200_harvest_accrete": ([ctypes.c_void_p, _i64], ctypes.c_int),
    "fb200_harvest_evaluate": ([ctypes.c_void_p, _i64, ctypes.POINTER(_i64), _dp, _dp], ctypes.c_int),
    "fb200_harvest_predicted": ([ctypes.c_void_p, ctypes.POINTER(ctypes.c_float)], ctypes.c_int),
    "fb200_harvest_effect": ([ctypes.c_void_p, _i64, _dp], ctypes.c_int),
    "fb200_harvest_launches": ([ctypes.c_void_p], _i64),
}
for _name in ("potential", "gx", "gy", "gz", "gxx", "gxy", "gxz", "gyy", "gyz", "gzz"):
    _SIGNATURES["fb200_prism_" + _name] = (_PRISM_GRAVITY, ctypes.c_int)

_lib = None


def lib():
    """The loaded library (raises ImportError when it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "%s is missing: build it with `python -m fatiando_b200.build` "
                "(fatiando_b200 has no CPU fallback)" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (argtypes, restype) in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.argtypes = argtypes
            fn.restype = restype
        _lib = L
    return _lib


def last_error():
    msg = lib().fb200_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc):
    """Turn a C-ABI return code into the Python exception the API promises."""
    if rc == OK:
        return
    msg = last_error()
    if rc in (EINVAL, EPROP, EUNSUP):
        raise ValueError(msg)
    raise RuntimeError("fatiando_b200 CUDA failure: " + msg)


def dptr(a):
    """Pointer to the data of a C-contiguous float64 array (or None)."""
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(_dp)


def device_count():
    n = ctypes.c_int(0)
    rc = lib().fb200_device_count(ctypes.byref(n))
    return n.value if rc == OK else 0


def require_device():
    n = ctypes.c_int(0)
    check(lib().fb200_device_count(ctypes.byref(n)))
    return n.value


# Page-locked buffers are expensive to create (the OS pins every page) and cheap to reuse: a
# released buffer goes back to this pool and serves the next request of the same size.
_PINNED_POOL = {}
_PINNED_POOL_BYTES = [0]
PINNED_POOL_LIMIT = 32 << 30          # bytes kept for reuse; beyond that buffers are freed


def _pinned_release(nbytes, address):
    if _lib is None:
        return
    if _PINNED_POOL_BYTES[0] + nbytes <= PINNED_POOL_LIMIT:
        _PINNED_POOL.setdefault(nbytes, []).append(address)
        _PINNED_POOL_BYTES[0] += nbytes
    else:
        _lib.fb200_host_free(ctypes.c_void_p(address))


def pinned_pool_clear():
    """Free every pooled page-locked buffer."""
    for nbytes, addresses in list(_PINNED_POOL.items()):
        for address in addresses:
            _lib.fb200_host_free(ctypes.c_void_p(address))
    _PINNED_POOL.clear()
    _PINNED_POOL_BYTES[0] = 0


def pinned_empty(shape, dtype="float64"):
    """
    An uninitialised numpy array in page-locked host memory (fb200_host_alloc): a destination
    that ``Model.jacobian(..., out=...)`` fills at PCIe DMA speed (pageable memory goes through
    the driver's bounce buffers, and a fresh pageable array pays a page fault per 4 KB on top).
    When the array and every view of it are gone the buffer returns to a pool and serves the
    next request of the same size.
    """
    import weakref
    import numpy
    shape = tuple(int(v) for v in (shape if isinstance(shape, (tuple, list)) else (shape,)))
    dt = numpy.dtype(dtype)
    nbytes = max(1, int(numpy.prod(shape, dtype=numpy.int64)) * dt.itemsize)
    pooled = _PINNED_POOL.get(nbytes)
    if pooled:
        address = pooled.pop()
        _PINNED_POOL_BYTES[0] -= nbytes
    else:
        ptr = ctypes.c_void_p()
        check(lib().fb200_host_alloc(nbytes, ctypes.byref(ptr)))
        address = ptr.value
    buf = (ctypes.c_char * nbytes).from_address(address)
    arr = numpy.frombuffer(buf, dtype=dt, count=nbytes // dt.itemsize).reshape(shape)
    weakref.finalize(buf, _pinned_release, nbytes, address)
    return arr
