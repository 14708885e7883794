"""Test helper: the system's libzstd called directly (ctypes) with the reference's parameters
(extensions/compressors/dxt-lossless-transform-zstd/src/lib.rs:150-209) — the checker for the zstd estimator tests."""
import ctypes as C
import ctypes.util

import numpy as np


def _load():
    for name in ("libzstd.so.1", ctypes.util.find_library("zstd")):
        if not name:
            continue
        try:
            return C.CDLL(name)
        except OSError:
            continue
    return None


_L = _load()
if _L is not None:
    _L.ZSTD_createCCtx.restype = C.c_void_p
    _L.ZSTD_freeCCtx.argtypes = [C.c_void_p]
    _L.ZSTD_CCtx_setParameter.argtypes = [C.c_void_p, C.c_int, C.c_int]
    _L.ZSTD_CCtx_setParameter.restype = C.c_size_t
    _L.ZSTD_compress2.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
    _L.ZSTD_compress2.restype = C.c_size_t
    _L.ZSTD_compressBound.argtypes = [C.c_size_t]
    _L.ZSTD_compressBound.restype = C.c_size_t
    _L.ZSTD_isError.argtypes = [C.c_size_t]
    _L.ZSTD_versionNumber.restype = C.c_uint


def available() -> bool:
    return _L is not None


def version() -> int:
    return int(_L.ZSTD_versionNumber())


def compress_bound(n: int) -> int:
    return int(_L.ZSTD_compressBound(n))


def compressed_size(data: np.ndarray, level: int) -> int:
    data = np.ascontiguousarray(data, dtype=np.uint8)
    if data.size == 0:
        return 0
    cctx = _L.ZSTD_createCCtx()
    try:
        _L.ZSTD_CCtx_setParameter(cctx, 100, level)   # ZSTD_c_compressionLevel
        _L.ZSTD_CCtx_setParameter(cctx, 10, 1)        # ZSTD_c_format = ZSTD_f_zstd1_magicless
        _L.ZSTD_CCtx_setParameter(cctx, 200, 0)       # contentSizeFlag
        _L.ZSTD_CCtx_setParameter(cctx, 201, 0)       # checksumFlag
        _L.ZSTD_CCtx_setParameter(cctx, 202, 0)       # dictIDFlag
        cap = compress_bound(data.size)
        out = np.empty(cap, np.uint8)
        r = _L.ZSTD_compress2(cctx, out.ctypes.data, cap, data.ctypes.data, data.size)
        assert not _L.ZSTD_isError(r)
        return int(r)
    finally:
        _L.ZSTD_freeCCtx(cctx)
