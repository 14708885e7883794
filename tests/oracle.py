"""ctypes binding of oracle/liboracle.so — the CPU oracle (test infrastructure only)."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"
LIB = ORACLE_DIR / "liboracle.so"


def build() -> Path:
    src = [ORACLE_DIR / "bcn_oracle.c", ORACLE_DIR / "bcn_oracle.h", ORACLE_DIR / "ltu_params.h"]
    if not LIB.exists() or any(p.stat().st_mtime > LIB.stat().st_mtime for p in src):
        subprocess.run(["make", "-C", str(ORACLE_DIR), "-s"], check=True)
    return LIB


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        L = C.CDLL(str(build()))
        sz, p, i = C.c_size_t, C.c_void_p, C.c_int
        L.orc_decorrelate.restype = C.c_uint16
        L.orc_decorrelate.argtypes = [C.c_uint16, i]
        L.orc_recorrelate.restype = C.c_uint16
        L.orc_recorrelate.argtypes = [C.c_uint16, i]
        for n in ("bc1", "bc2"):
            for d in ("transform", "untransform"):
                f = getattr(L, f"orc_{n}_{d}")
                f.restype, f.argtypes = None, [p, p, sz, i, i]
        for d in ("transform", "untransform"):
            f = getattr(L, f"orc_bc3_{d}")
            f.restype, f.argtypes = None, [p, p, sz, i, i, i]
        L.orc_split_color_endpoints.restype, L.orc_split_color_endpoints.argtypes = None, [p, p, sz]
        L.orc_ltu_num_lz_matches.restype, L.orc_ltu_num_lz_matches.argtypes = sz, [p, sz]
        L.orc_ltu_estimate.restype, L.orc_ltu_estimate.argtypes = sz, [p, sz]
        L.orc_ltu_num_lz_matches_params.restype, L.orc_ltu_num_lz_matches_params.argtypes = sz, [p, sz, i, i, i]
        L.orc_ltu_set_params.restype, L.orc_ltu_set_params.argtypes = i, [i, i, i]
        ip = C.POINTER(C.c_int)
        L.orc_bc1_transform_auto.restype, L.orc_bc1_transform_auto.argtypes = i, [p, p, sz, i, p, p, ip, ip]
        L.orc_bc2_transform_auto.restype, L.orc_bc2_transform_auto.argtypes = i, [p, p, sz, i, p, p, ip, ip]
        L.orc_bc3_transform_auto.restype, L.orc_bc3_transform_auto.argtypes = i, [p, p, sz, i, p, p, ip, ip, ip]
        for n in ("bc1", "bc2", "bc3"):
            f = getattr(L, f"orc_{n}_auto_estimates")
            f.restype, f.argtypes = i, [p, p, sz, i, C.POINTER(sz)]
            g = getattr(L, f"orc_generate_{n}_test_data")
            g.restype, g.argtypes = None, [p, sz]
        L.orc_bcn_run_mt.restype, L.orc_bcn_run_mt.argtypes = None, [i, i, p, p, sz, i, i, i, i]
        L.orc_bc1_normalize_blocks.restype, L.orc_bc1_normalize_blocks.argtypes = None, [p, p, sz, i]
        L.orc_bc1_normalize_blocks_all_modes.restype, L.orc_bc1_normalize_blocks_all_modes.argtypes = i, [p, p, p, p, sz]
        L.orc_bc1_normalize_split_blocks_in_place.restype = None
        L.orc_bc1_normalize_split_blocks_in_place.argtypes = [p, p, sz, i]
        L.orc_bc1_transform_with_normalize_blocks.restype = None
        L.orc_bc1_transform_with_normalize_blocks.argtypes = [p, p, sz, i, i, i]
        L.orc_bc1_transform_auto_with_normalization.restype = i
        L.orc_bc1_transform_auto_with_normalization.argtypes = [p, p, sz, i, p, p, ip, ip, ip]
        _lib = L
    return _lib


def _ptr(a: np.ndarray) -> int:
    assert a.dtype == np.uint8 and a.flags.c_contiguous
    return a.ctypes.data


def decorrelate(v: int, variant: int) -> int:
    return lib().orc_decorrelate(v, variant)


def recorrelate(v: int, variant: int) -> int:
    return lib().orc_recorrelate(v, variant)


def transform(fmt: int, data: np.ndarray, variant: int, split_alpha: bool, split_colour: bool, threads: int = 1) -> np.ndarray:
    out = np.empty_like(data)
    if threads > 1:
        lib().orc_bcn_run_mt(fmt, 0, _ptr(data), _ptr(out), data.size, variant, int(split_alpha), int(split_colour), threads)
    elif fmt == 3:
        lib().orc_bc3_transform(_ptr(data), _ptr(out), data.size, variant, int(split_alpha), int(split_colour))
    else:
        getattr(lib(), f"orc_bc{fmt}_transform")(_ptr(data), _ptr(out), data.size, variant, int(split_colour))
    return out


def untransform(fmt: int, data: np.ndarray, variant: int, split_alpha: bool, split_colour: bool, threads: int = 1) -> np.ndarray:
    out = np.empty_like(data)
    if threads > 1:
        lib().orc_bcn_run_mt(fmt, 1, _ptr(data), _ptr(out), data.size, variant, int(split_alpha), int(split_colour), threads)
    elif fmt == 3:
        lib().orc_bc3_untransform(_ptr(data), _ptr(out), data.size, variant, int(split_alpha), int(split_colour))
    else:
        getattr(lib(), f"orc_bc{fmt}_untransform")(_ptr(data), _ptr(out), data.size, variant, int(split_colour))
    return out


def generate_test_data(fmt: int, num_blocks: int) -> np.ndarray:
    out = np.empty(num_blocks * (8 if fmt == 1 else 16), np.uint8)
    getattr(lib(), f"orc_generate_bc{fmt}_test_data")(_ptr(out), num_blocks)
    return out


def ltu_estimate(data: np.ndarray) -> int:
    return lib().orc_ltu_estimate(_ptr(data) if data.size else None, data.size)


def ltu_matches(data: np.ndarray) -> int:
    return lib().orc_ltu_num_lz_matches(_ptr(data), data.size)


def ltu_matches_params(data: np.ndarray, hash_bits: int, index_top: bool, group: int) -> int:
    return lib().orc_ltu_num_lz_matches_params(_ptr(data), data.size, hash_bits, int(index_top), group)


def ltu_set_params(hash_bits: int = 16, index_top: bool = True, group: int = 4) -> None:
    """Process-wide parameters of the restated estimator (everything that estimates follows them)."""
    if lib().orc_ltu_set_params(hash_bits, int(index_top), group) != 0:
        raise ValueError("unsupported LTU parameters")


def auto(fmt: int, data: np.ndarray, use_all: bool):
    """Returns (transformed bytes, (variant, split_alpha, split_colour)) with the LTU restatement."""
    out = np.empty_like(data)
    v, sa, sc = C.c_int(), C.c_int(0), C.c_int()
    if fmt == 3:
        rc = lib().orc_bc3_transform_auto(_ptr(data), _ptr(out), data.size, int(use_all), None, None, C.byref(v), C.byref(sa), C.byref(sc))
    else:
        rc = getattr(lib(), f"orc_bc{fmt}_transform_auto")(_ptr(data), _ptr(out), data.size, int(use_all), None, None, C.byref(v), C.byref(sc))
    assert rc == 0
    return out, (v.value, bool(sa.value), bool(sc.value))


def auto_estimates(fmt: int, data: np.ndarray, use_all: bool) -> list[int]:
    scratch = np.empty_like(data)
    sizes = (C.c_size_t * 16)()
    k = getattr(lib(), f"orc_bc{fmt}_auto_estimates")(_ptr(data), _ptr(scratch), data.size, int(use_all), sizes)
    return [sizes[i] for i in range(k)]


def run_range(fmt: int, inverse: bool, src: np.ndarray, dst: np.ndarray, variant: int, split_alpha: bool,
              split_colour: bool, b0: int, b1: int) -> None:
    """Transform / untransform blocks [b0, b1) of the payload `src` into `dst` (full-size buffers)."""
    f = lib().orc_bcn_run_range
    f.restype = None
    f.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_size_t]
    f(fmt, int(inverse), _ptr(src), _ptr(dst), src.size, variant, int(split_alpha), int(split_colour), b0, b1)


# ---- experimental::normalize_blocks (BC1) ---------------------------------------------------------------
def normalize_blocks(data: np.ndarray, mode: int) -> np.ndarray:
    out = np.empty_like(data)
    lib().orc_bc1_normalize_blocks(_ptr(data), _ptr(out), data.size, mode)
    return out


def normalize_blocks_all_modes(data: np.ndarray):
    outs = [np.empty_like(data) for _ in range(3)]
    any_ = lib().orc_bc1_normalize_blocks_all_modes(_ptr(data), _ptr(outs[0]), _ptr(outs[1]), _ptr(outs[2]), data.size)
    return outs, bool(any_)


def normalize_split_blocks_in_place(colors: np.ndarray, indices: np.ndarray, mode: int) -> None:
    lib().orc_bc1_normalize_split_blocks_in_place(_ptr(colors), _ptr(indices), colors.size // 4, mode)


def transform_with_normalize_blocks(data: np.ndarray, norm: int, variant: int, split: bool) -> np.ndarray:
    out = np.empty_like(data)
    lib().orc_bc1_transform_with_normalize_blocks(_ptr(data), _ptr(out), data.size, norm, variant, int(split))
    return out


def auto_with_normalization(data: np.ndarray, use_all: bool):
    """Returns (transformed bytes, (norm, variant, split_colour)) with the LTU restatement."""
    out = np.empty_like(data)
    n, v, s = C.c_int(), C.c_int(), C.c_int()
    rc = lib().orc_bc1_transform_auto_with_normalization(_ptr(data), _ptr(out), data.size, int(use_all), None, None,
                                                         C.byref(n), C.byref(v), C.byref(s))
    assert rc == 0
    return out, (n.value, v.value, bool(s.value))
