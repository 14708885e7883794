"""ctypes binding of libdxt_lossless_transform_cuda.so (the C ABI in include/*.h).

There is no CPU fallback: if the shared library is missing this module raises at import of the
symbols, and every compute entry point needs a CUDA device.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import os

# DLT_LIB_PATH lets experiments load an alternative build of the same library (never a CPU fallback).
LIB_PATH = Path(os.environ.get("DLT_LIB_PATH") or Path(__file__).resolve().parent / "libdxt_lossless_transform_cuda.so")

_lib = None


class NativeLibraryMissing(RuntimeError):
    pass


# --- shared C types -------------------------------------------------------------------------------
class DltResult(C.Structure):
    _fields_ = [("error_code", C.c_int32)]


MaxCompressedSizeFn = C.CFUNCTYPE(C.c_uint32, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t))
EstimateCompressedSizeFn = C.CFUNCTYPE(
    C.c_uint32, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)
)


class DltSizeEstimator(C.Structure):
    _fields_ = [
        ("context", C.c_void_p),
        ("max_compressed_size", MaxCompressedSizeFn),
        ("estimate_compressed_size", EstimateCompressedSizeFn),
    ]


class CoreSettings(C.Structure):  # Dltbc{1,2}TransformSettings of the core crates
    _fields_ = [("split_colour_endpoints", C.c_bool), ("decorrelation_mode", C.c_uint8)]


class CoreAutoSettings(C.Structure):
    _fields_ = [("use_all_modes", C.c_bool)]


class CoreBc3Settings(C.Structure):
    _fields_ = [
        ("split_alpha_endpoints", C.c_bool),
        ("split_colour_endpoints", C.c_bool),
        ("decorrelation_mode", C.c_uint8),
    ]


class DltcudaSettings(C.Structure):
    _fields_ = [
        ("format", C.c_uint8),
        ("decorrelation_mode", C.c_uint8),
        ("split_alpha_endpoints", C.c_bool),
        ("split_colour_endpoints", C.c_bool),
    ]


class DltcudaPayload(C.Structure):
    _fields_ = [("input", C.c_void_p), ("output", C.c_void_p), ("len", C.c_size_t), ("settings", DltcudaSettings)]


class DltcudaAutoJob(C.Structure):
    _fields_ = [("format", C.c_uint8), ("input", C.c_void_p), ("output", C.c_void_p), ("len", C.c_size_t),
                ("out_settings", DltcudaSettings), ("status", C.c_int32)]


class DltffResult(C.Structure):  # dxt_lossless_transform_file_formats.h
    _fields_ = [("error_code", C.c_int32), ("detail_a", C.c_size_t), ("detail_b", C.c_size_t)]


class DdsInfo(C.Structure):  # dds/parse_dds.rs:36-42
    _fields_ = [("format", C.c_uint8), ("data_offset", C.c_uint8), ("data_length", C.c_uint32)]


class DltddsFile(C.Structure):
    _fields_ = [("input", C.c_void_p), ("input_len", C.c_size_t), ("output", C.c_void_p), ("output_len", C.c_size_t)]


# Every exported symbol: name -> (restype, argtypes).  tests/test_cabi_symbols.py checks this table
# against include/*.h and against the built library.
_P = C.c_void_p
_SZ = C.c_size_t
SIGNATURES: dict[str, tuple] = {}
for _n in (1, 2):
    _p = f"dltbc{_n}_"
    SIGNATURES.update(
        {
            _p + "new_ManualTransformBuilder": (_P, []),
            _p + "free_ManualTransformBuilder": (None, [_P]),
            _p + "clone_ManualTransformBuilder": (_P, [_P]),
            _p + "ManualTransformBuilder_SetDecorrelationMode": (None, [_P, C.c_uint8]),
            _p + "ManualTransformBuilder_SetSplitColourEndpoints": (None, [_P, C.c_bool]),
            _p + "ManualTransformBuilder_ResetToDefaults": (None, [_P]),
            _p + "ManualTransformBuilder_Transform": (DltResult, [_P, _SZ, _P, _SZ, _P]),
            _p + "ManualTransformBuilder_Untransform": (DltResult, [_P, _SZ, _P, _SZ, _P]),
            _p + "new_AutoTransformBuilder": (_P, [C.POINTER(DltSizeEstimator)]),
            _p + "free_AutoTransformBuilder": (None, [_P]),
            _p + "AutoTransformBuilder_SetUseAllDecorrelationModes": (DltResult, [_P, C.c_bool]),
            _p + "AutoTransformBuilder_Transform": (DltResult, [_P, _P, _SZ, _P, _SZ, C.POINTER(_P)]),
            _p + "error_message": (C.c_char_p, [C.c_int32]),
            f"dltbc{_n}core_transform": (DltResult, [_P, _SZ, _P, _SZ, CoreSettings]),
            f"dltbc{_n}core_untransform": (DltResult, [_P, _SZ, _P, _SZ, CoreSettings]),
            f"dltbc{_n}core_transform_auto": (
                DltResult,
                [_P, _SZ, _P, _SZ, C.POINTER(DltSizeEstimator), CoreAutoSettings, C.POINTER(CoreSettings)],
            ),
        }
    )
SIGNATURES.update(
    {
        "dltbc3core_transform": (DltResult, [_P, _SZ, _P, _SZ, CoreBc3Settings]),
        "dltbc3core_untransform": (DltResult, [_P, _SZ, _P, _SZ, CoreBc3Settings]),
        "dltbc3core_transform_auto": (
            DltResult,
            [_P, _SZ, _P, _SZ, C.POINTER(DltSizeEstimator), CoreAutoSettings, C.POINTER(CoreBc3Settings)],
        ),
        "dltltu_new_size_estimator": (C.POINTER(DltSizeEstimator), []),
        "dltltu_free_size_estimator": (None, [C.POINTER(DltSizeEstimator)]),
        "dltzstd_new_size_estimator": (C.POINTER(DltSizeEstimator), [C.c_int]),
        "dltzstd_free_size_estimator": (None, [C.POINTER(DltSizeEstimator)]),
        "dltzstd_version_number": (C.c_uint, []),
        "dltcuda_ManualTransformBuilder_GetSettings": (C.c_int, [_P, C.POINTER(C.c_uint8), C.POINTER(C.c_bool)]),
        "dltcuda_device_count": (C.c_int, []),
        "dltcuda_set_device": (None, [C.c_int]),
        "dltcuda_last_error": (C.c_char_p, []),
        "dltcuda_kernel_launch_count": (C.c_uint64, []),
        "dltcuda_alloc_pinned": (_P, [_SZ]),
        "dltcuda_free_pinned": (None, [_P]),
        "dltcuda_transform_device": (C.c_int, [_P, _P, _SZ, DltcudaSettings, _P]),
        "dltcuda_untransform_device": (C.c_int, [_P, _P, _SZ, DltcudaSettings, _P]),
        "dltcuda_transform_device_range": (C.c_int, [_P, _P, _SZ, _SZ, _SZ, DltcudaSettings, _P]),
        "dltcuda_untransform_device_range": (C.c_int, [_P, _P, _SZ, _SZ, _SZ, DltcudaSettings, _P]),
        "dltcuda_transform_device_streams": (C.c_int, [_P, C.POINTER(_P), _SZ, DltcudaSettings, _P]),
        "dltcuda_untransform_device_streams": (C.c_int, [C.POINTER(_P), _P, _SZ, DltcudaSettings, _P]),
        "dltcuda_stream_count": (C.c_int, [DltcudaSettings]),
        "dltcuda_stream_width": (C.c_int, [DltcudaSettings, C.c_int]),
        "dltcuda_split_color_endpoints_device": (C.c_int, [_P, _P, _SZ, _P]),
        "dltcuda_split_color_endpoints": (C.c_int, [_P, _P, _SZ]),
        "dltcuda_transform_batch": (C.c_int, [C.POINTER(DltcudaPayload), _SZ, C.c_bool]),
        "dltcuda_transform_batch_multi_gpu": (C.c_int, [C.POINTER(DltcudaPayload), _SZ, C.c_bool, C.POINTER(C.c_int), C.c_int]),
        "dltcuda_shard_first_block": (_SZ, [C.c_int, _SZ, C.c_int, C.c_int]),
        "dltcuda_ltu_estimate_device": (C.c_int, [_P, _SZ, C.POINTER(_SZ)]),
        "dltcuda_release_cached_memory": (_SZ, []),
        "dltcuda_ltu_set_params": (C.c_int, [C.c_int, C.c_bool, C.c_int]),
        "dltcuda_ltu_get_params": (None, [C.POINTER(C.c_int), C.POINTER(C.c_bool), C.POINTER(C.c_int)]),
        "dltcuda_transform_auto_device": (
            C.c_int,
            [C.c_int, _P, _P, _SZ, C.c_bool, C.POINTER(DltcudaSettings), C.POINTER(_SZ)],
        ),
        "dltcuda_auto_candidates": (C.c_int, [C.c_int, C.c_bool, C.POINTER(DltcudaSettings)]),
        "dltcuda_bc1_normalize_blocks": (C.c_int, [_P, _P, _SZ, C.c_int]),
        "dltcuda_bc1_normalize_blocks_device": (C.c_int, [_P, _P, _SZ, C.c_int, _P]),
        "dltcuda_bc1_normalize_blocks_all_modes": (C.c_int, [_P, _P, _P, _P, _SZ, C.POINTER(C.c_bool)]),
        "dltcuda_bc1_normalize_split_blocks_in_place": (C.c_int, [_P, _P, _SZ, C.c_int]),
        "dltcuda_bc1_transform_with_normalize_blocks": (C.c_int, [_P, _P, _SZ, C.c_int, C.c_uint8, C.c_bool]),
        "dltcuda_bc1_transform_with_normalize_blocks_device": (C.c_int, [_P, _P, _SZ, C.c_int, C.c_uint8, C.c_bool, _P]),
        "dltcuda_bc1_transform_auto_with_normalization": (
            C.c_int, [_P, _P, _SZ, C.c_bool, C.POINTER(C.c_int), C.POINTER(C.c_uint8), C.POINTER(C.c_bool), C.POINTER(_SZ)]),
        "dltcuda_transform_auto_batch": (C.c_int, [C.POINTER(DltcudaAutoJob), _SZ, C.c_bool]),
        "dltcuda_transform_auto_batch_multi_gpu": (C.c_int, [C.POINTER(DltcudaAutoJob), _SZ, C.c_bool, C.POINTER(C.c_int), C.c_int]),
    }
)


_U32 = C.c_uint32
SIGNATURES.update(
    {
        "dltff_error_message": (C.c_char_p, [C.c_int32]),
        "dltff_TransformHeader_new": (_U32, [C.c_int32, _U32]),
        "dltff_TransformHeader_format": (C.c_bool, [_U32, C.POINTER(C.c_int32)]),
        "dltff_TransformHeader_format_data": (_U32, [_U32]),
        "dltff_TransformHeader_read": (_U32, [_P]),
        "dltff_TransformHeader_write": (None, [_U32, _P]),
        "dltff_bc1_header_from_settings": (_U32, [C.c_uint8, C.c_bool]),
        "dltff_bc2_header_from_settings": (_U32, [C.c_uint8, C.c_bool]),
        "dltff_bc1_settings_from_header": (DltffResult, [_U32, C.POINTER(C.c_uint8), C.POINTER(C.c_bool)]),
        "dltff_bc2_settings_from_header": (DltffResult, [_U32, C.POINTER(C.c_uint8), C.POINTER(C.c_bool)]),
        "dltff_new_TransformBundle": (_P, []),
        "dltff_TransformBundle_default_all": (_P, []),
        "dltff_free_TransformBundle": (None, [_P]),
        "dltff_TransformBundle_with_bc1_manual": (DltffResult, [_P, _P]),
        "dltff_TransformBundle_with_bc1_auto": (DltffResult, [_P, _P]),
        "dltff_TransformBundle_with_bc2_manual": (DltffResult, [_P, _P]),
        "dltff_TransformBundle_with_bc2_auto": (DltffResult, [_P, _P]),
        "dltff_dispatch_transform": (DltffResult, [C.c_int32, _P, _SZ, _P, _SZ, _P, C.POINTER(_U32)]),
        "dltff_dispatch_untransform": (DltffResult, [_U32, _P, _SZ, _P, _SZ]),
        "is_dds": (C.c_bool, [_P, _SZ]),
        "parse_dds": (DdsInfo, [_P, _SZ]),
        "dltdds_parse_dds_ignore_magic": (DdsInfo, [_P, _SZ]),
        "dltdds_can_handle": (C.c_bool, [_P, _SZ, C.c_char_p]),
        "dltdds_can_handle_untransform": (C.c_bool, [_P, _SZ, C.c_char_p]),
        "dltdds_transform_bundle": (DltffResult, [_P, _SZ, _P, _SZ, _P]),
        "dltdds_untransform": (DltffResult, [_P, _SZ, _P, _SZ]),
        "dltdds_transform_bundle_batch": (C.c_int, [C.POINTER(DltddsFile), _SZ, _P, C.POINTER(DltffResult), C.POINTER(C.c_int), C.c_int]),
        "dltdds_untransform_batch": (C.c_int, [C.POINTER(DltddsFile), _SZ, C.POINTER(DltffResult), C.POINTER(C.c_int), C.c_int]),
    }
)


def lib() -> C.CDLL:
    """The loaded shared library (loaded once).  Raises NativeLibraryMissing if it was not built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise NativeLibraryMissing(
                f"{LIB_PATH} is missing: build it with `python -m dxt_lossless_transform_b200.build` "
                "(there is no CPU fallback)"
            )
        handle = C.CDLL(str(LIB_PATH))
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the library lacks a declared symbol
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
    return _lib
