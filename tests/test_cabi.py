"""CPU tests of the C ABI: the library loads, exports every symbol include/*.h declares, and the
argument validation (which runs before any CUDA call) returns the reference's error codes."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

import dxt_lossless_transform_b200 as dlt
from dxt_lossless_transform_b200 import _native as N
from dxt_lossless_transform_b200 import api, sharding

ROOT = Path(__file__).resolve().parent.parent
INCLUDE = ROOT / "include"


def declared_functions() -> set[str]:
    names = set()
    for h in INCLUDE.glob("*.h"):
        text = re.sub(r"/\*.*?\*/", "", h.read_text(), flags=re.S)
        names |= set(re.findall(r"\b(dlt[a-z0-9]*_\w+|is_dds|parse_dds)\s*\(", text))
    return names


def exported_functions() -> set[str]:
    out = subprocess.run(["nm", "-D", "--defined-only", str(N.LIB_PATH)], capture_output=True, text=True, check=True)
    return {l.split()[-1] for l in out.stdout.splitlines() if " T " in l}


def test_library_exports_exactly_what_the_headers_declare():
    declared, exported = declared_functions(), exported_functions()
    assert declared - exported == set(), f"declared but not exported: {declared - exported}"
    assert {e for e in exported if e.startswith("dlt") or e in ("is_dds", "parse_dds")} - declared == set()
    assert set(N.SIGNATURES) == declared
    N.lib()  # resolves every symbol through ctypes


@pytest.mark.parametrize("lang", ["c", "c++"])
def test_headers_compile_together(lang, tmp_path):
    src = tmp_path / ("t.c" if lang == "c" else "t.cpp")
    src.write_text("".join(f'#include "{h.name}"\n' for h in sorted(INCLUDE.glob("*.h"))) +
                   "int main(void){ Dltbc1CoreTransformSettings s = {true, DltCoreYCoCgVariant_Variant1};"
                   " return sizeof(s) == 2 && sizeof(Dltbc1Result) == 4 && sizeof(DltSizeEstimator) == 24 ? 0 : 1; }\n")
    exe = tmp_path / "t"
    subprocess.run(["gcc" if lang == "c" else "g++", "-Wall", "-Werror", f"-I{INCLUDE}", str(src), "-o", str(exe)], check=True)
    assert subprocess.run([str(exe)]).returncode == 0


def test_struct_layouts_match_ctypes():
    assert C.sizeof(N.CoreSettings) == 2 and C.sizeof(N.CoreBc3Settings) == 3
    assert C.sizeof(N.DltcudaSettings) == 4 and C.sizeof(N.DltResult) == 4
    assert N.CoreSettings.split_colour_endpoints.offset == 0 and N.CoreSettings.decorrelation_mode.offset == 1


# ---- core ABI: check order of core/.../c_api/transform_with_settings.rs:73-100 -------------------
@pytest.mark.parametrize("n,bpb", [(1, 8), (2, 16), (3, 16)])
def test_core_validation_codes(n, bpb):
    L = N.lib()
    S = N.CoreBc3Settings(True, True, 1) if n == 3 else N.CoreSettings(True, 1)
    buf = np.zeros(64, np.uint8)
    for name in ("transform", "untransform"):
        f = getattr(L, f"dltbc{n}core_{name}")
        assert f(None, 16, buf.ctypes.data, 16, S).error_code == api.CORE_NULL_DATA
        assert f(buf.ctypes.data, 16, None, 16, S).error_code == api.CORE_NULL_OUTPUT
        assert f(None, 16, None, 16, S).error_code == api.CORE_NULL_DATA           # input checked first
        assert f(buf.ctypes.data, bpb - 1, buf.ctypes.data, 64, S).error_code == api.CORE_INVALID_LENGTH
        assert f(buf.ctypes.data, bpb + 1, buf.ctypes.data, 0, S).error_code == api.CORE_INVALID_LENGTH  # length before size
        assert f(buf.ctypes.data, 2 * bpb, buf.ctypes.data, 2 * bpb - 1, S).error_code == api.CORE_OUTPUT_TOO_SMALL
        assert f(buf.ctypes.data, 0, buf.ctypes.data, 0, S).error_code == api.CORE_SUCCESS  # len 0 is a no-op


@pytest.mark.parametrize("n", [1, 2, 3])
def test_core_auto_null_checks(n):
    L = N.lib()
    f = getattr(L, f"dltbc{n}core_transform_auto")
    buf = np.zeros(64, np.uint8)
    est = L.dltltu_new_size_estimator()
    details = N.CoreBc3Settings() if n == 3 else N.CoreSettings()
    a = N.CoreAutoSettings(False)
    assert f(None, 16, buf.ctypes.data, 16, est, a, C.byref(details)).error_code == api.CORE_NULL_DATA
    assert f(buf.ctypes.data, 16, None, 16, est, a, C.byref(details)).error_code == api.CORE_NULL_OUTPUT
    assert f(buf.ctypes.data, 16, buf.ctypes.data, 16, None, a, C.byref(details)).error_code == api.CORE_NULL_ESTIMATOR
    assert f(buf.ctypes.data, 16, buf.ctypes.data, 16, est, a, None).error_code == api.CORE_NULL_SETTINGS
    assert f(buf.ctypes.data, 17, buf.ctypes.data, 32, est, a, C.byref(details)).error_code == api.CORE_INVALID_LENGTH
    assert f(buf.ctypes.data, 32, buf.ctypes.data, 16, est, a, C.byref(details)).error_code == api.CORE_OUTPUT_TOO_SMALL
    L.dltltu_free_size_estimator(est)
    L.dltltu_free_size_estimator(None)


# ---- stable ABI (api crates) -------------------------------------------------------------------------
@pytest.mark.parametrize("n,bpb", [(1, 8), (2, 16)])
def test_stable_manual_builder_codes_and_lifecycle(n, bpb):
    L = N.lib()
    g = lambda name: getattr(L, f"dltbc{n}_{name}")
    b = g("new_ManualTransformBuilder")()
    assert b
    buf = np.zeros(64, np.uint8)
    p = buf.ctypes.data
    for name in ("ManualTransformBuilder_Transform", "ManualTransformBuilder_Untransform"):
        f = g(name)
        assert f(None, 16, p, 16, b).error_code == 5     # NullDataPointer
        assert f(p, 16, None, 16, b).error_code == 9     # NullOutputBufferPointer
        assert f(p, 16, p, 16, None).error_code == 10    # NullManualTransformBuilderPointer
        assert f(p, bpb + 3, p, 64, b).error_code == 1   # InvalidLength
        assert f(p, 2 * bpb, p, bpb, b).error_code == 2  # OutputBufferTooSmall
        assert f(p, 0, p, 0, b).error_code == 0
    # defaults (Variant1 = stable 0, split) / setters / reset / clone
    mode, split = C.c_uint8(9), C.c_bool(False)
    L.dltcuda_ManualTransformBuilder_GetSettings(b, C.byref(mode), C.byref(split))
    assert (mode.value, split.value) == (0, True)
    g("ManualTransformBuilder_SetDecorrelationMode")(b, 3)  # None
    g("ManualTransformBuilder_SetSplitColourEndpoints")(b, False)
    c = g("clone_ManualTransformBuilder")(b)
    L.dltcuda_ManualTransformBuilder_GetSettings(c, C.byref(mode), C.byref(split))
    assert (mode.value, split.value) == (3, False)
    g("ManualTransformBuilder_ResetToDefaults")(b)
    L.dltcuda_ManualTransformBuilder_GetSettings(b, C.byref(mode), C.byref(split))
    assert (mode.value, split.value) == (0, True)
    # null-safety of the setters and lifecycle functions (manual_transform_builder.rs:86,107,150,183,203)
    g("ManualTransformBuilder_SetDecorrelationMode")(None, 1)
    g("ManualTransformBuilder_SetSplitColourEndpoints")(None, True)
    g("ManualTransformBuilder_ResetToDefaults")(None)
    assert not g("clone_ManualTransformBuilder")(None)
    g("free_ManualTransformBuilder")(None)
    g("free_ManualTransformBuilder")(b)
    g("free_ManualTransformBuilder")(c)


@pytest.mark.parametrize("n", [1, 2])
def test_stable_auto_builder_codes(n):
    L = N.lib()
    g = lambda name: getattr(L, f"dltbc{n}_{name}")
    assert not g("new_AutoTransformBuilder")(None)
    est = L.dltltu_new_size_estimator()
    b = g("new_AutoTransformBuilder")(est)
    L.dltltu_free_size_estimator(est)  # the builder holds a copy
    assert g("AutoTransformBuilder_SetUseAllDecorrelationModes")(None, True).error_code == 11
    assert g("AutoTransformBuilder_SetUseAllDecorrelationModes")(b, True).error_code == 0
    buf = np.zeros(64, np.uint8)
    p = buf.ctypes.data
    out = C.c_void_p(1234)
    f = g("AutoTransformBuilder_Transform")
    assert f(None, p, 16, p, 16, C.byref(out)).error_code == 11
    assert f(b, None, 16, p, 16, C.byref(out)).error_code == 5
    assert f(b, p, 16, None, 16, C.byref(out)).error_code == 9
    assert f(b, p, 16, p, 16, None).error_code == 12
    assert f(b, p, 17, p, 32, C.byref(out)).error_code == 1 and out.value is None  # *out = NULL on error
    out = C.c_void_p(1234)
    assert f(b, p, 32, p, 16, C.byref(out)).error_code == 2 and out.value is None
    g("free_AutoTransformBuilder")(b)
    g("free_AutoTransformBuilder")(None)


def test_error_messages():
    L = N.lib()
    assert L.dltbc1_error_message(0) == b"Success"
    assert L.dltbc1_error_message(1) == b"Invalid input length: Length must be divisible by 8 (BC1 block size)"
    assert L.dltbc2_error_message(1) == b"Invalid input length: Length must be divisible by 16 (BC2 block size)"
    assert L.dltbc1_error_message(10) == b"Null pointer provided for Dltbc1ManualTransformBuilder parameter"
    assert L.dltbc2_error_message(10) == b"Null pointer provided for Dltbc2ManualTransformBuilder parameter"
    for code in range(13):
        assert L.dltbc1_error_message(code) and L.dltbc2_error_message(code)


def test_ltu_estimator_vtable_edge_cases():
    # extensions/estimators/dxt-lossless-transform-ltu/src/c_api.rs:106-150, lib.rs:96-119
    L = N.lib()
    e = L.dltltu_new_size_estimator()
    out = C.c_size_t(77)
    assert e.contents.max_compressed_size(e.contents.context, 1000, C.byref(out)) == 0 and out.value == 0
    assert e.contents.max_compressed_size(None, 1000, C.byref(out)) == 1
    out = C.c_size_t(77)
    assert e.contents.estimate_compressed_size(e.contents.context, None, 0, None, 0, C.byref(out)) == 0 and out.value == 0
    assert e.contents.estimate_compressed_size(None, None, 0, None, 0, C.byref(out)) == 1
    L.dltltu_free_size_estimator(e)


# ---- host-side logic --------------------------------------------------------------------------------
def test_python_mirror_validation_errors():
    with pytest.raises(api.InvalidLength):
        dlt.transform_bc1_with_settings(np.zeros(7, np.uint8), np.zeros(7, np.uint8))
    with pytest.raises(api.OutputBufferTooSmall):
        dlt.transform_bc3_with_settings(np.zeros(32, np.uint8), np.zeros(16, np.uint8))
    with pytest.raises(api.BcnError) as e:
        dlt.Bc1ManualTransformBuilder().transform(np.zeros(9, np.uint8), np.zeros(16, np.uint8))
    assert e.value.code == 1
    assert dlt.Bc1TransformSettings() == dlt.Bc1TransformSettings(dlt.YCoCgVariant.Variant1, True)
    assert dlt.Bc3TransformSettings() == dlt.Bc3TransformSettings(dlt.YCoCgVariant.Variant1, True, True)
    assert len(list(dlt.Bc1TransformSettings.all_combinations())) == 8
    assert len(list(dlt.Bc3TransformSettings.all_combinations())) == 16
    assert [dlt.YCoCgVariant.from_stable(v.to_stable()) for v in dlt.YCoCgVariant] == list(dlt.YCoCgVariant)


def test_candidate_orders_match_reference_tables():
    V = dlt.YCoCgVariant
    fast = [(V.NONE, False), (V.NONE, True), (V.Variant1, False), (V.Variant1, True)]
    full = [(V.Variant2, False), (V.NONE, False), (V.NONE, True), (V.Variant3, False), (V.Variant3, True),
            (V.Variant2, True), (V.Variant1, False), (V.Variant1, True)]
    for fmt in (1, 2):
        assert [(s.decorrelation_mode, s.split_colour_endpoints) for s in api.auto_candidates(fmt, False)] == fast
        assert [(s.decorrelation_mode, s.split_colour_endpoints) for s in api.auto_candidates(fmt, True)] == full
    fast3 = [(V.Variant1, True, False), (V.Variant1, True, True), (V.NONE, True, False), (V.NONE, False, True),
             (V.NONE, True, True), (V.Variant1, False, True), (V.NONE, False, False), (V.Variant1, False, False)]
    got = [(s.decorrelation_mode, s.split_alpha_endpoints, s.split_colour_endpoints) for s in api.auto_candidates(3, False)]
    assert got == fast3
    all3 = api.auto_candidates(3, True)
    assert len(all3) == 16 and len(set(all3)) == 16
    assert all3[0] == dlt.Bc3TransformSettings(V.Variant2, True, False)
    assert all3[-1] == dlt.Bc3TransformSettings(V.Variant1, False, False)


@pytest.mark.parametrize("fmt", [1, 2, 3])
def test_shard_ranges_partition_the_payload_on_tile_boundaries(fmt):
    tile = 2048 if fmt == 1 else 1024
    for total in (0, 1, tile - 1, tile, 10 * tile + 17, 1 << 20, 134217728):
        for shards in (1, 2, 3, 4, 8):
            r = sharding.shard_ranges(fmt, total, shards)
            assert sum(c for _, c in r) == total
            pos = 0
            for first, count in r:
                assert first == pos and (first % tile == 0 or first == total)
                pos += count
    s = dlt.Bc3TransformSettings() if fmt == 3 else dlt.Bc1TransformSettings()
    n = 10 * tile + 17
    slices = sharding.stream_slices(fmt, s, n, 0, n)
    assert slices[0][0] == 0 and sum(l for _, l in slices) == n * (8 if fmt == 1 else 16)
    for (o0, l0), (o1, _) in zip(slices, slices[1:]):
        assert o0 + l0 == o1


def test_no_gpu_means_a_loud_error_not_a_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("this box has a GPU")
    data = np.arange(64, dtype=np.uint8)
    with pytest.raises(api.TransformationError):
        dlt.transform_bc1_with_settings(data, np.zeros_like(data))
    with pytest.raises(api.BcnError) as e:
        dlt.Bc2ManualTransformBuilder().transform(data, np.zeros_like(data))
    assert e.value.code == 3
    assert api.device_count() == 0


def test_cpp_host_header_links_and_validates(tmp_path):
    """include/dxt_lossless_transform.hpp (the compiled host layer above the C ABI) builds against the
    shared library and reports the reference's validation errors (no GPU needed for those)."""
    src = tmp_path / "host.cpp"
    src.write_text(r'''
#include "dxt_lossless_transform.hpp"
#include <cstdio>
#include <vector>
namespace dlt = dxt_lossless_transform;
int main() {
    std::vector<uint8_t> in(64), out(64);
    int ok = 0;
    try { dlt::transform_bc1_with_settings(in.data(), 7, out.data(), 64); } catch (const dlt::InvalidLength& e) { ok += e.length == 7; }
    try { dlt::untransform_bc3_with_settings(in.data(), 32, out.data(), 16); } catch (const dlt::OutputBufferTooSmall& e) { ok += e.needed == 32 && e.actual == 16; }
    try { dlt::transform_bc2_auto(in.data(), 17, out.data(), 64, nullptr); } catch (const dlt::DeviceError& e) { ok += e.core_code == 3; }
    dlt::transform_bc2_with_settings(in.data(), 0, out.data(), 0);  // len 0 is a no-op
    dlt::Bc3TransformSettings d;
    ok += d.decorrelation_mode == dlt::YCoCgVariant::Variant1 && d.split_alpha_endpoints && d.split_colour_endpoints;
    dlt::LosslessTransformUtilsSizeEstimation ltu;
    ok += ltu.estimate_compressed_size(nullptr, 0) == 0;
    try { dlt::ZStandardSizeEstimation z(23); } catch (const dlt::InvalidLevel&) { ok += 1; }
    try { dlt::experimental::transform_bc1_with_normalize_blocks(in.data(), 12, out.data(), 64); } catch (const dlt::InvalidLength& e) { ok += e.length == 12; }
    try { dlt::experimental::transform_bc1_auto_with_normalization(in.data(), 64, out.data(), 8); } catch (const dlt::OutputBufferTooSmall& e) { ok += e.needed == 64; }
    dlt::experimental::Bc1TransformDetailsWithNormalization nd;
    ok += nd.color_normalization_mode == dlt::experimental::ColorNormalizationMode::None && nd.split_colour_endpoints;
    std::printf("%d\n", ok);
    return ok == 9 ? 0 : 1;
}
''')
    exe = tmp_path / "host"
    libdir = N.LIB_PATH.parent
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", f"-I{INCLUDE}", str(src), "-o", str(exe), f"-L{libdir}",
                    "-ldxt_lossless_transform_cuda", f"-Wl,-rpath,{libdir}"], check=True)
    assert subprocess.run([str(exe)]).returncode == 0


# ---- C code written against the REFERENCE's generated headers (cbindgen's spelling) --------------------------
CBINDGEN_USER_API = r'''
/* What a user of the reference's generated `dxt-lossless-transform-bc%(n)d-api` header writes: bare enumerators,
 * PascalCase fields (.github/cbindgen_c.toml), the documented builder flow (c_api/mod.rs). */
#include "dxt_lossless_transform_bc%(n)d_api.h"
#include <stdio.h>
#include <string.h>
int main(int argc, char **argv) {
    enum { BLOCK = %(bpb)d, N = 1024 };
    static uint8_t in[BLOCK * N], mid[BLOCK * N], back[BLOCK * N];
    for (size_t i = 0; i < sizeof in; i++) in[i] = (uint8_t)(i * 131u + (i >> 7));
    Dltbc%(n)dManualTransformBuilder *b = dltbc%(n)d_new_ManualTransformBuilder();
    if (!b) return 10;
    dltbc%(n)d_ManualTransformBuilder_SetDecorrelationMode(b, Variant2);
    dltbc%(n)d_ManualTransformBuilder_SetSplitColourEndpoints(b, false);
    Dltbc%(n)dTransformSettings s = { .DecorrelationMode = None, .SplitColourEndpoints = true };
    (void)s;
    Dltbc%(n)dResult r = dltbc%(n)d_ManualTransformBuilder_Transform(NULL, sizeof in, mid, sizeof mid, b);
    if (r.ErrorCode != NullDataPointer) return 11;
    r = dltbc%(n)d_ManualTransformBuilder_Transform(in, sizeof in, NULL, sizeof mid, b);
    if (r.ErrorCode != NullOutputBufferPointer) return 12;
    r = dltbc%(n)d_ManualTransformBuilder_Transform(in, sizeof in, mid, sizeof mid, NULL);
    if (r.ErrorCode != NullManualTransformBuilderPointer) return 13;
    r = dltbc%(n)d_ManualTransformBuilder_Transform(in, BLOCK + 1, mid, sizeof mid, b);
    if (r.ErrorCode != InvalidLength) return 14;
    r = dltbc%(n)d_ManualTransformBuilder_Transform(in, sizeof in, mid, BLOCK, b);
    if (r.ErrorCode != OutputBufferTooSmall) return 15;
    if (strcmp(dltbc%(n)d_error_message(Success), dltbc%(n)d_error_message(InvalidLength)) == 0) return 16;
    DltSizeEstimator *est = dltltu_new_size_estimator();
    if (!est || !est->MaxCompressedSize || !est->EstimateCompressedSize) return 17;
    Dltbc%(n)dAutoTransformBuilder *ab = dltbc%(n)d_new_AutoTransformBuilder(est);
    if (!ab) return 18;
    if (dltbc%(n)d_AutoTransformBuilder_SetUseAllDecorrelationModes(NULL, true).ErrorCode != NullBuilderPointer) return 19;
    Dltbc%(n)dManualTransformBuilder *won = (Dltbc%(n)dManualTransformBuilder *)1;
    r = dltbc%(n)d_AutoTransformBuilder_Transform(ab, NULL, sizeof in, mid, sizeof mid, &won);
    if (r.ErrorCode != NullDataPointer || won != (Dltbc%(n)dManualTransformBuilder *)1) return 20; /* pointer checks return before *out is written */
    r = dltbc%(n)d_AutoTransformBuilder_Transform(ab, in, BLOCK + 1, mid, sizeof mid, &won);
    if (r.ErrorCode != InvalidLength || won != NULL) return 21;                                    /* a failed search nulls it */
    if (argc > 1) { /* a GPU is present: the full documented flow */
        r = dltbc%(n)d_ManualTransformBuilder_Transform(in, sizeof in, mid, sizeof mid, b);
        if (r.ErrorCode != Success) return 30;
        r = dltbc%(n)d_ManualTransformBuilder_Untransform(mid, sizeof mid, back, sizeof back, b);
        if (r.ErrorCode != Success || memcmp(in, back, sizeof in)) return 31;
        r = dltbc%(n)d_AutoTransformBuilder_Transform(ab, in, sizeof in, mid, sizeof mid, &won);
        if (r.ErrorCode != Success || !won) return 32;
        r = dltbc%(n)d_ManualTransformBuilder_Untransform(mid, sizeof mid, back, sizeof back, won);
        if (r.ErrorCode != Success || memcmp(in, back, sizeof in)) return 33;
        dltbc%(n)d_free_ManualTransformBuilder(won);
    }
    dltbc%(n)d_free_AutoTransformBuilder(ab);
    dltltu_free_size_estimator(est);
    dltbc%(n)d_free_ManualTransformBuilder(b);
    return 0;
}
'''
CBINDGEN_USER_CORE = r'''
/* ... and of the core crate's generated header (`dxt-lossless-transform-bc%(n)d`, feature c-exports): same type names,
 * different layout and values (SURVEY section 8b). */
#include "dxt_lossless_transform_bc%(n)d.h"
#include <string.h>
int main(int argc, char **argv) {
    enum { BLOCK = %(bpb)d, N = 777 };
    static uint8_t in[BLOCK * N], mid[BLOCK * N], back[BLOCK * N];
    for (size_t i = 0; i < sizeof in; i++) in[i] = (uint8_t)(i * 29u + (i >> 5));
    Dltbc%(n)dTransformSettings s = { .SplitColourEndpoints = true, .DecorrelationMode = Variant3 };
    Dltbc%(n)dUntransformSettings u = s;
    if (sizeof s != 2 || None != 0 || Variant3 != 3) return 10;
    if (dltbc%(n)dcore_transform(NULL, sizeof in, mid, sizeof mid, s).ErrorCode != NullDataPointer) return 11;
    if (dltbc%(n)dcore_transform(in, sizeof in, NULL, sizeof mid, s).ErrorCode != NullOutputBufferPointer) return 12;
    if (dltbc%(n)dcore_transform(in, BLOCK - 1, mid, sizeof mid, s).ErrorCode != InvalidDataLength) return 13;
    if (dltbc%(n)dcore_untransform(in, sizeof in, mid, BLOCK, u).ErrorCode != OutputBufferTooSmall) return 14;
    Dltbc%(n)dAutoTransformSettings a = { .UseAllModes = true };
    Dltbc%(n)dTransformSettings best;
    if (dltbc%(n)dcore_transform_auto(in, sizeof in, mid, sizeof mid, NULL, a, &best).ErrorCode != NullEstimatorPointer) return 15;
    if (argc > 1) {
        if (dltbc%(n)dcore_transform(in, sizeof in, mid, sizeof mid, s).ErrorCode != Success) return 30;
        if (dltbc%(n)dcore_untransform(mid, sizeof mid, back, sizeof back, u).ErrorCode != Success || memcmp(in, back, sizeof in)) return 31;
        DltSizeEstimator *est = dltltu_new_size_estimator();
        if (dltbc%(n)dcore_transform_auto(in, sizeof in, mid, sizeof mid, est, a, &best).ErrorCode != Success) return 32;
        if (dltbc%(n)dcore_untransform(mid, sizeof mid, back, sizeof back, best).ErrorCode != Success || memcmp(in, back, sizeof in)) return 33;
        dltltu_free_size_estimator(est);
    }
    return 0;
}
'''


@pytest.mark.parametrize("which", ["bc1_api", "bc2_api", "bc1", "bc2"])
def test_code_written_against_cbindgen_headers_compiles_and_links(which, tmp_path):
    """INTEGRATION.md section 1 claims that C code written against the reference's cbindgen headers links unchanged: the
    generated compatibility headers (include/cbindgen_compat, tools/gen_cbindgen_compat.py) are current, a program in
    cbindgen's spelling compiles with -Wall -Werror, links to the product library, and sees the reference's error codes
    (and, on a GPU box, round-trips through the documented builder flow)."""
    import torch

    assert subprocess.run(["python", str(ROOT / "tools" / "gen_cbindgen_compat.py"), "--check"]).returncode == 0
    n = 1 if "bc1" in which else 2
    text = (CBINDGEN_USER_API if which.endswith("_api") else CBINDGEN_USER_CORE) % {"n": n, "bpb": 8 if n == 1 else 16}
    src, exe = tmp_path / "user.c", tmp_path / "user"
    src.write_text(text)
    libdir = N.LIB_PATH.parent
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", "-Wno-unused-parameter", f"-I{INCLUDE / 'cbindgen_compat'}", str(src), "-o",
                    str(exe), f"-L{libdir}", "-ldxt_lossless_transform_cuda", f"-Wl,-rpath,{libdir}"], check=True)
    args = [str(exe)] + (["gpu"] if torch.cuda.is_available() else [])
    assert subprocess.run(args).returncode == 0


def test_the_rust_crate_builds_from_the_same_lists_as_build_py():
    """rust/dxt-lossless-transform-cuda/build.rs used to carry its own (stale) copy of the source list: both builds
    now read csrc/SOURCES.txt and csrc/NVCC_FLAGS.txt, and the list is every .cu file of csrc/."""
    from dxt_lossless_transform_b200 import build

    rs = (ROOT / "rust" / "dxt-lossless-transform-cuda" / "build.rs").read_text()
    assert "SOURCES.txt" in rs and "NVCC_FLAGS.txt" in rs and ".cu\"" not in rs
    assert "rustc-link-lib=dylib=dl" in rs
    assert sorted(build.SOURCES) == sorted(p.name for p in build.CSRC.glob("*.cu"))
    assert "-ldl" in build.NVCC_FLAGS and "arch=compute_100a,code=sm_100a" in build.NVCC_FLAGS
