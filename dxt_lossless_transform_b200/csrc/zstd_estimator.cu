// zstd_estimator.cu — ZStandardSizeEstimation behind the DltSizeEstimator vtable.
//
// Reference: extensions/compressors/dxt-lossless-transform-zstd/src/lib.rs
//   :60-69   ZStandardSizeEstimation::new(level): 1..=22, anything else is InvalidLevel
//   :103-112 max_compressed_size(len) = 0 for len == 0, else ZSTD_compressBound(len)
//   :114-139 estimate_compressed_size: null or empty input -> 0, else the size ZSTD_compress2 produces
//   :150-189 compress(): fresh CCtx per call, freed afterwards
//   :193-209 parameters: compressionLevel, format = ZSTD_f_zstd1_magicless, contentSizeFlag = checksumFlag =
//            dictIDFlag = 0
// The reference links zstd statically through zstd-sys (2.0.16+zstd.1.5.7, src/Cargo.lock:1724-1727).  zstd is a
// third-party library and not part of the reference tree, so — exactly like the reference — this file CALLS it; it is
// bound at run time (dlopen of the system's libzstd.so.1), which keeps libdxt_lossless_transform_cuda.so free of a link
// dependency.  Without a loadable libzstd the factory returns NULL: there is no substitute estimator.  Compressed sizes
// equal the reference's whenever the library versions match (dltzstd_version_number() reports the one in use).
//
// What runs where: the estimator is a HOST callback by contract (api-common/src/c_api/size_estimation.rs:76-115).  In
// transform_bcN_auto the candidates are transformed on the GPU; the reference then compresses them one after another on
// the calling thread, here (cabi.cu, auto_host) every candidate is compressed by its own host thread as soon as its
// endpoint streams have arrived from the device, so the search costs one zstd pass of wall time instead of K.
#include <dlfcn.h>

#include <cstdint>
#include <cstdlib>
#include <mutex>
#include <new>

#include "cabi_internal.h"
#include "zstd_estimator.h"

#define DLT_EXPORT extern "C" __attribute__((visibility("default")))

namespace dlt {
namespace {

// zstd.h (stable since 1.4.0): ZSTD_c_compressionLevel = 100, ZSTD_c_contentSizeFlag = 200, ZSTD_c_checksumFlag = 201,
// ZSTD_c_dictIDFlag = 202, ZSTD_c_experimentalParam2 (= ZSTD_c_format) = 10, ZSTD_f_zstd1_magicless = 1.
constexpr int kParamLevel = 100, kParamFormat = 10, kParamContentSize = 200, kParamChecksum = 201, kParamDictId = 202;
constexpr int kFormatMagicless = 1;

struct ZstdApi {
    void* (*create_cctx)();
    size_t (*free_cctx)(void*);
    size_t (*set_parameter)(void*, int, int);
    size_t (*compress2)(void*, void*, size_t, const void*, size_t);
    size_t (*compress_bound)(size_t);
    unsigned (*is_error)(size_t);
    unsigned (*version_number)();
};

const ZstdApi* zstd_api() {
    static ZstdApi api{};
    static bool ok = false;
    static std::once_flag once;
    std::call_once(once, [] {
        void* h = nullptr;
        const char* override_path = std::getenv("DLTCUDA_LIBZSTD");
        if (override_path && *override_path) h = dlopen(override_path, RTLD_NOW | RTLD_LOCAL);
        if (!h) h = dlopen("libzstd.so.1", RTLD_NOW | RTLD_LOCAL);
        if (!h) h = dlopen("libzstd.so", RTLD_NOW | RTLD_LOCAL);
        if (!h) return;
        auto sym = [h](const char* name) { return dlsym(h, name); };
        api.create_cctx = reinterpret_cast<void* (*)()>(sym("ZSTD_createCCtx"));
        api.free_cctx = reinterpret_cast<size_t (*)(void*)>(sym("ZSTD_freeCCtx"));
        api.set_parameter = reinterpret_cast<size_t (*)(void*, int, int)>(sym("ZSTD_CCtx_setParameter"));
        api.compress2 = reinterpret_cast<size_t (*)(void*, void*, size_t, const void*, size_t)>(sym("ZSTD_compress2"));
        api.compress_bound = reinterpret_cast<size_t (*)(size_t)>(sym("ZSTD_compressBound"));
        api.is_error = reinterpret_cast<unsigned (*)(size_t)>(sym("ZSTD_isError"));
        api.version_number = reinterpret_cast<unsigned (*)()>(sym("ZSTD_versionNumber"));
        ok = api.create_cctx && api.free_cctx && api.set_parameter && api.compress2 && api.compress_bound &&
             api.is_error && api.version_number;
    });
    return ok ? &api : nullptr;
}

struct ZstdContext {
    int level;
};

uint32_t zstd_max_compressed_size(void* context, size_t len, size_t* out_size) {
    const ZstdApi* z = zstd_api();
    if (!context || !out_size || !z) return 1;
    *out_size = len == 0 ? 0 : z->compress_bound(len);   // lib.rs:103-112
    return 0;
}

uint32_t zstd_estimate_compressed_size(void* context, const uint8_t* input, size_t len, uint8_t* output,
                                       size_t output_len, size_t* out_size) {
    if (!context || !out_size) return 1;
    if (!input || len == 0) {   // lib.rs:121-127
        *out_size = 0;
        return 0;
    }
    if (!output) return 1;
    return zstd_compressed_size(static_cast<const ZstdContext*>(context)->level, input, len, output, output_len, out_size);
}

}  // namespace

uint32_t zstd_compressed_size(int level, const uint8_t* input, size_t len, uint8_t* output, size_t output_len,
                              size_t* out_size) {
    const ZstdApi* z = zstd_api();
    if (!z) return 3;
    void* cctx = z->create_cctx();   // lib.rs:152-157
    if (!cctx) return 2;
    z->set_parameter(cctx, kParamLevel, level);
    z->set_parameter(cctx, kParamFormat, kFormatMagicless);
    z->set_parameter(cctx, kParamContentSize, 0);
    z->set_parameter(cctx, kParamChecksum, 0);
    z->set_parameter(cctx, kParamDictId, 0);
    const size_t r = z->compress2(cctx, output, output_len, input, len);
    z->free_cctx(cctx);
    if (z->is_error(r)) return 3;    // lib.rs:178-186: ZStandardInternal
    *out_size = r;
    return 0;
}

bool is_zstd_estimator(const DltSizeEstimator& e) { return e.estimate_compressed_size == &zstd_estimate_compressed_size; }
int zstd_estimator_level(const DltSizeEstimator& e) { return static_cast<const ZstdContext*>(e.context)->level; }

}  // namespace dlt

using namespace dlt;

// ZStandardSizeEstimation::new (lib.rs:60-69): NULL for a level outside 1..=22, when libzstd cannot be loaded, or on
// allocation failure.  Additive: the reference crate has no C exports (its Rust type is what the CLI presets use).
DLT_EXPORT DltSizeEstimator* dltzstd_new_size_estimator(int compression_level) {
    if (compression_level < 1 || compression_level > 22) return nullptr;
    if (!zstd_api()) return nullptr;
    ZstdContext* ctx = new (std::nothrow) ZstdContext{compression_level};
    if (!ctx) return nullptr;
    DltSizeEstimator* e = new (std::nothrow) DltSizeEstimator{ctx, &zstd_max_compressed_size, &zstd_estimate_compressed_size};
    if (!e) delete ctx;
    return e;
}

DLT_EXPORT void dltzstd_free_size_estimator(DltSizeEstimator* estimator) {
    if (!estimator) return;
    if (is_zstd_estimator(*estimator)) delete static_cast<ZstdContext*>(estimator->context);
    delete estimator;
}

// ZSTD_versionNumber() of the library in use (e.g. 10505 = 1.5.5), 0 when none could be loaded.
DLT_EXPORT unsigned dltzstd_version_number(void) {
    const ZstdApi* z = zstd_api();
    return z ? z->version_number() : 0u;
}
