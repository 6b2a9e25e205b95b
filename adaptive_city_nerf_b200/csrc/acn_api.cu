// Context, version and error plumbing of the C ABI (include/acn_b200.h).
#include "acn_common.cuh"
#include <string.h>

static thread_local char g_err[512] = "";

void acn_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

unsigned int* acn_scratch_word(acn_ctx* ctx) {
    if (!ctx || !ctx->scratch) return nullptr;
    unsigned int i = __atomic_fetch_add(&ctx->scratch_next, 1u, __ATOMIC_RELAXED);
    return ctx->scratch + (i % ACN_SCRATCH_WORDS);
}

extern "C" int acn_version(void) { return ACN_VERSION; }

extern "C" const char* acn_last_error(void) { return g_err; }

extern "C" int acn_create(int device, acn_ctx** out) {
    ACN_REQUIRE(out != nullptr, ACN_EINVAL, "acn_create: null out pointer");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    ACN_REQUIRE(e == cudaSuccess && n > 0, ACN_ENODEVICE, "acn_create: no CUDA device (%s)",
                e == cudaSuccess ? "count = 0" : cudaGetErrorString(e));
    ACN_REQUIRE(device >= 0 && device < n, ACN_EINVAL, "acn_create: device %d out of range [0,%d)", device, n);
    cudaDeviceProp p;
    ACN_CUDA(cudaGetDeviceProperties(&p, device));
    ACN_REQUIRE(p.major == 10, ACN_ENODEVICE,
                "acn_create: device %d is sm_%d%d; this library is built for sm_100a only", device, p.major, p.minor);
    acn_ctx* c = new acn_ctx();
    c->device = device;
    c->sm_count = p.multiProcessorCount;
    c->cc_major = p.major;
    c->cc_minor = p.minor;
    c->l2_bytes = p.l2CacheSize;
    c->max_smem_optin = (int)p.sharedMemPerBlockOptin;
    c->scratch = nullptr;
    c->scratch_next = 0;
    {
        int prev = -1;
        cudaGetDevice(&prev);
        cudaError_t es = cudaSetDevice(device);
        if (es == cudaSuccess) es = cudaMalloc((void**)&c->scratch, ACN_SCRATCH_WORDS * sizeof(unsigned int));
        if (prev >= 0 && prev != device) cudaSetDevice(prev);
        if (es != cudaSuccess) {
            delete c;
            acn_set_error("acn_create: scratch allocation failed: %s", cudaGetErrorString(es));
            return ACN_ECUDA;
        }
    }
    *out = c;
    return ACN_OK;
}

extern "C" int acn_destroy(acn_ctx* ctx) {
    ACN_REQUIRE(ctx != nullptr, ACN_EINVAL, "acn_destroy: null context");
    if (ctx->scratch) cudaFree(ctx->scratch);
    delete ctx;
    return ACN_OK;
}

extern "C" int acn_device_info(acn_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor, int64_t* l2_bytes) {
    ACN_REQUIRE(ctx != nullptr, ACN_EINVAL, "acn_device_info: null context");
    if (sm_count) *sm_count = ctx->sm_count;
    if (cc_major) *cc_major = ctx->cc_major;
    if (cc_minor) *cc_minor = ctx->cc_minor;
    if (l2_bytes) *l2_bytes = ctx->l2_bytes;
    return ACN_OK;
}
