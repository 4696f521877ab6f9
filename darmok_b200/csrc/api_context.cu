// api_context.cu -- C ABI of include/dmk.h: context, error text, stream, kernel timing
#include "api_internal.hpp"

thread_local std::string g_create_error;

extern "C" {


const char* dmk_version(void) { return "darmok_b200 0.1 (sm_100a)"; }

dmk_status dmk_context_create(int device, void* cuda_stream, dmk_context_t* out) {
    if (!out) return fail(nullptr, DMK_ERR_INVALID, "out is null");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, DMK_ERR_CUDA, std::string("no CUDA device (this library has no CPU fallback): ") +
                                               (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
    if (device < 0 || device >= count) return fail(nullptr, DMK_ERR_INVALID, "device index out of range");
    if ((e = cudaSetDevice(device)) != cudaSuccess) return cuda_fail(nullptr, e, "cudaSetDevice");
    auto ctx = std::make_unique<dmk_context_s>();
    ctx->device = device;
    cudaDeviceProp prop{};
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return cuda_fail(nullptr, e, "cudaGetDeviceProperties");
    if (prop.major < 10)
        return fail(nullptr, DMK_ERR_CUDA, std::string("device is not sm_100a class (found sm_") + std::to_string(prop.major) +
                                               std::to_string(prop.minor) + "); kernels are built for B200 only");
    ctx->num_sms = prop.multiProcessorCount;
    if (cuda_stream) {
        ctx->stream = static_cast<cudaStream_t>(cuda_stream);
    } else {
        if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess)
            return cuda_fail(nullptr, e, "cudaStreamCreate");
        ctx->owns_stream = true;
    }
    *out = ctx.release();
    return DMK_OK;
}

void dmk_context_destroy(dmk_context_t ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    dev_free(ctx->cull_scratch);
    drain_timers(ctx);
    for (cudaEvent_t e : ctx->event_pool) cudaEventDestroy(e);
    if (ctx->owns_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* dmk_last_error(dmk_context_t ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }
void* dmk_context_stream(dmk_context_t ctx) { return ctx ? ctx->stream : nullptr; }
uint64_t dmk_context_launch_count(dmk_context_t ctx) { return ctx ? ctx->launches : 0; }

dmk_status dmk_context_enable_kernel_timing(dmk_context_t ctx, int on) {
    if (!ctx) return DMK_ERR_INVALID;
    ctx->timing = on != 0;
    return DMK_OK;
}
dmk_status dmk_context_kernel_time(dmk_context_t ctx, int kernel, double* total_ms, uint64_t* launches) {
    if (!ctx || kernel < 0 || kernel >= DMK_KERNEL_COUNT) return DMK_ERR_INVALID;
    if (bind(ctx)) return DMK_ERR_CUDA;
    drain_timers(ctx);
    if (total_ms) *total_ms = ctx->kernel_ms[kernel];
    if (launches) *launches = ctx->kernel_launches[kernel];
    return DMK_OK;
}
dmk_status dmk_context_reset_kernel_timing(dmk_context_t ctx) {
    if (!ctx) return DMK_ERR_INVALID;
    if (bind(ctx)) return DMK_ERR_CUDA;
    drain_timers(ctx);
    for (int k = 0; k < DMK_KERNEL_COUNT; ++k) {
        ctx->kernel_ms[k] = 0;
        ctx->kernel_launches[k] = 0;
    }
    return DMK_OK;
}

dmk_status dmk_sync(dmk_context_t ctx) {
    if (!ctx) return DMK_ERR_INVALID;
    if (bind(ctx)) return DMK_ERR_CUDA;
    DMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return DMK_OK;
}

}  // extern "C"
