// api_exchange.cu -- C ABI of include/dmk.h: visible-palette exchange over peer memory, stand-alone cull
#include "api_internal.hpp"

extern "C" {

// ---- visible-palette exchange over peer memory ---------------------------------------------------------------
static dmk_status exchange_check_shape(dmk_context_t ctx, int world, int rank, int64_t capacity, int palette_size) {
    if (world < 1 || world > kExchangeMaxRanks) return fail(ctx, DMK_ERR_INVALID, "world must be in [1, DMK_EXCHANGE_MAX_RANKS]");
    if (rank < 0 || rank >= world) return fail(ctx, DMK_ERR_INVALID, "rank out of range");
    if (capacity < 0 || capacity > 0x7fffffff) return fail(ctx, DMK_ERR_INVALID, "capacity out of range");
    if (palette_size < 1 || palette_size > kMaxPalette) return fail(ctx, DMK_ERR_INVALID, "palette_size out of range");
    return DMK_OK;
}
static dmk_status exchange_alloc_local(dmk_exchange_t x) {
    DMK_CUDA(x->ctx, dev_alloc(&x->d_local, 2));
    DMK_CUDA(x->ctx, cudaMemsetAsync(x->d_local, 0, 2 * sizeof(uint32_t), x->ctx->stream));
    DMK_CUDA(x->ctx, cudaStreamSynchronize(x->ctx->stream));
    return DMK_OK;
}
void dmk_exchange_destroy(dmk_exchange_t x) {
    if (!x) return;
    cudaSetDevice(x->ctx->device);
    cudaDeviceSynchronize();   // the exchange may have run on side streams
    if (x->base) {
        if (x->kind == dmk_exchange_s::Owner) cudaFree(x->base);
        else if (x->kind == dmk_exchange_s::Opened) cudaIpcCloseMemHandle(x->base);
    }
    dev_free(x->d_local);
    dev_free(x->d_counts);
    dev_free(x->d_stage);
    delete x;
}
dmk_status dmk_exchange_create(dmk_context_t ctx, int world, int rank, int64_t capacity, int palette_size, uint8_t* handle_out,
                               dmk_exchange_t* out) {
    if (!ctx || !out) return fail(ctx, DMK_ERR_INVALID, "null argument");
    *out = nullptr;
    static_assert(sizeof(cudaIpcMemHandle_t) == DMK_EXCHANGE_HANDLE_BYTES, "DMK_EXCHANGE_HANDLE_BYTES");
    static_assert(kExchangeMaxRanks == DMK_EXCHANGE_MAX_RANKS, "DMK_EXCHANGE_MAX_RANKS");
    if (dmk_status st = exchange_check_shape(ctx, world, rank, capacity, palette_size)) return st;
    if (bind(ctx)) return DMK_ERR_CUDA;
    Building<dmk_exchange_s, dmk_exchange_destroy> x(new dmk_exchange_s());
    x->ctx = ctx;
    x->kind = dmk_exchange_s::Owner;
    x->world = world;
    x->rank = rank;
    x->capacity = capacity;
    x->palette_size = palette_size;
    DMK_CUDA(ctx, cudaMalloc(&x->base, x->bytes()));
    DMK_CUDA(ctx, cudaMemsetAsync(x->base, 0, kExchangeHeaderBytes, ctx->stream));
    DMK_CUDA(ctx, dev_alloc(&x->d_counts, 2 * kExchangeMaxRanks));
    DMK_CUDA(ctx, cudaMemsetAsync(x->d_counts, 0, 2 * kExchangeMaxRanks * sizeof(int32_t), ctx->stream));
    if (dmk_status st = exchange_alloc_local(x.get())) return st;
    if (handle_out) {
        cudaIpcMemHandle_t h;
        DMK_CUDA(ctx, cudaIpcGetMemHandle(&h, x->base));
        std::memcpy(handle_out, &h, sizeof h);
    }
    *out = x.release();
    return DMK_OK;
}
dmk_status dmk_exchange_open(dmk_context_t ctx, int world, int rank, int64_t capacity, int palette_size, const uint8_t* handle,
                             dmk_exchange_t* out) {
    if (!ctx || !out || !handle) return fail(ctx, DMK_ERR_INVALID, "null argument");
    *out = nullptr;
    if (dmk_status st = exchange_check_shape(ctx, world, rank, capacity, palette_size)) return st;
    if (bind(ctx)) return DMK_ERR_CUDA;
    Building<dmk_exchange_s, dmk_exchange_destroy> x(new dmk_exchange_s());
    x->ctx = ctx;
    x->kind = dmk_exchange_s::Opened;
    x->world = world;
    x->rank = rank;
    x->capacity = capacity;
    x->palette_size = palette_size;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, sizeof h);
    DMK_CUDA(ctx, cudaIpcOpenMemHandle(&x->base, h, cudaIpcMemLazyEnablePeerAccess));
    if (dmk_status st = exchange_alloc_local(x.get())) return st;
    *out = x.release();
    return DMK_OK;
}
dmk_status dmk_exchange_attach(dmk_exchange_t owner, dmk_context_t ctx, int rank, dmk_exchange_t* out) {
    if (!owner || !ctx || !out) return fail(ctx, DMK_ERR_INVALID, "null argument");
    *out = nullptr;
    if (owner->kind != dmk_exchange_s::Owner) return fail(ctx, DMK_ERR_INVALID, "attach needs the owning exchange");
    if (rank < 0 || rank >= owner->world) return fail(ctx, DMK_ERR_INVALID, "rank out of range");
    if (bind(ctx)) return DMK_ERR_CUDA;
    if (ctx->device != owner->ctx->device) {
        int can = 0;
        DMK_CUDA(ctx, cudaDeviceCanAccessPeer(&can, ctx->device, owner->ctx->device));
        if (!can) return fail(ctx, DMK_ERR_UNSUPPORTED, "no peer access to the render GPU");
        const cudaError_t e = cudaDeviceEnablePeerAccess(owner->ctx->device, 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
        else if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaDeviceEnablePeerAccess");
    }
    Building<dmk_exchange_s, dmk_exchange_destroy> x(new dmk_exchange_s());
    x->ctx = ctx;
    x->kind = dmk_exchange_s::Attached;
    x->world = owner->world;
    x->rank = rank;
    x->capacity = owner->capacity;
    x->palette_size = owner->palette_size;
    x->base = owner->base;
    x->timeout_ns = owner->timeout_ns;
    if (dmk_status st = exchange_alloc_local(x.get())) return st;
    *out = x.release();
    return DMK_OK;
}
dmk_status dmk_exchange_set_timeout_ms(dmk_exchange_t x, int64_t ms) {
    if (!x) return DMK_ERR_INVALID;
    if (ms < 1) return fail(x->ctx, DMK_ERR_INVALID, "timeout must be at least 1 ms");
    x->timeout_ns = static_cast<uint64_t>(ms) * 1000000ull;
    return DMK_OK;
}
static dmk_status exchange_check_crowd(dmk_crowd_t c, dmk_exchange_t x) {
    dmk_context_t ctx = c->ctx;
    if (x->ctx != ctx) return fail(ctx, DMK_ERR_INVALID, "the exchange belongs to another context");
    if (!c->arm) return fail(ctx, DMK_ERR_INVALID, "needs an armature");
    if (c->n > 0 && !c->culled_once) return fail(ctx, DMK_ERR_INVALID, "needs a cull step");
    if (c->arm->num_joints + 1 != x->palette_size) return fail(ctx, DMK_ERR_INVALID, "palette size differs from the exchange's");
    return DMK_OK;
}
// copy-engine transport, first half: the visible palettes into this rank's local staging buffer (pack kernel on `num_sms` SMs)
static dmk_status exchange_stage(dmk_crowd_t c, dmk_exchange_t x, int num_sms, cudaStream_t s) {
    dmk_context_t ctx = c->ctx;
    const size_t floats = static_cast<size_t>(x->capacity) * static_cast<size_t>(x->palette_size) * 16;
    if (!x->d_stage && floats) DMK_CUDA(ctx, dev_alloc(&x->d_stage, floats));
    KernelTimer timer(ctx, DMK_KERNEL_PACK, s);
    if (x->capacity > 0)
        DMK_CUDA(ctx, launch_pack_palettes(c->palette_3x4_only ? c->d_palette34 : c->d_palette, x->palette_size * 16, c->d_visible_ids, c->d_visible_count,
                                           x->capacity, x->d_stage, num_sms, s, c->palette_3x4_only));
    ++ctx->launches;
    return DMK_OK;
}
// second half: flow control, ONE DMA copy of the rank's slab into the render GPU's memory, publication
static dmk_status exchange_send(dmk_crowd_t c, dmk_exchange_t x, uint32_t frame, cudaStream_t s) {
    dmk_context_t ctx = c->ctx;
    const size_t floats = static_cast<size_t>(x->capacity) * static_cast<size_t>(x->palette_size) * 16;
    float* slab = x->buffer(frame) + static_cast<size_t>(x->rank) * floats;
    DMK_CUDA(ctx, launch_wait_consumed(x->header(), frame, x->d_local + 1, x->timeout_ns, s));
    if (floats) DMK_CUDA(ctx, cudaMemcpyAsync(slab, x->d_stage, floats * sizeof(float), cudaMemcpyDefault, s));
    DMK_CUDA(ctx, launch_publish(x->header(), x->rank, frame, c->d_visible_count, x->d_local + 1, s));
    ctx->launches += 2;
    x->last_pushed = frame;
    return DMK_OK;
}
dmk_status dmk_crowd_stage_visible_palettes(dmk_crowd_t c, dmk_exchange_t x, void* cuda_stream) {
    if (!c || !x) return DMK_ERR_INVALID;
    if (dmk_status st = exchange_check_crowd(c, x)) return st;
    if (bind(c->ctx)) return DMK_ERR_CUDA;
    x->staged = true;
    return exchange_stage(c, x, c->ctx->num_sms, cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : c->ctx->stream);
}
dmk_status dmk_exchange_send_staged(dmk_exchange_t x, dmk_crowd_t c, uint32_t frame, void* cuda_stream) {
    if (!c || !x) return DMK_ERR_INVALID;
    dmk_context_t ctx = c->ctx;
    if (dmk_status st = exchange_check_crowd(c, x)) return st;
    if (!x->staged) return fail(ctx, DMK_ERR_INVALID, "nothing staged: call dmk_crowd_stage_visible_palettes first");
    if (frame != x->last_pushed + 1) return fail(ctx, DMK_ERR_INVALID, "frames are pushed in order: expected the previous frame + 1");
    if (bind(ctx)) return DMK_ERR_CUDA;
    x->staged = false;
    return exchange_send(c, x, frame, cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->stream);
}
dmk_status dmk_crowd_push_visible_palettes(dmk_crowd_t c, dmk_exchange_t x, uint32_t frame, int num_ctas, void* cuda_stream) {
    if (!c || !x) return DMK_ERR_INVALID;
    dmk_context_t ctx = c->ctx;
    if (x->ctx != ctx) return fail(ctx, DMK_ERR_INVALID, "the exchange belongs to another context");
    if (!c->arm) return fail(ctx, DMK_ERR_INVALID, "needs an armature");
    if (c->n > 0 && !c->culled_once) return fail(ctx, DMK_ERR_INVALID, "needs a cull step");
    if (c->arm->num_joints + 1 != x->palette_size) return fail(ctx, DMK_ERR_INVALID, "palette size differs from the exchange's");
    if (frame != x->last_pushed + 1) return fail(ctx, DMK_ERR_INVALID, "frames are pushed in order: expected the previous frame + 1");
    if (bind(ctx)) return DMK_ERR_CUDA;
    cudaStream_t s = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->stream;
    const bool copy_engine = num_ctas < 0;
    if (num_ctas == 0) num_ctas = ctx->num_sms * 4;
    float* slab = x->buffer(frame) + static_cast<size_t>(x->rank) * static_cast<size_t>(x->capacity) * static_cast<size_t>(x->palette_size) * 16;
    if (copy_engine) {
        // pack locally on -num_ctas SMs' worth of blocks, one DMA copy of the whole rank slab (padding included: the host does not
        // know the count) into the render GPU's memory, publish.  Nothing of the transfer runs on an SM.
        if (dmk_status st = exchange_stage(c, x, -num_ctas, s)) return st;
        return exchange_send(c, x, frame, s);
    }
    {
        KernelTimer timer(ctx, DMK_KERNEL_PACK, s);
        DMK_CUDA(ctx, launch_push_palettes(c->palette_3x4_only ? c->d_palette34 : c->d_palette, x->palette_size * 16, c->d_visible_ids, c->d_visible_count,
                                           x->capacity, slab, x->header(), x->rank, frame, x->d_local, x->d_local + 1, x->timeout_ns, num_ctas, s,
                                           c->palette_3x4_only));
    }
    ++ctx->launches;
    x->last_pushed = frame;
    return DMK_OK;
}
dmk_status dmk_exchange_wait(dmk_exchange_t x, uint32_t frame, void* cuda_stream) {
    if (!x) return DMK_ERR_INVALID;
    dmk_context_t ctx = x->ctx;
    if (x->kind != dmk_exchange_s::Owner) return fail(ctx, DMK_ERR_INVALID, "only the render rank waits");
    if (frame != x->last_waited + 1) return fail(ctx, DMK_ERR_INVALID, "frames are waited for in order: expected the previous frame + 1");
    if (bind(ctx)) return DMK_ERR_CUDA;
    cudaStream_t s = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->stream;
    DMK_CUDA(ctx, launch_wait_published(x->header(), x->world, frame, x->d_counts + (frame & 1u) * kExchangeMaxRanks, x->capacity, x->timeout_ns, s));
    ++ctx->launches;
    x->last_waited = frame;
    return DMK_OK;
}
dmk_status dmk_exchange_frame_dev(dmk_exchange_t x, uint32_t frame, float** out_slab_dev, int32_t** out_counts_dev) {
    if (!x) return DMK_ERR_INVALID;
    if (x->kind != dmk_exchange_s::Owner) return fail(x->ctx, DMK_ERR_INVALID, "only the render rank reads the slab");
    if (out_slab_dev) *out_slab_dev = x->buffer(frame);
    if (out_counts_dev) *out_counts_dev = x->d_counts + (frame & 1u) * kExchangeMaxRanks;
    return DMK_OK;
}
dmk_status dmk_exchange_release(dmk_exchange_t x, uint32_t frame, void* cuda_stream) {
    if (!x) return DMK_ERR_INVALID;
    dmk_context_t ctx = x->ctx;
    if (x->kind != dmk_exchange_s::Owner) return fail(ctx, DMK_ERR_INVALID, "only the render rank releases");
    if (frame != x->last_released + 1 || frame > x->last_waited)
        return fail(ctx, DMK_ERR_INVALID, "frames are released in order, after their wait");
    if (bind(ctx)) return DMK_ERR_CUDA;
    cudaStream_t s = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->stream;
    DMK_CUDA(ctx, launch_release_frame(x->header(), frame, s));
    ++ctx->launches;
    x->last_released = frame;
    return DMK_OK;
}
dmk_status dmk_exchange_status(dmk_exchange_t x, void* cuda_stream, int* out_error) {
    if (!x) return DMK_ERR_INVALID;
    dmk_context_t ctx = x->ctx;
    if (bind(ctx)) return DMK_ERR_CUDA;
    cudaStream_t s = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->stream;
    uint32_t local = 0, owner = 0, overflow = 0;
    DMK_CUDA(ctx, cudaMemcpyAsync(&local, x->d_local + 1, sizeof local, cudaMemcpyDeviceToHost, s));
    if (x->kind == dmk_exchange_s::Owner) {
        DMK_CUDA(ctx, cudaMemcpyAsync(&owner, &x->header()->error, sizeof owner, cudaMemcpyDeviceToHost, s));
        DMK_CUDA(ctx, cudaMemcpyAsync(&overflow, &x->header()->overflow_frame, sizeof overflow, cudaMemcpyDeviceToHost, s));
    }
    DMK_CUDA(ctx, cudaStreamSynchronize(s));
    const int err = local ? static_cast<int>(local) : owner ? static_cast<int>(owner) : overflow ? static_cast<int>(kExchangeErrOverflow) : 0;
    if (out_error) *out_error = err;
    if (err == static_cast<int>(kExchangeErrOverflow))
        return fail(ctx, DMK_ERR_UNSUPPORTED, "palette exchange: a rank had more visible characters than the slab's capacity (last in frame " +
                                                  std::to_string(overflow) + "): the characters beyond it were not transferred");
    if (err == static_cast<int>(kExchangeErrTimeoutProducer))
        return fail(ctx, DMK_ERR_CUDA, "palette exchange: this rank's push gave up waiting for the render rank to release a buffer");
    if (err) return fail(ctx, DMK_ERR_CUDA, "palette exchange: the render rank gave up waiting for a rank's frame");
    return DMK_OK;
}

// ---- stand-alone cull ----------------------------------------------------------------------------------------
dmk_status dmk_cull_dev(dmk_context_t ctx, const float* corners_host, const float* world_inverse_dev, const float* bbox_dev, int64_t n,
                        uint32_t* vis_bits_dev) {
    if (!ctx) return DMK_ERR_INVALID;
    if (!corners_host || n < 0 || (n > 0 && (!world_inverse_dev || !bbox_dev || !vis_bits_dev))) return fail(ctx, DMK_ERR_INVALID, "null argument");
    if (bind(ctx)) return DMK_ERR_CUDA;
    CullParams p{};
    p.world_inverse = world_inverse_dev;
    p.bbox = bbox_dev;
    p.n = n;
    p.vis_bits = vis_bits_dev;
    std::memcpy(p.corners, corners_host, sizeof p.corners);
    DMK_CUDA(ctx, launch_cull(p, ctx->stream));
    ++ctx->launches;
    return DMK_OK;
}

dmk_status dmk_cull_trs_dev(dmk_context_t ctx, const float* corners_host, const float* position_dev, const float* rotation_dev,
                            const float* scale_dev, const float* bbox_dev, int64_t n, uint32_t* vis_bits_dev, float* world_inverse_out_dev) {
    if (!ctx) return DMK_ERR_INVALID;
    if (!corners_host || n < 0 || (n > 0 && (!position_dev || !rotation_dev || !scale_dev || !bbox_dev || !vis_bits_dev)))
        return fail(ctx, DMK_ERR_INVALID, "null argument");
    if (bind(ctx)) return DMK_ERR_CUDA;
    CullParams p{};
    p.position = position_dev;
    p.rotation = rotation_dev;
    p.scale = scale_dev;
    p.world_inverse_out = world_inverse_out_dev;
    p.bbox = bbox_dev;
    p.n = n;
    p.vis_bits = vis_bits_dev;
    std::memcpy(p.corners, corners_host, sizeof p.corners);
    DMK_CUDA(ctx, launch_cull(p, ctx->stream));
    ++ctx->launches;
    return DMK_OK;
}

dmk_status dmk_cull_host(dmk_context_t ctx, const float* view_proj, float near_depth, const float* world_inverse, const float* bbox,
                         int64_t n, uint32_t* vis_bits) {
    if (!ctx) return DMK_ERR_INVALID;
    if (!view_proj || n < 0 || (n > 0 && (!world_inverse || !bbox || !vis_bits))) return fail(ctx, DMK_ERR_INVALID, "null argument");
    if (n == 0) return DMK_OK;
    if (bind(ctx)) return DMK_ERR_CUDA;
    const size_t words = static_cast<size_t>((n + 31) / 32);
    const size_t need = static_cast<size_t>(n) * 88 + words * 4 + 256;
    if (ctx->cull_scratch_bytes < need) {
        dev_free(ctx->cull_scratch);
        ctx->cull_scratch = nullptr;
        ctx->cull_scratch_bytes = 0;
        DMK_CUDA(ctx, cudaMalloc(&ctx->cull_scratch, need));
        ctx->cull_scratch_bytes = need;
    }
    float* d_wi = static_cast<float*>(ctx->cull_scratch);
    float* d_bb = d_wi + static_cast<size_t>(n) * 16;
    uint32_t* d_bits = reinterpret_cast<uint32_t*>(d_bb + static_cast<size_t>(n) * 6 + ((n * 6) % 2 ? 1 : 0));
    float corners[24];
    frustum_corners_glm(view_proj, near_depth, corners);
    DMK_CUDA(ctx, cudaMemcpyAsync(d_wi, world_inverse, static_cast<size_t>(n) * 64, cudaMemcpyHostToDevice, ctx->stream));
    DMK_CUDA(ctx, cudaMemcpyAsync(d_bb, bbox, static_cast<size_t>(n) * 24, cudaMemcpyHostToDevice, ctx->stream));
    if (dmk_status st = dmk_cull_dev(ctx, corners, d_wi, d_bb, n, d_bits)) return st;
    DMK_CUDA(ctx, cudaMemcpyAsync(vis_bits, d_bits, words * 4, cudaMemcpyDeviceToHost, ctx->stream));
    DMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return DMK_OK;
}

}  // extern "C"
