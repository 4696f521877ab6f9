// api_results.cu -- C ABI of include/dmk.h: crowd results (palette, matrices, key indices, skinned vertices, visibility) and further skins per character
#include "api_internal.hpp"

extern "C" {

// ---- results -----------------------------------------------------------------------------------------------
const float* dmk_crowd_palette_dev(dmk_crowd_t c) { return c && !c->palette_3x4_only ? c->d_palette : nullptr; }
const float* dmk_crowd_skinned_dev(dmk_crowd_t c) { return c ? c->d_skinned : nullptr; }
const uint32_t* dmk_crowd_visibility_dev(dmk_crowd_t c) { return c ? c->d_vis_bits : nullptr; }
const int32_t* dmk_crowd_visible_ids_dev(dmk_crowd_t c) { return c ? c->d_visible_ids : nullptr; }
const int32_t* dmk_crowd_visible_count_dev(dmk_crowd_t c) { return c ? c->d_visible_count : nullptr; }

static dmk_status finish_and_check(dmk_crowd_t c) {
    dmk_context_t ctx = c->ctx;
    DMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (*c->h_error != 0)
        return fail(ctx, DMK_ERR_INVALID, "character " + std::to_string(*c->h_error - 1) + " references a clip outside the clip set");
    return DMK_OK;
}
static dmk_status check_range(dmk_crowd_t c, int64_t first, int64_t count, const void* out) {
    if (!c) return DMK_ERR_INVALID;
    if (!out || first < 0 || count < 0 || first + count > c->n) return fail(c->ctx, DMK_ERR_INVALID, "character range out of bounds");
    if (bind(c->ctx)) return DMK_ERR_CUDA;
    return DMK_OK;
}

dmk_status dmk_crowd_get_palette(dmk_crowd_t c, int64_t first, int64_t count, float* out) {
    if (dmk_status st = check_range(c, first, count, out)) return st;
    if (!c->arm) return fail(c->ctx, DMK_ERR_INVALID, "crowd has no armature");
    if (count == 0) return DMK_OK;
    const size_t per = static_cast<size_t>(c->arm->num_joints + 1) * 16;
    if (c->palette_3x4_only) {   // DMK_OPT_PALETTE_3X4_ONLY: expand the row-major 3x4 palette on the host
        DMK_CUDA(c->ctx, cudaMemcpyAsync(out, c->d_palette34 + first * per, count * per * 4, cudaMemcpyDeviceToHost, c->ctx->stream));
        if (dmk_status st = finish_and_check(c)) return st;
        for (size_t s = 0; s < static_cast<size_t>(count) * per; s += 16) {
            float m[16];
            for (int col = 0; col < 4; ++col) {
                for (int r = 0; r < 3; ++r) m[col * 4 + r] = out[s + r * 4 + col];
                m[col * 4 + 3] = col == 3 ? 1.f : 0.f;
            }
            std::memcpy(out + s, m, sizeof m);
        }
        return DMK_OK;
    }
    DMK_CUDA(c->ctx, cudaMemcpyAsync(out, c->d_palette + first * per, count * per * 4, cudaMemcpyDeviceToHost, c->ctx->stream));
    return finish_and_check(c);
}
dmk_status dmk_crowd_get_models(dmk_crowd_t c, int64_t first, int64_t count, float* out) {
    if (dmk_status st = check_range(c, first, count, out)) return st;
    if (!c->d_models) return fail(c->ctx, DMK_ERR_INVALID, "enable DMK_OUT_MODELS with dmk_crowd_enable_outputs before the step");
    const size_t per = static_cast<size_t>(c->skel->data.num_joints) * 16;
    DMK_CUDA(c->ctx, cudaMemcpyAsync(out, c->d_models + first * per, count * per * 4, cudaMemcpyDeviceToHost, c->ctx->stream));
    return finish_and_check(c);
}
dmk_status dmk_crowd_get_locals(dmk_crowd_t c, int64_t first, int64_t count, float* out) {
    if (dmk_status st = check_range(c, first, count, out)) return st;
    if (!c->d_locals) return fail(c->ctx, DMK_ERR_INVALID, "enable DMK_OUT_LOCALS with dmk_crowd_enable_outputs before the step");
    const size_t per = static_cast<size_t>(c->skel->data.num_soa()) * kSoaFloats;
    DMK_CUDA(c->ctx, cudaMemcpyAsync(out, c->d_locals + first * per, count * per * 4, cudaMemcpyDeviceToHost, c->ctx->stream));
    return finish_and_check(c);
}
dmk_status dmk_crowd_get_key_indices(dmk_crowd_t c, int64_t first, int64_t count, int32_t* out) {
    if (dmk_status st = check_range(c, first, count, out)) return st;
    if (!c->d_key_indices) return fail(c->ctx, DMK_ERR_INVALID, "enable DMK_OUT_KEY_INDICES with dmk_crowd_enable_outputs before the step");
    // device layout uses the current num_layers as the layer stride
    const size_t per = static_cast<size_t>(c->num_layers) * 3 * c->skel->data.num_soa() * 4 * 2;
    DMK_CUDA(c->ctx, cudaMemcpyAsync(out, c->d_key_indices + first * per, count * per * 4, cudaMemcpyDeviceToHost, c->ctx->stream));
    return finish_and_check(c);
}
dmk_status dmk_crowd_get_bone_matrices(dmk_crowd_t c, int64_t first, int64_t count, const float* dir, float* out, uint8_t* out_has_bone) {
    if (dmk_status st = check_range(c, first, count, out)) return st;
    dmk_context_t ctx = c->ctx;
    if (!dir) return fail(ctx, DMK_ERR_INVALID, "null direction");
    if (!c->d_models) return fail(ctx, DMK_ERR_INVALID, "enable DMK_OUT_MODELS with dmk_crowd_enable_outputs before the step");
    const int J = c->skel->data.num_joints;
    std::vector<int16_t> first_child(static_cast<size_t>(J), -1);
    for (int i = 0; i < J; ++i) {
        const int p = c->skel->data.parents[static_cast<size_t>(i)];
        if (p >= 0 && first_child[static_cast<size_t>(p)] < 0) first_child[static_cast<size_t>(p)] = static_cast<int16_t>(i);
    }
    if (out_has_bone)
        for (int i = 0; i < J; ++i) out_has_bone[i] = first_child[static_cast<size_t>(i)] >= 0;
    if (count == 0) return DMK_OK;
    int16_t* d_first = nullptr;
    float* d_out = nullptr;
    DMK_CUDA(ctx, upload(ctx, &d_first, first_child));
    cudaError_t e = dev_alloc(&d_out, static_cast<size_t>(count) * J * 16);
    // angleAxis(pi, axis) of glm::rotation's opposite-direction case: sin / cos of pi / 2 in float, from the host's libm
    const float half = 3.14159265358979323846264338327950288f * 0.5f;
    if (e == cudaSuccess)
        e = launch_bone_matrices(c->d_models + first * J * 16, d_first, J, count, dir, std::sin(half), std::cos(half), d_out, ctx->stream);
    ++ctx->launches;
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, static_cast<size_t>(count) * J * 64, cudaMemcpyDeviceToHost, ctx->stream);
    const cudaError_t e2 = cudaStreamSynchronize(ctx->stream);
    dev_free(d_first);
    dev_free(d_out);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "bone matrices");
    if (e2 != cudaSuccess) return cuda_fail(ctx, e2, "bone matrices");
    if (*c->h_error != 0)
        return fail(ctx, DMK_ERR_INVALID, "character " + std::to_string(*c->h_error - 1) + " references a clip outside the clip set");
    return DMK_OK;
}
// [count][V][9] on the host from either device layout
static dmk_status read_skinned(dmk_crowd_t c, dmk_mesh_t mesh, const float* d_skinned, int64_t first, int64_t count, float* out) {
    const size_t vpad = static_cast<size_t>(mesh->vpad), nv = static_cast<size_t>(mesh->num_vertices);
    const size_t pitch = vpad * 9, row = nv * 9;
    if (!c->skin_variant) {
        DMK_CUDA(c->ctx, cudaMemcpy2DAsync(out, row * 4, d_skinned + first * pitch, pitch * 4, row * 4, count, cudaMemcpyDeviceToHost, c->ctx->stream));
        return finish_and_check(c);
    }
    // on the device: nine component planes [9][vpad] (1, 2) or three float3 streams [3][vpad][3] (3) per character
    std::vector<float> planes;
    try {
        planes.resize(static_cast<size_t>(count) * pitch);
    } catch (const std::bad_alloc&) {   // no exception crosses the C ABI
        return fail(c->ctx, DMK_ERR_INVALID, "not enough host memory to re-interleave that many characters: read a smaller range");
    }
    DMK_CUDA(c->ctx, cudaMemcpyAsync(planes.data(), d_skinned + first * pitch, planes.size() * 4, cudaMemcpyDeviceToHost, c->ctx->stream));
    if (dmk_status st = finish_and_check(c)) return st;
    for (size_t i = 0; i < static_cast<size_t>(count); ++i) {
        if (c->skin_variant == 3) {
            for (size_t a = 0; a < 3; ++a) {
                const float* src = planes.data() + i * pitch + a * vpad * 3;
                float* dst = out + i * row + a * 3;
                for (size_t v = 0; v < nv; ++v)
                    for (size_t r = 0; r < 3; ++r) dst[v * 9 + r] = src[v * 3 + r];
            }
            continue;
        }
        for (size_t k = 0; k < 9; ++k) {
            const float* src = planes.data() + i * pitch + k * vpad;
            float* dst = out + i * row + k;
            for (size_t v = 0; v < nv; ++v) dst[v * 9] = src[v];
        }
    }
    return DMK_OK;
}
dmk_status dmk_crowd_get_skinned(dmk_crowd_t c, int64_t first, int64_t count, float* out) {
    if (dmk_status st = check_range(c, first, count, out)) return st;
    if (!c->mesh) return fail(c->ctx, DMK_ERR_INVALID, "crowd has no mesh");
    if (count == 0) return DMK_OK;
    return read_skinned(c, c->mesh, c->d_skinned, first, count, out);
}
// ---- several skinned meshes per character -----------------------------------------------------------------------
dmk_status dmk_crowd_add_skin(dmk_crowd_t c, dmk_armature_t arm, dmk_mesh_t mesh, int* out_skin) {
    if (!c) return DMK_ERR_INVALID;
    dmk_context_t ctx = c->ctx;
    if (!arm) return fail(ctx, DMK_ERR_INVALID, "a skin needs an armature");
    if (arm->ctx != ctx || (mesh && mesh->ctx != ctx)) return fail(ctx, DMK_ERR_INVALID, "armature / mesh belong to another context");
    if (mesh && mesh->palette_size != arm->num_joints + 1) return fail(ctx, DMK_ERR_INVALID, "mesh palette_size does not match the armature");
    if (!c->arm) return fail(ctx, DMK_ERR_INVALID, "the crowd was created without an armature: pass the first one to dmk_crowd_create");
    if (bind(ctx)) return DMK_ERR_CUDA;
    dmk_crowd_s::ExtraSkin x;
    x.arm = arm;
    x.mesh = mesh;
    const size_t ps = static_cast<size_t>(arm->num_joints + 1), n = static_cast<size_t>(c->n);
    cudaError_t e = dev_alloc(&x.d_palette, n * ps * 16);
    if (e == cudaSuccess) e = dev_alloc(&x.d_palette34, n * ps * 16);
    if (e == cudaSuccess && mesh) e = dev_alloc(&x.d_skinned, n * static_cast<size_t>(mesh->vpad) * 9);
    if (e != cudaSuccess) {
        dev_free(x.d_palette);
        dev_free(x.d_palette34);
        dev_free(x.d_skinned);
        return cuda_fail(ctx, e, "dmk_crowd_add_skin allocation");
    }
    if (mesh && c->skin_variant && plan_skin(mesh->vpad, mesh->palette_size, ctx->num_sms, c->skin_variant, c->skin_legacy).replicas != 8) {
        dev_free(x.d_palette);
        dev_free(x.d_palette34);
        dev_free(x.d_skinned);
        return fail(ctx, DMK_ERR_UNSUPPORTED, "this skinned-output layout needs a palette of at most 226 slots (181 for layouts 1 and 2)");
    }
    if (mesh) x.plan = plan_skin(mesh->vpad, mesh->palette_size, ctx->num_sms - c->reserved_sms, c->skin_variant, c->skin_legacy);
    c->extra_skins.push_back(x);
    if (out_skin) *out_skin = static_cast<int>(c->extra_skins.size());
    return DMK_OK;
}
int dmk_crowd_num_skins(dmk_crowd_t c) { return c ? (c->arm ? 1 : 0) + static_cast<int>(c->extra_skins.size()) : 0; }
namespace {
struct SkinView {
    dmk_armature_t arm = nullptr;
    dmk_mesh_t mesh = nullptr;
    float *palette = nullptr, *skinned = nullptr;
};
bool skin_view(dmk_crowd_t c, int skin, SkinView* v) {
    if (skin == 0) {
        *v = {c->arm, c->mesh, c->d_palette, c->d_skinned};
        return c->arm != nullptr;
    }
    if (skin < 0 || static_cast<size_t>(skin) > c->extra_skins.size()) return false;
    const auto& x = c->extra_skins[static_cast<size_t>(skin) - 1];
    *v = {x.arm, x.mesh, x.d_palette, x.d_skinned};
    return true;
}
}  // namespace
const float* dmk_crowd_skin_palette_dev(dmk_crowd_t c, int skin) {
    SkinView v;
    return c && skin_view(c, skin, &v) ? v.palette : nullptr;
}
const float* dmk_crowd_skin_skinned_dev(dmk_crowd_t c, int skin) {
    SkinView v;
    return c && skin_view(c, skin, &v) ? v.skinned : nullptr;
}
dmk_status dmk_crowd_get_skin_palette(dmk_crowd_t c, int skin, int64_t first, int64_t count, float* out) {
    if (dmk_status st = check_range(c, first, count, out)) return st;
    SkinView v;
    if (!skin_view(c, skin, &v)) return fail(c->ctx, DMK_ERR_INVALID, "no such skin");
    if (count == 0) return DMK_OK;
    const size_t per = static_cast<size_t>(v.arm->num_joints + 1) * 16;
    DMK_CUDA(c->ctx, cudaMemcpyAsync(out, v.palette + first * per, count * per * 4, cudaMemcpyDeviceToHost, c->ctx->stream));
    return finish_and_check(c);
}
dmk_status dmk_crowd_get_skin_skinned(dmk_crowd_t c, int skin, int64_t first, int64_t count, float* out) {
    if (dmk_status st = check_range(c, first, count, out)) return st;
    SkinView v;
    if (!skin_view(c, skin, &v)) return fail(c->ctx, DMK_ERR_INVALID, "no such skin");
    if (!v.mesh) return fail(c->ctx, DMK_ERR_INVALID, "the skin has no mesh");
    if (count == 0) return DMK_OK;
    return read_skinned(c, v.mesh, v.skinned, first, count, out);
}

dmk_status dmk_crowd_get_visibility(dmk_crowd_t c, uint32_t* out_bits, int64_t* out_visible_count) {
    if (!c) return DMK_ERR_INVALID;
    if (!c->culled_once) return fail(c->ctx, DMK_ERR_INVALID, "no cull step has run");
    if (bind(c->ctx)) return DMK_ERR_CUDA;
    const size_t words = static_cast<size_t>((c->n + 31) / 32);
    if (out_bits) DMK_CUDA(c->ctx, cudaMemcpyAsync(out_bits, c->d_vis_bits, words * 4, cudaMemcpyDeviceToHost, c->ctx->stream));
    int32_t count = 0;
    DMK_CUDA(c->ctx, cudaMemcpyAsync(&count, c->d_visible_count, 4, cudaMemcpyDeviceToHost, c->ctx->stream));
    if (dmk_status st = finish_and_check(c)) return st;
    if (out_visible_count) *out_visible_count = count;
    return DMK_OK;
}
dmk_status dmk_crowd_collect_visibility(dmk_crowd_t c, uint32_t* out_bits, int64_t* out_visible_count) {
    if (!c) return DMK_ERR_INVALID;
    if (c->vis_pending == 0) return fail(c->ctx, DMK_ERR_INVALID, "no step with DMK_STEP_READBACK_VISIBILITY is outstanding");
    if (bind(c->ctx)) return DMK_ERR_CUDA;
    const int b = c->vis_head;
    DMK_CUDA(c->ctx, cudaEventSynchronize(c->vis_ready[b]));
    const size_t words = static_cast<size_t>((c->n + 31) / 32);
    if (out_bits) std::memcpy(out_bits, c->h_vis[b], words * 4);
    if (out_visible_count) *out_visible_count = static_cast<int32_t>(c->h_vis[b][words]);
    c->vis_head ^= 1;
    --c->vis_pending;
    return DMK_OK;
}
dmk_status dmk_crowd_get_ratios(dmk_crowd_t c, float* out_ratio, uint8_t* out_looped) {
    if (!c) return DMK_ERR_INVALID;
    if (!c->layers_set) return fail(c->ctx, DMK_ERR_INVALID, "dmk_crowd_set_layers has not been called");
    if (bind(c->ctx)) return DMK_ERR_CUDA;
    const size_t ln = static_cast<size_t>(c->num_layers) * static_cast<size_t>(c->n);
    if (dmk_status st = join_layer_upload(c)) return st;
    if (out_ratio) DMK_CUDA(c->ctx, cudaMemcpyAsync(out_ratio, c->d_ratio, ln * 4, cudaMemcpyDeviceToHost, c->ctx->stream));
    if (out_looped) DMK_CUDA(c->ctx, cudaMemcpyAsync(out_looped, c->d_looped, ln, cudaMemcpyDeviceToHost, c->ctx->stream));
    return finish_and_check(c);
}

dmk_status dmk_crowd_pack_visible_palettes(dmk_crowd_t c, float* out_dev, int64_t capacity) {
    if (!c) return DMK_ERR_INVALID;
    if (!out_dev || capacity < 0) return fail(c->ctx, DMK_ERR_INVALID, "null output");
    if (!c->arm || !c->culled_once) return fail(c->ctx, DMK_ERR_INVALID, "needs an armature and a cull step");
    if (c->n == 0) return DMK_OK;
    if (bind(c->ctx)) return DMK_ERR_CUDA;
    {
        KernelTimer timer(c->ctx, DMK_KERNEL_PACK);
        DMK_CUDA(c->ctx, launch_pack_palettes(c->palette_3x4_only ? c->d_palette34 : c->d_palette, (c->arm->num_joints + 1) * 16, c->d_visible_ids,
                                              c->d_visible_count, capacity, out_dev, c->ctx->num_sms, c->ctx->stream, c->palette_3x4_only));
    }
    ++c->ctx->launches;
    return DMK_OK;
}

}  // extern "C"
