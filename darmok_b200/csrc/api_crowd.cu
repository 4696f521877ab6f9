// api_crowd.cu -- C ABI of include/dmk.h: crowd: creation, options, per-frame inputs, device animators, the frame step
#include "api_internal.hpp"

#include <cstdlib>

extern "C" {

// ------------------------------------------------------------------------------------------------------------
// crowd
// ------------------------------------------------------------------------------------------------------------
dmk_status dmk_crowd_create(dmk_context_t ctx, dmk_skeleton_t skel, dmk_clipset_t clips, dmk_armature_t arm, dmk_mesh_t mesh,
                            int64_t n, int max_layers, dmk_crowd_t* out) {
    if (!ctx || !skel || !clips || !out) return fail(ctx, DMK_ERR_INVALID, "null argument");
    *out = nullptr;
    if (n < 0 || n > 0x7fffffff) return fail(ctx, DMK_ERR_INVALID, "num_characters out of range");
    if (max_layers < 1 || max_layers > kMaxLayers) return fail(ctx, DMK_ERR_INVALID, "max_layers must be in [1, 32]");
    if (mesh && !arm) return fail(ctx, DMK_ERR_INVALID, "a mesh needs an armature");
    if (mesh && mesh->palette_size != arm->num_joints + 1)
        return fail(ctx, DMK_ERR_INVALID, "mesh palette_size does not match the armature");
    if (skel->data.num_joints == 0) return fail(ctx, DMK_ERR_INVALID, "skeleton has no joints");
    if (bind(ctx)) return DMK_ERR_CUDA;
    Building<dmk_crowd_s, dmk_crowd_destroy> crowd(new dmk_crowd_s());
    dmk_crowd_s* c = crowd.get();
    c->ctx = ctx;
    c->skel = skel;
    c->clips = clips;
    c->arm = arm;
    c->mesh = mesh;
    c->n = n;
    c->max_layers = max_layers;
    const size_t ln = static_cast<size_t>(max_layers) * static_cast<size_t>(n);
    DMK_CUDA(ctx, dev_alloc(&c->d_clip, ln));
    DMK_CUDA(ctx, dev_alloc(&c->d_ratio, ln));
    DMK_CUDA(ctx, dev_alloc(&c->d_weight, ln));
    DMK_CUDA(ctx, dev_alloc(&c->d_speed, ln));
    DMK_CUDA(ctx, dev_alloc(&c->d_loop, ln));
    DMK_CUDA(ctx, dev_alloc(&c->d_looped, ln));
    DMK_CUDA(ctx, dev_alloc(&c->d_local_matrices, static_cast<size_t>(n) * skel->data.num_soa() * 4 * 16));
    DMK_CUDA(ctx, dev_alloc(&c->d_error, 1));
    DMK_CUDA(ctx, cudaMemsetAsync(c->d_error, 0, sizeof(int32_t), ctx->stream));
    DMK_CUDA(ctx, cudaHostAlloc(reinterpret_cast<void**>(&c->h_error), sizeof(int32_t), cudaHostAllocDefault));
    *c->h_error = 0;
    if (arm) {
        const size_t ps = static_cast<size_t>(arm->num_joints + 1);
        DMK_CUDA(ctx, dev_alloc(&c->d_palette, static_cast<size_t>(n) * ps * 16));
        DMK_CUDA(ctx, dev_alloc(&c->d_palette34, static_cast<size_t>(n) * ps * 16));
    }
    if (mesh) {
        DMK_CUDA(ctx, dev_alloc(&c->d_skinned, static_cast<size_t>(n) * static_cast<size_t>(mesh->vpad) * 9));
        c->plan = plan_skin(mesh->vpad, mesh->palette_size, ctx->num_sms);
    }
    const size_t words = static_cast<size_t>((n + 31) / 32);
    DMK_CUDA(ctx, dev_alloc(&c->d_vis_bits, words + 1));
    DMK_CUDA(ctx, dev_alloc(&c->d_visible_ids, static_cast<size_t>(n) + 1));
    DMK_CUDA(ctx, dev_alloc(&c->d_visible_count, 1));
    DMK_CUDA(ctx, dev_alloc(&c->d_scratch, compact_scratch_ints(n)));
    DMK_CUDA(ctx, cudaMemsetAsync(c->d_visible_count, 0, sizeof(int32_t), ctx->stream));
    *out = crowd.release();
    return DMK_OK;
}

void dmk_crowd_destroy(dmk_crowd_t c) {
    if (!c) return;
    cudaSetDevice(c->ctx->device);
    cudaStreamSynchronize(c->ctx->stream);
    for (void* p : {(void*)c->d_clip, (void*)c->d_ratio, (void*)c->d_weight, (void*)c->d_speed, (void*)c->d_loop, (void*)c->d_looped,
                    (void*)c->d_program, (void*)c->d_local_matrices, (void*)c->d_palette, (void*)c->d_palette34, (void*)c->d_skinned, (void*)c->d_models, (void*)c->d_locals,
                    (void*)c->d_key_indices, (void*)c->d_error, (void*)c->d_world_inverse, (void*)c->d_bbox, (void*)c->d_position, (void*)c->d_rotation, (void*)c->d_scale, (void*)c->d_vis_bits,
                    (void*)c->d_visible_ids, (void*)c->d_visible_count, (void*)c->d_scratch, (void*)c->d_anim_states, (void*)c->d_anim_clips,
                    (void*)c->d_anim_transitions, (void*)c->d_anim_tables, (void*)c->d_anim_slots, (void*)c->d_anim_transition,
                    (void*)c->d_anim_transition_time, (void*)c->d_anim_events, (void*)c->d_anim_info, (void*)c->d_node_position, (void*)c->d_node_rotation,
                    (void*)c->d_node_scale, (void*)c->d_node_world, (void*)c->d_node_world_inverse, (void*)c->d_node_parent, (void*)c->d_entity_parent, (void*)c->d_anim_cmd, (void*)c->d_anim_status,
                    (void*)c->d_anim_cmd_speed, (void*)c->d_anim_blend, (void*)c->d_anim_speed, (void*)c->d_skin_next})
        dev_free(p);
    if (c->upload_stream) {
        cudaStreamSynchronize(c->upload_stream);
        cudaStreamDestroy(c->upload_stream);
    }
    if (c->layers_idle) cudaEventDestroy(c->layers_idle);
    if (c->upload_done) cudaEventDestroy(c->upload_done);
    for (auto& x : c->extra_skins) {
        dev_free(x.d_palette);
        dev_free(x.d_palette34);
        dev_free(x.d_skinned);
    }
    for (int b = 0; b < 2; ++b) {
        if (c->h_vis[b]) cudaFreeHost(c->h_vis[b]);
        if (c->vis_ready[b]) cudaEventDestroy(c->vis_ready[b]);
        if (c->h_stage[b]) cudaFreeHost(c->h_stage[b]);
        if (c->stage_done[b]) cudaEventDestroy(c->stage_done[b]);
    }
    if (c->h_error) cudaFreeHost(c->h_error);
    delete c;
}
int64_t dmk_crowd_size(dmk_crowd_t c) { return c ? c->n : 0; }
dmk_status dmk_crowd_set_option(dmk_crowd_t c, int option, int value) {
    if (!c) return DMK_ERR_INVALID;
    if (option == DMK_OPT_FORCE_KEY_SEARCH) {
        c->force_search = value != 0;
        return DMK_OK;
    }
    if (option == DMK_OPT_SKIN_RESERVED_SMS) {
        if (value < 0 || value >= c->ctx->num_sms) return fail(c->ctx, DMK_ERR_INVALID, "reserved SM count out of range");
        c->reserved_sms = value;
        if (c->mesh) c->plan = plan_skin(c->mesh->vpad, c->mesh->palette_size, c->ctx->num_sms - value, c->skin_variant, c->skin_legacy);
        for (auto& x : c->extra_skins)
            if (x.mesh) x.plan = plan_skin(x.mesh->vpad, x.mesh->palette_size, c->ctx->num_sms - value, c->skin_variant, c->skin_legacy);
        return DMK_OK;
    }
    if (option == DMK_OPT_SKIN_PLANAR_OUTPUT) {
        if (value < 0 || value > 4) return fail(c->ctx, DMK_ERR_INVALID, "DMK_OPT_SKIN_PLANAR_OUTPUT takes 0, 1, 2, 3 or 4");
        if (value) {
            auto supported = [&](dmk_mesh_t m) { return !m || plan_skin(m->vpad, m->palette_size, c->ctx->num_sms, value, c->skin_legacy).replicas == 8; };
            bool ok = supported(c->mesh);
            for (auto& x : c->extra_skins) ok = ok && supported(x.mesh);
            if (!ok) return fail(c->ctx, DMK_ERR_UNSUPPORTED, "this skinned-output layout needs a palette of at most 226 slots (181 for layouts 1 and 2)");
        }
        c->skin_variant = value;
        if (c->mesh) c->plan = plan_skin(c->mesh->vpad, c->mesh->palette_size, c->ctx->num_sms - c->reserved_sms, value, c->skin_legacy);
        for (auto& x : c->extra_skins)
            if (x.mesh) x.plan = plan_skin(x.mesh->vpad, x.mesh->palette_size, c->ctx->num_sms - c->reserved_sms, value, c->skin_legacy);
        return DMK_OK;
    }
    if (option == DMK_OPT_SKIN_ROUND1_KERNEL) {
        c->skin_legacy = value != 0;
        if (c->mesh) c->plan = plan_skin(c->mesh->vpad, c->mesh->palette_size, c->ctx->num_sms - c->reserved_sms, c->skin_variant, c->skin_legacy);
        for (auto& x : c->extra_skins)
            if (x.mesh) x.plan = plan_skin(x.mesh->vpad, x.mesh->palette_size, c->ctx->num_sms - c->reserved_sms, c->skin_variant, c->skin_legacy);
        return DMK_OK;
    }
    if (option == DMK_OPT_SKIN_REPLICATE_IN_SMEM) {
        c->skin_s2s = value != 0;
        return DMK_OK;
    }
    if (option == DMK_OPT_POSE_FUSED) {
        c->pose_fused = value != 0;
        return DMK_OK;
    }
    if (option == DMK_OPT_PALETTE_3X4_ONLY) {
        if (!c->mesh && !c->d_palette34) return fail(c->ctx, DMK_ERR_INVALID, "the crowd has no armature");
        c->palette_3x4_only = value != 0;
        return DMK_OK;
    }
    if (option == DMK_OPT_KEEP_WORLD_INVERSE) {
        if (bind(c->ctx)) return DMK_ERR_CUDA;
        if (value && !c->d_world_inverse) DMK_CUDA(c->ctx, dev_alloc(&c->d_world_inverse, static_cast<size_t>(c->n) * 16));
        c->keep_world_inverse = value != 0;
        return DMK_OK;
    }
    return fail(c->ctx, DMK_ERR_INVALID, "unknown option");
}

dmk_status dmk_skin_plan(int vertex_pitch, int palette_size, int num_sms, int layout, int64_t* out7) {
    if (!out7 || vertex_pitch < 8 || vertex_pitch % 8 || palette_size < 1 || palette_size > kMaxPalette || num_sms < 1 || layout < 0 || layout > 4)
        return fail(nullptr, DMK_ERR_INVALID, "dmk_skin_plan: argument out of range");
    const SkinPlan plan = plan_skin(vertex_pitch, palette_size, num_sms, layout);
    cudaGetLastError();   // without a device the occupancy query fails and the plan assumes one block per SM
    out7[0] = plan.threads;
    out7[1] = plan.num_tiles;
    out7[2] = plan.tile_vertices;
    out7[3] = plan.grid;
    out7[4] = plan.replicas;
    out7[5] = static_cast<int64_t>(plan.smem);
    out7[6] = (layout == 1 || layout == 2) ? 4 : plan.ws_vpt;
    return DMK_OK;
}

dmk_status dmk_crowd_enable_outputs(dmk_crowd_t c, uint32_t outputs) {
    if (!c) return DMK_ERR_INVALID;
    dmk_context_t ctx = c->ctx;
    if (bind(ctx)) return DMK_ERR_CUDA;
    const size_t n = static_cast<size_t>(c->n);
    if ((outputs & DMK_OUT_MODELS) && !c->d_models) DMK_CUDA(ctx, dev_alloc(&c->d_models, n * c->skel->data.num_joints * 16));
    if ((outputs & DMK_OUT_LOCALS) && !c->d_locals) DMK_CUDA(ctx, dev_alloc(&c->d_locals, n * c->skel->data.num_soa() * kSoaFloats));
    if ((outputs & DMK_OUT_KEY_INDICES) && !c->d_key_indices)
        DMK_CUDA(ctx, dev_alloc(&c->d_key_indices, n * c->max_layers * 3 * c->skel->data.num_soa() * 4 * 2));
    return DMK_OK;
}

// Returns a pinned staging buffer of at least `bytes` that no in-flight copy is reading; stage_release() must be
// called after the copies out of it have been enqueued.
static dmk_status stage_acquire(dmk_crowd_t c, size_t bytes, char** out) {
    const int b = c->stage_next;
    if (c->stage_done[b]) DMK_CUDA(c->ctx, cudaEventSynchronize(c->stage_done[b]));
    else DMK_CUDA(c->ctx, cudaEventCreateWithFlags(&c->stage_done[b], cudaEventDisableTiming));
    if (c->h_stage_bytes[b] < bytes) {
        if (c->h_stage[b]) cudaFreeHost(c->h_stage[b]);
        c->h_stage[b] = nullptr;
        c->h_stage_bytes[b] = 0;
        DMK_CUDA(c->ctx, cudaHostAlloc(&c->h_stage[b], bytes, cudaHostAllocDefault));
        c->h_stage_bytes[b] = bytes;
    }
    *out = static_cast<char*>(c->h_stage[b]);
    return DMK_OK;
}
static dmk_status stage_release(dmk_crowd_t c) {
    DMK_CUDA(c->ctx, cudaEventRecord(c->stage_done[c->stage_next], c->ctx->stream));
    c->stage_next ^= 1;
    return DMK_OK;
}

dmk_status dmk_crowd_set_layers(dmk_crowd_t c, int num_layers, const int32_t* clip, const float* ratio, const float* weight,
                                const float* speed, const uint8_t* loop, float threshold) {
    if (!c) return DMK_ERR_INVALID;
    dmk_context_t ctx = c->ctx;
    if (num_layers < 1 || num_layers > c->max_layers) return fail(ctx, DMK_ERR_INVALID, "num_layers out of range");
    if (!clip || !ratio || !weight) return fail(ctx, DMK_ERR_INVALID, "null layer array");
    if (bind(ctx)) return DMK_ERR_CUDA;
    const size_t ln = static_cast<size_t>(num_layers) * static_cast<size_t>(c->n);
    for (size_t i = 0; i < ln; ++i)
        if (clip[i] < 0 || clip[i] >= c->clips->num_clips)
            return fail(ctx, DMK_ERR_INVALID, "layer references clip " + std::to_string(clip[i]) + " outside the clip set");
    cudaStream_t s = ctx->stream;
    if (dmk_status st = join_layer_upload(c)) return st;
    c->layers_idle_recorded = false;   // an upload that follows orders itself behind these copies
    DMK_CUDA(ctx, cudaMemcpyAsync(c->d_clip, clip, ln * 4, cudaMemcpyHostToDevice, s));
    DMK_CUDA(ctx, cudaMemcpyAsync(c->d_ratio, ratio, ln * 4, cudaMemcpyHostToDevice, s));
    DMK_CUDA(ctx, cudaMemcpyAsync(c->d_weight, weight, ln * 4, cudaMemcpyHostToDevice, s));
    if (speed) {
        DMK_CUDA(ctx, cudaMemcpyAsync(c->d_speed, speed, ln * 4, cudaMemcpyHostToDevice, s));
    } else {
        std::vector<float> ones(ln, 1.f);
        DMK_CUDA(ctx, cudaMemcpyAsync(c->d_speed, ones.data(), ln * 4, cudaMemcpyHostToDevice, s));
        DMK_CUDA(ctx, cudaStreamSynchronize(s));
    }
    c->has_loop = loop != nullptr;
    if (loop) DMK_CUDA(ctx, cudaMemcpyAsync(c->d_loop, loop, ln, cudaMemcpyHostToDevice, s));
    DMK_CUDA(ctx, cudaMemsetAsync(c->d_looped, 0, ln, s));
    DMK_CUDA(ctx, cudaStreamSynchronize(s));  // host arrays are the caller's
    c->num_layers = num_layers;
    c->threshold = threshold;
    c->layers_set = true;
    c->use_program = false;
    return DMK_OK;
}

dmk_status dmk_crowd_set_programs(dmk_crowd_t c, int num_layers, const int32_t* clip, const float* ratio, const float* weight,
                                  const dmk_pose_program* programs) {
    if (!c) return DMK_ERR_INVALID;
    dmk_context_t ctx = c->ctx;
    static_assert(sizeof(dmk_pose_program) == sizeof(PoseProgram), "C ABI struct and device struct must match");
    if (num_layers < 1 || num_layers > c->max_layers) return fail(ctx, DMK_ERR_INVALID, "num_layers out of range");
    if (!clip || !ratio || !weight || !programs) return fail(ctx, DMK_ERR_INVALID, "null argument");
    if (bind(ctx)) return DMK_ERR_CUDA;
    const size_t n = static_cast<size_t>(c->n);
    for (size_t i = 0; i < n; ++i) {
        const dmk_pose_program& g = programs[i];
        if (g.top > DMK_TOP_SKIP || g.mode0 > DMK_GROUP_ZERO || g.mode1 > DMK_GROUP_ZERO || g.num_layers0 + g.num_layers1 > num_layers)
            return fail(ctx, DMK_ERR_INVALID, "malformed pose program for character " + std::to_string(i));
        if (g.top == DMK_TOP_SKIP || g.top == DMK_TOP_REST_POSE) continue;
        const int groups = g.top == DMK_TOP_TRANSITION ? 2 : 1;
        for (int k = 0; k < groups; ++k) {
            const int cnt = k ? g.num_layers1 : g.num_layers0, rest = k ? g.rest1 : g.rest0, mode = k ? g.mode1 : g.mode0;
            const float th = k ? g.threshold1 : g.threshold0;
            if (cnt < 1 || rest >= cnt) return fail(ctx, DMK_ERR_INVALID, "malformed pose program for character " + std::to_string(i));
            // BlendingJob::Validate: threshold > 0 (darmok reports "error in the blending job", src/skeleton_ozz.cpp:605-608)
            if (mode == DMK_GROUP_BLEND && !(th > 0.f)) return fail(ctx, DMK_ERR_JOB, "error in the blending job");
        }
        for (int l = 0; l < g.num_layers0 + g.num_layers1; ++l)
            if (clip[static_cast<size_t>(l) * n + i] < 0 || clip[static_cast<size_t>(l) * n + i] >= c->clips->num_clips)
                return fail(ctx, DMK_ERR_INVALID, "layer references a clip outside the clip set");
    }
    if (!c->d_program) DMK_CUDA(ctx, dev_alloc(&c->d_program, n));
    const size_t lbytes = static_cast<size_t>(num_layers) * n * 4, pbytes = n * sizeof(PoseProgram);
    char* h = nullptr;
    if (dmk_status st = stage_acquire(c, 3 * lbytes + pbytes, &h)) return st;
    std::memcpy(h, clip, lbytes);
    std::memcpy(h + lbytes, ratio, lbytes);
    std::memcpy(h + 2 * lbytes, weight, lbytes);
    std::memcpy(h + 3 * lbytes, programs, pbytes);
    cudaStream_t s = ctx->stream;
    if (dmk_status st = join_layer_upload(c)) return st;
    c->layers_idle_recorded = false;
    DMK_CUDA(ctx, cudaMemcpyAsync(c->d_clip, h, lbytes, cudaMemcpyHostToDevice, s));
    DMK_CUDA(ctx, cudaMemcpyAsync(c->d_ratio, h + lbytes, lbytes, cudaMemcpyHostToDevice, s));
    DMK_CUDA(ctx, cudaMemcpyAsync(c->d_weight, h + 2 * lbytes, lbytes, cudaMemcpyHostToDevice, s));
    DMK_CUDA(ctx, cudaMemcpyAsync(c->d_program, h + 3 * lbytes, pbytes, cudaMemcpyHostToDevice, s));
    if (dmk_status st = stage_release(c)) return st;
    c->num_layers = num_layers;
    c->layers_set = true;
    c->use_program = true;
    c->has_loop = false;
    return DMK_OK;
}

dmk_status dmk_crowd_update_layers(dmk_crowd_t c, int num_layers, const int32_t* clip, const float* ratio, const float* weight) {
    if (!c) return DMK_ERR_INVALID;
    dmk_context_t ctx = c->ctx;
    if (!c->layers_set || num_layers != c->num_layers) return fail(ctx, DMK_ERR_INVALID, "call dmk_crowd_set_layers first (same num_layers)");
    if (!clip || !ratio || !weight) return fail(ctx, DMK_ERR_INVALID, "null layer array");
    // same contract as dmk_crowd_set_layers: a clip index outside the set is refused here, on the host, not discovered by the kernel
    const size_t entries = static_cast<size_t>(num_layers) * static_cast<size_t>(c->n);
    const uint32_t num_clips = static_cast<uint32_t>(c->clips->num_clips);
    uint32_t worst = 0;
    for (size_t i = 0; i < entries; ++i) worst = std::max(worst, static_cast<uint32_t>(clip[i]));   // negative indices wrap to large values
    if (entries && worst >= num_clips) {
        size_t i = 0;
        while (static_cast<uint32_t>(clip[i]) < num_clips) ++i;
        return fail(ctx, DMK_ERR_INVALID, "character " + std::to_string(i % static_cast<size_t>(c->n)) + " references a clip outside the clip set");
    }
    if (bind(ctx)) return DMK_ERR_CUDA;
    const size_t bytes = static_cast<size_t>(num_layers) * static_cast<size_t>(c->n) * 4;
    char* h = nullptr;
    if (dmk_status st = stage_acquire(c, 3 * bytes, &h)) return st;
    std::memcpy(h, clip, bytes);
    std::memcpy(h + bytes, ratio, bytes);
    std::memcpy(h + 2 * bytes, weight, bytes);
    // The copies go on the crowd's upload stream: behind the last step that used the layer arrays, not behind everything the context's
    // stream already holds (the skinning kernel of the previous frame) -- the next step waits for them (join_layer_upload).
    if (!c->upload_stream) {
        DMK_CUDA(ctx, cudaStreamCreateWithFlags(&c->upload_stream, cudaStreamNonBlocking));
        DMK_CUDA(ctx, cudaEventCreateWithFlags(&c->upload_done, cudaEventDisableTiming));
    }
    if (!c->layers_idle_recorded) {   // nothing has stepped yet: order behind what the setters enqueued
        if (dmk_status st = mark_layers_idle(c)) return st;
    }
    cudaStream_t up = c->upload_stream;
    DMK_CUDA(ctx, cudaStreamWaitEvent(up, c->layers_idle, 0));
    DMK_CUDA(ctx, cudaMemcpyAsync(c->d_clip, h, bytes, cudaMemcpyHostToDevice, up));
    DMK_CUDA(ctx, cudaMemcpyAsync(c->d_ratio, h + bytes, bytes, cudaMemcpyHostToDevice, up));
    DMK_CUDA(ctx, cudaMemcpyAsync(c->d_weight, h + 2 * bytes, bytes, cudaMemcpyHostToDevice, up));
    DMK_CUDA(ctx, cudaEventRecord(c->upload_done, up));
    c->upload_pending = true;
    DMK_CUDA(ctx, cudaEventRecord(c->stage_done[c->stage_next], up));   // (stage_release on the upload stream)
    c->stage_next ^= 1;
    return DMK_OK;
}

// ---- device-side animators ------------------------------------------------------------
// (the host-side flattening of the definition is in animator_flatten.hpp)
static AnimatorParams animator_params(dmk_crowd_t c, float dt) {
    AnimatorParams a{};
    a.states = c->d_anim_states;
    a.clips = c->d_anim_clips;
    a.transitions = c->d_anim_transitions;
    a.num_states = c->anim_num_states;
    a.num_transitions = c->anim_num_transitions;
    a.tables = c->d_anim_tables;
    a.slots = c->d_anim_slots;
    a.transition_def = c->d_anim_transition;
    a.transition_time = c->d_anim_transition_time;
    a.cmd_state = c->d_anim_cmd;
    a.cmd_speed = c->d_anim_cmd_speed;
    a.in_blend = c->d_anim_blend;
    a.in_speed = c->d_anim_speed;
    a.events = c->d_anim_events;
    a.status = c->d_anim_status;
    a.n = c->n;
    a.num_layers = c->anim_layers;
    a.dt = dt;
    a.clip = c->d_clip;
    a.ratio = c->d_ratio;
    a.weight = c->d_weight;
    a.program = c->d_program;
    return a;
}

dmk_status dmk_crowd_set_animator(dmk_crowd_t c, const dmk_anim_state_def* states, int num_states, const dmk_anim_clip_def* clips,
                                  int num_clips, const dmk_anim_transition_def* transitions, int num_transitions) {
    if (!c) return DMK_ERR_INVALID;
    dmk_context_t ctx = c->ctx;
    if (!states || !clips || num_states < 1 || num_states > 32767 || num_clips < 1 || num_transitions < 0 || (num_transitions && !transitions))
        return fail(ctx, DMK_ERR_INVALID, "malformed animator definition");
    if (bind(ctx)) return DMK_ERR_CUDA;
    FlatAnimator flat;
    if (const std::string msg = flatten_animator(states, num_states, clips, num_clips, transitions, num_transitions, c->clips->durations, flat); !msg.empty())
        return fail(ctx, DMK_ERR_INVALID, msg);
    const int max_clips = flat.max_clips;
    if (2 * max_clips > c->max_layers)
        return fail(ctx, DMK_ERR_INVALID, "the crowd needs max_layers >= " + std::to_string(2 * max_clips) + " for this animator (two states blend in a transition)");
    const std::vector<AnimStateDef>& hs = flat.states;
    const std::vector<AnimClipDef>& hc = flat.clips;
    const std::vector<AnimTransitionDef>& ht = flat.transitions;
    const std::vector<float>& tables = flat.tables;
    for (void* p : {(void*)c->d_anim_states, (void*)c->d_anim_clips, (void*)c->d_anim_transitions, (void*)c->d_anim_tables}) dev_free(p);
    c->d_anim_states = nullptr; c->d_anim_clips = nullptr; c->d_anim_transitions = nullptr; c->d_anim_tables = nullptr;
    DMK_CUDA(ctx, upload(ctx, &c->d_anim_tables, tables));
    DMK_CUDA(ctx, upload(ctx, &c->d_anim_states, hs));
    DMK_CUDA(ctx, upload(ctx, &c->d_anim_clips, hc));
    DMK_CUDA(ctx, upload(ctx, &c->d_anim_transitions, ht));
    const size_t n = static_cast<size_t>(c->n);
    if (!c->d_anim_slots) {
        DMK_CUDA(ctx, dev_alloc(&c->d_anim_slots, static_cast<size_t>(kNumSlots) * kSlotWords * n));
        DMK_CUDA(ctx, dev_alloc(&c->d_anim_transition, n));
        DMK_CUDA(ctx, dev_alloc(&c->d_anim_transition_time, n));
        DMK_CUDA(ctx, dev_alloc(&c->d_anim_events, n));
        DMK_CUDA(ctx, dev_alloc(&c->d_anim_cmd, n));
        DMK_CUDA(ctx, dev_alloc(&c->d_anim_status, n));
        DMK_CUDA(ctx, dev_alloc(&c->d_anim_cmd_speed, n));
        DMK_CUDA(ctx, dev_alloc(&c->d_anim_blend, 2 * n));
        DMK_CUDA(ctx, dev_alloc(&c->d_anim_speed, n));
    }
    if (!c->d_program) DMK_CUDA(ctx, dev_alloc(&c->d_program, n));
    if (dmk_status st = join_layer_upload(c)) return st;
    DMK_CUDA(ctx, launch_animator_init(animator_params(c, 0.f), c->d_anim_cmd_speed, c->d_anim_blend, c->d_anim_speed, ctx->stream));
    DMK_CUDA(ctx, cudaMemsetAsync(c->d_anim_status, 0, n * sizeof(int32_t), ctx->stream));
    DMK_CUDA(ctx, cudaMemsetAsync(c->d_anim_events, 0xff, n * sizeof(AnimEvents), ctx->stream));
    DMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // the upload sources are locals
    ++ctx->launches;
    c->anim_num_states = num_states;
    c->anim_num_transitions = num_transitions;
    c->anim_layers = 2 * max_clips;
    c->animator_set = true;
    c->animator_ticked = false;
    return DMK_OK;
}

dmk_status dmk_crowd_animator_commands(dmk_crowd_t c, const int32_t* state, const float* speed) {
    if (!c) return DMK_ERR_INVALID;
    dmk_context_t ctx = c->ctx;
    if (!c->animator_set) return fail(ctx, DMK_ERR_INVALID, "dmk_crowd_set_animator has not been called");
    if (!state) return fail(ctx, DMK_ERR_INVALID, "null command array");
    if (bind(ctx)) return DMK_ERR_CUDA;
    const size_t n = static_cast<size_t>(c->n);
    // an unknown state name is "nothing happens" in the reference (play() returns false): ids outside the table do the same
    char* h = nullptr;
    if (dmk_status st = stage_acquire(c, 2 * n * 4, &h)) return st;
    std::memcpy(h, state, n * 4);
    float* hs = reinterpret_cast<float*>(h + n * 4);
    if (speed) std::memcpy(hs, speed, n * 4);
    else std::fill(hs, hs + n, 1.f);
    DMK_CUDA(ctx, cudaMemcpyAsync(c->d_anim_cmd, h, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    DMK_CUDA(ctx, cudaMemcpyAsync(c->d_anim_cmd_speed, hs, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    return stage_release(c);
}

dmk_status dmk_crowd_animator_set_inputs(dmk_crowd_t c, const float* blend_xy, const float* speed) {
    if (!c) return DMK_ERR_INVALID;
    dmk_context_t ctx = c->ctx;
    if (!c->animator_set) return fail(ctx, DMK_ERR_INVALID, "dmk_crowd_set_animator has not been called");
    if (!blend_xy && !speed) return DMK_OK;
    if (bind(ctx)) return DMK_ERR_CUDA;
    const size_t n = static_cast<size_t>(c->n);
    char* h = nullptr;
    if (dmk_status st = stage_acquire(c, 3 * n * 4, &h)) return st;
    if (blend_xy) {
        std::memcpy(h, blend_xy, 2 * n * 4);
        DMK_CUDA(ctx, cudaMemcpyAsync(c->d_anim_blend, h, 2 * n * 4, cudaMemcpyHostToDevice, ctx->stream));
    }
    if (speed) {
        std::memcpy(h + 2 * n * 4, speed, n * 4);
        DMK_CUDA(ctx, cudaMemcpyAsync(c->d_anim_speed, h + 2 * n * 4, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    }
    return stage_release(c);
}

dmk_status dmk_crowd_get_animator_events(dmk_crowd_t c, dmk_anim_events* out_events, int32_t* out_status) {
    if (!c) return DMK_ERR_INVALID;
    dmk_context_t ctx = c->ctx;
    static_assert(sizeof(dmk_anim_events) == sizeof(AnimEvents), "C ABI struct and device struct must match");
    if (!c->animator_set) return fail(ctx, DMK_ERR_INVALID, "dmk_crowd_set_animator has not been called");
    if (bind(ctx)) return DMK_ERR_CUDA;
    const size_t n = static_cast<size_t>(c->n);
    if (out_events) DMK_CUDA(ctx, cudaMemcpyAsync(out_events, c->d_anim_events, n * sizeof(AnimEvents), cudaMemcpyDeviceToHost, ctx->stream));
    if (out_status) DMK_CUDA(ctx, cudaMemcpyAsync(out_status, c->d_anim_status, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    DMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return DMK_OK;
}

dmk_status dmk_crowd_get_animator_state(dmk_crowd_t c, dmk_anim_state* out_states) {
    if (!c) return DMK_ERR_INVALID;
    dmk_context_t ctx = c->ctx;
    static_assert(sizeof(dmk_anim_state) == sizeof(AnimStateInfo), "C ABI struct and device struct must match");
    if (!c->animator_set) return fail(ctx, DMK_ERR_INVALID, "dmk_crowd_set_animator has not been called");
    if (!out_states) return fail(ctx, DMK_ERR_INVALID, "null output");
    if (bind(ctx)) return DMK_ERR_CUDA;
    const size_t n = static_cast<size_t>(c->n);
    if (!c->d_anim_info) DMK_CUDA(ctx, dev_alloc(&c->d_anim_info, n));
    if (dmk_status st = join_layer_upload(c)) return st;
    DMK_CUDA(ctx, launch_animator_query(animator_params(c, 0.f), c->d_anim_info, ctx->stream));
    ++ctx->launches;
    DMK_CUDA(ctx, cudaMemcpyAsync(out_states, c->d_anim_info, n * sizeof(AnimStateInfo), cudaMemcpyDeviceToHost, ctx->stream));
    DMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return DMK_OK;
}

dmk_status dmk_crowd_get_programs(dmk_crowd_t c, int num_layers, int32_t* out_clip, float* out_ratio, float* out_weight,
                                  dmk_pose_program* out_programs) {
    if (!c) return DMK_ERR_INVALID;
    dmk_context_t ctx = c->ctx;
    if (num_layers < 0 || num_layers > c->max_layers) return fail(ctx, DMK_ERR_INVALID, "num_layers out of range");
    if (out_programs && !c->d_program) return fail(ctx, DMK_ERR_INVALID, "the crowd has no pose programs");
    if (bind(ctx)) return DMK_ERR_CUDA;
    const size_t n = static_cast<size_t>(c->n), ln = static_cast<size_t>(num_layers) * n;
    if (dmk_status st = join_layer_upload(c)) return st;
    if (out_clip) DMK_CUDA(ctx, cudaMemcpyAsync(out_clip, c->d_clip, ln * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_ratio) DMK_CUDA(ctx, cudaMemcpyAsync(out_ratio, c->d_ratio, ln * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_weight) DMK_CUDA(ctx, cudaMemcpyAsync(out_weight, c->d_weight, ln * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_programs) DMK_CUDA(ctx, cudaMemcpyAsync(out_programs, c->d_program, n * sizeof(PoseProgram), cudaMemcpyDeviceToHost, ctx->stream));
    DMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return DMK_OK;
}

dmk_status dmk_crowd_set_instances(dmk_crowd_t c, const float* world_inverse, const float* bbox) {
    if (!c) return DMK_ERR_INVALID;
    dmk_context_t ctx = c->ctx;
    if (!world_inverse || !bbox) return fail(ctx, DMK_ERR_INVALID, "null instance array");
    if (bind(ctx)) return DMK_ERR_CUDA;
    const size_t n = static_cast<size_t>(c->n);
    if (!c->d_world_inverse) DMK_CUDA(ctx, dev_alloc(&c->d_world_inverse, n * 16));
    if (!c->d_bbox) DMK_CUDA(ctx, dev_alloc(&c->d_bbox, n * 6));
    DMK_CUDA(ctx, cudaMemcpyAsync(c->d_world_inverse, world_inverse, n * 64, cudaMemcpyHostToDevice, ctx->stream));
    DMK_CUDA(ctx, cudaMemcpyAsync(c->d_bbox, bbox, n * 24, cudaMemcpyHostToDevice, ctx->stream));
    DMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    c->instances_set = true;
    c->use_trs = false;
    return DMK_OK;
}

dmk_status dmk_crowd_set_transforms(dmk_crowd_t c, const float* position, const float* rotation, const float* scale, const float* bbox) {
    if (!c) return DMK_ERR_INVALID;
    dmk_context_t ctx = c->ctx;
    if (!position || !rotation || !scale || !bbox) return fail(ctx, DMK_ERR_INVALID, "null transform array");
    if (bind(ctx)) return DMK_ERR_CUDA;
    const size_t n = static_cast<size_t>(c->n);
    if (!c->d_position) DMK_CUDA(ctx, dev_alloc(&c->d_position, n * 3));
    if (!c->d_rotation) DMK_CUDA(ctx, dev_alloc(&c->d_rotation, n * 4));
    if (!c->d_scale) DMK_CUDA(ctx, dev_alloc(&c->d_scale, n * 3));
    if (!c->d_bbox) DMK_CUDA(ctx, dev_alloc(&c->d_bbox, n * 6));
    DMK_CUDA(ctx, cudaMemcpyAsync(c->d_position, position, n * 12, cudaMemcpyHostToDevice, ctx->stream));
    DMK_CUDA(ctx, cudaMemcpyAsync(c->d_rotation, rotation, n * 16, cudaMemcpyHostToDevice, ctx->stream));
    DMK_CUDA(ctx, cudaMemcpyAsync(c->d_scale, scale, n * 12, cudaMemcpyHostToDevice, ctx->stream));
    DMK_CUDA(ctx, cudaMemcpyAsync(c->d_bbox, bbox, n * 24, cudaMemcpyHostToDevice, ctx->stream));
    DMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    c->instances_set = true;
    c->use_trs = true;
    return DMK_OK;
}

dmk_status dmk_crowd_set_transform_nodes(dmk_crowd_t c, int num_nodes, const float* position, const float* rotation, const float* scale,
                                         const int32_t* node_parent, const int32_t* entity_parent) {
    if (!c) return DMK_ERR_INVALID;
    dmk_context_t ctx = c->ctx;
    if (num_nodes < 0) return fail(ctx, DMK_ERR_INVALID, "negative node count");
    if (bind(ctx)) return DMK_ERR_CUDA;
    DMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // a running cull may still read the old arrays
    for (void* p : {(void*)c->d_node_position, (void*)c->d_node_rotation, (void*)c->d_node_scale, (void*)c->d_node_world,
                    (void*)c->d_node_world_inverse, (void*)c->d_node_parent, (void*)c->d_entity_parent})
        dev_free(p);
    c->d_node_position = c->d_node_rotation = c->d_node_scale = c->d_node_world = c->d_node_world_inverse = nullptr;
    c->d_node_parent = c->d_entity_parent = nullptr;
    c->num_nodes = 0;
    if (num_nodes == 0) return DMK_OK;
    if (!position || !rotation || !scale || !node_parent || !entity_parent) return fail(ctx, DMK_ERR_INVALID, "null node array");
    const size_t m = static_cast<size_t>(num_nodes), n = static_cast<size_t>(c->n);
    for (size_t i = 0; i < m; ++i) {   // a forest: parents in range, no cycles, bounded depth
        int depth = 0;
        for (int k = static_cast<int>(i); k >= 0; k = node_parent[k]) {
            if (k >= num_nodes) return fail(ctx, DMK_ERR_INVALID, "node " + std::to_string(i) + ": parent index out of range");
            if (++depth > 32) return fail(ctx, DMK_ERR_INVALID, "node " + std::to_string(i) + ": parent chain deeper than 32 (or cyclic)");
        }
    }
    for (size_t i = 0; i < n; ++i)
        if (entity_parent[i] < -1 || entity_parent[i] >= num_nodes)
            return fail(ctx, DMK_ERR_INVALID, "entity " + std::to_string(i) + ": parent node out of range");
    cudaStream_t s = ctx->stream;
    DMK_CUDA(ctx, dev_alloc(&c->d_node_position, m * 3));
    DMK_CUDA(ctx, dev_alloc(&c->d_node_rotation, m * 4));
    DMK_CUDA(ctx, dev_alloc(&c->d_node_scale, m * 3));
    DMK_CUDA(ctx, dev_alloc(&c->d_node_world, m * 16));
    DMK_CUDA(ctx, dev_alloc(&c->d_node_world_inverse, m * 16));
    DMK_CUDA(ctx, dev_alloc(&c->d_node_parent, m));
    DMK_CUDA(ctx, dev_alloc(&c->d_entity_parent, n));
    DMK_CUDA(ctx, cudaMemcpyAsync(c->d_node_position, position, m * 12, cudaMemcpyHostToDevice, s));
    DMK_CUDA(ctx, cudaMemcpyAsync(c->d_node_rotation, rotation, m * 16, cudaMemcpyHostToDevice, s));
    DMK_CUDA(ctx, cudaMemcpyAsync(c->d_node_scale, scale, m * 12, cudaMemcpyHostToDevice, s));
    DMK_CUDA(ctx, cudaMemcpyAsync(c->d_node_parent, node_parent, m * 4, cudaMemcpyHostToDevice, s));
    DMK_CUDA(ctx, cudaMemcpyAsync(c->d_entity_parent, entity_parent, n * 4, cudaMemcpyHostToDevice, s));
    DMK_CUDA(ctx, cudaStreamSynchronize(s));
    c->num_nodes = num_nodes;
    return DMK_OK;
}

dmk_status dmk_crowd_get_world_inverses(dmk_crowd_t c, int64_t first, int64_t count, float* out) {
    if (!c) return DMK_ERR_INVALID;
    dmk_context_t ctx = c->ctx;
    if (!out || first < 0 || count < 0 || first + count > c->n) return fail(ctx, DMK_ERR_INVALID, "character range out of bounds");
    if (!c->d_world_inverse || !c->culled_once || (c->use_trs && !c->keep_world_inverse))
        return fail(ctx, DMK_ERR_INVALID, "needs DMK_OPT_KEEP_WORLD_INVERSE (or dmk_crowd_set_instances) and a cull step");
    if (bind(ctx)) return DMK_ERR_CUDA;
    DMK_CUDA(ctx, cudaMemcpyAsync(out, c->d_world_inverse + first * 16, static_cast<size_t>(count) * 64, cudaMemcpyDeviceToHost, ctx->stream));
    DMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return DMK_OK;
}
const float* dmk_crowd_world_inverse_dev(dmk_crowd_t c) { return c ? c->d_world_inverse : nullptr; }

dmk_status dmk_crowd_get_transform_nodes(dmk_crowd_t c, float* out_world, float* out_world_inverse) {
    if (!c) return DMK_ERR_INVALID;
    dmk_context_t ctx = c->ctx;
    if (!c->num_nodes || !c->culled_once) return fail(ctx, DMK_ERR_INVALID, "needs dmk_crowd_set_transform_nodes and a cull step");
    if (bind(ctx)) return DMK_ERR_CUDA;
    const size_t bytes = static_cast<size_t>(c->num_nodes) * 64;
    if (out_world) DMK_CUDA(ctx, cudaMemcpyAsync(out_world, c->d_node_world, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_world_inverse) DMK_CUDA(ctx, cudaMemcpyAsync(out_world_inverse, c->d_node_world_inverse, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    DMK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return DMK_OK;
}

dmk_status dmk_frustum_corners(const float* view_proj, float near_depth, float* out_corners) {
    if (!view_proj || !out_corners) return DMK_ERR_INVALID;
    frustum_corners_glm(view_proj, near_depth, out_corners);
    return DMK_OK;
}

dmk_status dmk_crowd_set_camera(dmk_crowd_t c, const float* view_proj, float near_depth) {
    if (!c) return DMK_ERR_INVALID;
    if (!view_proj) return fail(c->ctx, DMK_ERR_INVALID, "null view_proj");
    frustum_corners_glm(view_proj, near_depth, c->corners);
    c->camera_set = true;
    return DMK_OK;
}
dmk_status dmk_crowd_set_frustum_corners(dmk_crowd_t c, const float* corners) {
    if (!c) return DMK_ERR_INVALID;
    if (!corners) return fail(c->ctx, DMK_ERR_INVALID, "null corners");
    std::memcpy(c->corners, corners, sizeof c->corners);
    c->camera_set = true;
    return DMK_OK;
}

dmk_status dmk_crowd_step(dmk_crowd_t c, float dt, uint32_t flags) {
    if (!c) return DMK_ERR_INVALID;
    dmk_context_t ctx = c->ctx;
    if (bind(ctx)) return DMK_ERR_CUDA;
    cudaStream_t s = ctx->stream;
    if ((flags & DMK_STEP_ADVANCE) && !(flags & DMK_STEP_POSE))
        return fail(ctx, DMK_ERR_UNSUPPORTED, "DMK_STEP_ADVANCE runs inside the pose kernel: combine it with DMK_STEP_POSE");
    // refused before anything is enqueued: a third outstanding read-back has no slot to land in
    if ((flags & DMK_STEP_CULL) && (flags & DMK_STEP_READBACK_VISIBILITY) && c->vis_pending == 2)
        return fail(ctx, DMK_ERR_INVALID, "two visibility read-backs are already outstanding: collect one first");
    if (flags & (DMK_STEP_ANIMATOR | DMK_STEP_POSE | DMK_STEP_ADVANCE)) {
        if (dmk_status st = join_layer_upload(c)) return st;
    }
    if (flags & DMK_STEP_ANIMATOR) {
        if (!c->animator_set) return fail(ctx, DMK_ERR_INVALID, "dmk_crowd_set_animator has not been called");
        if (flags & DMK_STEP_ADVANCE) return fail(ctx, DMK_ERR_UNSUPPORTED, "the device animators own the clip clocks: do not combine them with DMK_STEP_ADVANCE");
        const AnimatorParams a = animator_params(c, dt);
        {
            KernelTimer timer(ctx, DMK_KERNEL_ANIMATOR);
            DMK_CUDA(ctx, launch_animator(a, s));
        }
        ++ctx->launches;
        c->num_layers = c->anim_layers;
        c->layers_set = true;
        c->use_program = true;
        c->has_loop = false;
        c->animator_ticked = true;
    }
    if (flags & DMK_STEP_POSE) {
        if (!c->layers_set) return fail(ctx, DMK_ERR_INVALID, "dmk_crowd_set_layers has not been called");
        // BlendingJob::Validate: threshold > 0 (darmok: "error in the blending job", src/skeleton_ozz.cpp:605-608)
        if (!c->use_program && c->num_layers > 1 && !(c->threshold > 0.f)) return fail(ctx, DMK_ERR_JOB, "error in the blending job");
        if (c->use_program && (flags & DMK_STEP_ADVANCE))
            return fail(ctx, DMK_ERR_UNSUPPORTED, "pose programs carry host-resolved ratios: do not combine them with DMK_STEP_ADVANCE");
        PoseParams p{};
        p.clips.dir = c->clips->d_dir;
        for (int k = 0; k < 3; ++k) {
            p.clips.ratio[k] = c->clips->d_ratio[k];
            p.clips.value[k] = c->clips->d_value[k];
            p.clips.stream[k] = c->clips->d_stream[k];
        }
        p.clips.duration = c->clips->d_duration;
        if (c->clips->has_buckets && !c->force_search) {
            p.clips.bucket_dir = c->clips->d_bucket_dir;
            p.clips.bucket_a = c->clips->d_bucket_a;
            p.clips.bucket_b = c->clips->d_bucket_b;
            p.clips.bucket_c = c->clips->d_bucket_c;
        }
        p.clips.num_clips = c->clips->num_clips;
        p.clips.padded_tracks = c->clips->padded_tracks;
        p.skel.parents = c->skel->d_parents;
        p.skel.schedule = c->skel->d_schedule;
        p.skel.rest = c->skel->d_rest;
        p.skel.num_joints = c->skel->data.num_joints;
        p.skel.padded_joints = c->skel->data.num_soa() * 4;
        p.skel.num_soa = c->skel->data.num_soa();
        p.skel.rounds = c->skel->rounds;
        if (c->arm) {
            p.arm.remap = c->arm->d_remap;
            p.arm.inv_bind = c->arm->d_inv_bind;
            p.arm.num_joints = c->arm->num_joints;
        }
        p.n = c->n;
        p.num_layers = c->num_layers;
        p.clip = c->d_clip;
        p.ratio = c->d_ratio;
        p.weight = c->d_weight;
        p.speed = c->d_speed;
        p.loop = c->has_loop ? c->d_loop : nullptr;
        p.looped = c->d_looped;
        p.threshold = c->threshold;
        p.program = c->use_program ? c->d_program : nullptr;
        p.dt = dt;
        p.advance = (flags & DMK_STEP_ADVANCE) ? 1 : 0;
        p.palette = c->palette_3x4_only ? nullptr : c->d_palette;
        p.palette34 = c->d_palette34;
        p.fused = (c->pose_fused && c->extra_skins.empty()) ? 1 : 0;
        p.models = c->d_models;
        p.locals = c->d_locals;
        p.key_indices = c->d_key_indices;
        p.error_flag = c->d_error;
        // the error word describes THIS pose step: a bad clip index of an earlier frame does not outlive the layers that had it
        DMK_CUDA(ctx, cudaMemsetAsync(c->d_error, 0, sizeof(int32_t), s));
        p.scratch = c->d_local_matrices;
        int pose_launches = 0;
        {
            KernelTimer timer(ctx, DMK_KERNEL_POSE);
            DMK_CUDA(ctx, launch_pose(p, ctx->num_sms, s, &pose_launches));
            // every further armature: its own palette from the same local matrices (beforeRenderEntity runs once per Skinnable)
            for (const auto& x : c->extra_skins) {
                PoseParams q = p;
                q.arm.remap = x.arm->d_remap;
                q.arm.inv_bind = x.arm->d_inv_bind;
                q.arm.num_joints = x.arm->num_joints;
                q.palette = x.d_palette;
                q.palette34 = x.d_palette34;
                q.models = nullptr;
                q.locals = nullptr;
                q.key_indices = nullptr;
                DMK_CUDA(ctx, launch_model(q, ctx->num_sms, s, &pose_launches));
            }
        }
        ctx->launches += pose_launches;
    }
    if (flags & (DMK_STEP_ANIMATOR | DMK_STEP_POSE | DMK_STEP_ADVANCE)) {
        if (dmk_status st = mark_layers_idle(c)) return st;   // the layer arrays may be overwritten by the next frame's upload from here on
    }
    if (flags & DMK_STEP_CULL) {
        if (!c->instances_set || !c->camera_set)
            return fail(ctx, DMK_ERR_INVALID, "cull needs dmk_crowd_set_instances and dmk_crowd_set_camera");
        CullParams p{};
        int node_launches = 0;
        if (c->use_trs) {
            p.position = c->d_position;
            p.rotation = c->d_rotation;
            p.scale = c->d_scale;
            if (c->num_nodes) {
                DMK_CUDA(ctx, launch_transform_nodes(c->d_node_position, c->d_node_rotation, c->d_node_scale, c->d_node_parent, c->num_nodes,
                                                     c->d_node_world, c->d_node_world_inverse, s));
                node_launches = 1;
                p.entity_parent = c->d_entity_parent;
                p.node_world_inverse = c->d_node_world_inverse;
            }
            if (c->keep_world_inverse) p.world_inverse_out = c->d_world_inverse;
        } else {
            p.world_inverse = c->d_world_inverse;
        }
        p.bbox = c->d_bbox;
        p.n = c->n;
        p.vis_bits = c->d_vis_bits;
        std::memcpy(p.corners, c->corners, sizeof p.corners);
        int extra = 0;
        {
            KernelTimer timer(ctx, DMK_KERNEL_CULL);
            DMK_CUDA(ctx, launch_cull(p, s));
            DMK_CUDA(ctx, launch_compact(c->d_vis_bits, c->n, c->d_visible_ids, c->d_visible_count, c->d_scratch, s, &extra));
        }
        ctx->launches += 1 + extra + node_launches;
        c->culled_once = true;
        if (flags & DMK_STEP_READBACK_VISIBILITY) {
            if (c->vis_pending == 2) return fail(ctx, DMK_ERR_INVALID, "two visibility read-backs are already outstanding: collect one first");
            const int b = (c->vis_head + c->vis_pending) & 1;
            const size_t words = static_cast<size_t>((c->n + 31) / 32);
            for (int k = 0; k < 2 && !c->h_vis[1]; ++k) {   // both ring slots at first use: pinned allocation synchronises
                DMK_CUDA(ctx, cudaHostAlloc(reinterpret_cast<void**>(&c->h_vis[k]), (words + 1) * 4, cudaHostAllocDefault));
                DMK_CUDA(ctx, cudaEventCreateWithFlags(&c->vis_ready[k], cudaEventDisableTiming));
            }
            DMK_CUDA(ctx, cudaMemcpyAsync(c->h_vis[b], c->d_vis_bits, words * 4, cudaMemcpyDeviceToHost, s));
            DMK_CUDA(ctx, cudaMemcpyAsync(c->h_vis[b] + words, c->d_visible_count, 4, cudaMemcpyDeviceToHost, s));
            DMK_CUDA(ctx, cudaEventRecord(c->vis_ready[b], s));
            ++c->vis_pending;
        }
    }
    if (flags & DMK_STEP_SKIN) {
        bool any_mesh = c->mesh != nullptr;
        for (const auto& x : c->extra_skins) any_mesh = any_mesh || x.mesh != nullptr;
        if (!any_mesh || !c->arm) return fail(ctx, DMK_ERR_INVALID, "skinning needs a mesh and an armature");
        const bool visible_only = (flags & DMK_STEP_SKIN_VISIBLE_ONLY) != 0;
        if (visible_only && !c->culled_once) return fail(ctx, DMK_ERR_INVALID, "DMK_STEP_SKIN_VISIBLE_ONLY needs a cull step");
        SkinParams p{};
        {   // per-tile counters of the dynamic character assignment (DMK_SKIN_STATIC_SPLIT=1 in the environment: the static split, for A/B)
            static const bool static_split = [] { const char* e = std::getenv("DMK_SKIN_STATIC_SPLIT"); return e && std::atoi(e) == 1; }();
            int tiles = c->mesh ? c->plan.num_tiles : 0;
            for (const auto& x : c->extra_skins) tiles = std::max(tiles, x.mesh ? x.plan.num_tiles : 0);
            if (!static_split && tiles > c->skin_next_capacity) {
                dev_free(c->d_skin_next);
                c->d_skin_next = nullptr;
                c->skin_next_capacity = 0;
                DMK_CUDA(ctx, dev_alloc(&c->d_skin_next, static_cast<size_t>(tiles)));
                c->skin_next_capacity = tiles;
            }
            p.next_item = static_split ? nullptr : c->d_skin_next;
        }
        p.char_ids = visible_only ? c->d_visible_ids : nullptr;
        p.char_count = visible_only ? c->d_visible_count : nullptr;
        p.n = c->n;
        if (c->mesh) {
            p.palette34 = c->d_palette34;
            p.pos = c->mesh->d_pos;
            p.nrm = c->mesh->d_nrm;
            p.tan = c->mesh->d_tan;
            p.weights = c->mesh->d_weights;
            p.slots = c->mesh->d_slots;
            p.out = c->d_skinned;
            p.vpad = c->mesh->vpad;
            p.palette_size = c->mesh->palette_size;
            p.num_tiles = c->plan.num_tiles;
            p.tile_vertices = c->plan.tile_vertices;
            p.vertex_order_mode = c->mesh->strided ? 3 : c->mesh->clustered ? 2 : c->mesh->sorted ? 1 : 0;
            {
                KernelTimer timer(ctx, DMK_KERNEL_SKIN);
                SkinPlan plan = c->plan;
                plan.replicate_in_smem = c->skin_s2s;
                DMK_CUDA(ctx, launch_skin(p, plan, s));
            }
            ++ctx->launches;
        }
        for (const auto& x : c->extra_skins) {
            if (!x.mesh) continue;
            SkinParams q = p;
            q.palette34 = x.d_palette34;
            q.pos = x.mesh->d_pos;
            q.nrm = x.mesh->d_nrm;
            q.tan = x.mesh->d_tan;
            q.weights = x.mesh->d_weights;
            q.slots = x.mesh->d_slots;
            q.out = x.d_skinned;
            q.vpad = x.mesh->vpad;
            q.palette_size = x.mesh->palette_size;
            q.num_tiles = x.plan.num_tiles;
            q.tile_vertices = x.plan.tile_vertices;
            q.vertex_order_mode = x.mesh->strided ? 3 : x.mesh->clustered ? 2 : x.mesh->sorted ? 1 : 0;
            {
                KernelTimer timer(ctx, DMK_KERNEL_SKIN);
                SkinPlan plan = x.plan;
                plan.replicate_in_smem = c->skin_s2s;
                DMK_CUDA(ctx, launch_skin(q, plan, s));
            }
            ++ctx->launches;
        }
    }
    if (flags & DMK_STEP_POSE)
        DMK_CUDA(ctx, cudaMemcpyAsync(c->h_error, c->d_error, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    return DMK_OK;
}

}  // extern "C"
