// api_assets.cu -- C ABI of include/dmk.h: skeleton, clips, clip set (SoA key tracks + bucket tables), armature, mesh
#include "api_internal.hpp"

namespace {
struct PackedComponent {
    std::vector<float> ratio;
    std::vector<uint2> value;
    std::vector<int32_t> stream;
};
uint2 pack_key(const Float3Key& k) {
    return make_uint2(static_cast<uint32_t>(k.value[0]) | (static_cast<uint32_t>(k.value[1]) << 16), static_cast<uint32_t>(k.value[2]));
}
uint2 pack_key(const QuaternionKey& k) {
    return make_uint2(static_cast<uint32_t>(static_cast<uint16_t>(k.value[0])) | (static_cast<uint32_t>(static_cast<uint16_t>(k.value[1])) << 16),
                      static_cast<uint32_t>(static_cast<uint16_t>(k.value[2])) | (static_cast<uint32_t>(k.largest) << 16) |
                          (static_cast<uint32_t>(k.sign) << 18) | (static_cast<uint32_t>(k.format) << 19));
}
// keys of the ozz stream regrouped per track (stream order within a track = time order)
template <typename Key>
bool regroup(const std::vector<Key>& keys, int padded, PackedComponent& pc, uint2* dir, std::string& err) {
    std::vector<uint32_t> count(padded, 0);
    for (const auto& k : keys) ++count[k.track];
    uint32_t first = static_cast<uint32_t>(pc.ratio.size());
    std::vector<uint32_t> cursor(padded);
    for (int t = 0; t < padded; ++t) {
        if (count[t] < 2) {
            err = "corrupt ozz animation archive (a track has fewer than 2 keys)";
            return false;
        }
        dir[t] = make_uint2(first, count[t]);
        cursor[t] = first;
        first += count[t];
    }
    pc.ratio.resize(first);
    pc.value.resize(first);
    pc.stream.resize(first);
    for (size_t i = 0; i < keys.size(); ++i) {
        const uint32_t at = cursor[keys[i].track]++;
        pc.ratio[at] = keys[i].ratio;
        pc.value[at] = pack_key(keys[i]);
        pc.stream[at] = static_cast<int32_t>(i);
    }
    for (int t = 0; t < padded; ++t)
        for (uint32_t i = dir[t].x + 1; i < dir[t].x + dir[t].y; ++i)
            if (pc.ratio[i] < pc.ratio[i - 1]) {
                err = "corrupt ozz animation archive (track keys are not sorted by ratio)";
                return false;
            }
    return true;
}

// Bucket table of one clip (see ClipSetDev).  Returns false when no G <= kMaxBuckets separates the keys.
constexpr int kMinBuckets = 16, kMaxBuckets = 1024;
struct BucketTable {
    std::vector<uint2> dir;
    std::vector<float4> a;
    std::vector<uint4> b;
    std::vector<uint2> c;
};
bool build_buckets(const PackedComponent* pc, const uint2* dir, int padded, BucketTable& tbl) {
    for (int G = kMinBuckets; G <= kMaxBuckets; G *= 2) {
        bool ok = true;
        const size_t first = tbl.a.size();
        tbl.a.resize(first + (size_t)3 * G * padded);
        tbl.b.resize(tbl.a.size());
        tbl.c.resize(tbl.a.size());
        for (int comp = 0; comp < 3 && ok; ++comp) {
            for (int t = 0; t < padded && ok; ++t) {
                const uint2 d = dir[comp * padded + t];
                const float* r = pc[comp].ratio.data() + d.x;
                const uint2* v = pc[comp].value.data() + d.x;
                const int n = (int)d.y;
                int base = 1;
                for (int b = 0; b < G; ++b) {
                    const float tb = (float)b / (float)G, te = (float)(b + 1) / (float)G;
                    while (base < n - 1 && !(r[base] > tb)) ++base;  // first key (>= 1) with ratio > tb, else the last
                    if (!(base >= n - 2 || r[base + 1] >= te)) { ok = false; break; }
                    const size_t e = first + ((size_t)comp * G + b) * padded + t;
                    const bool last = base == n - 1;
                    tbl.a[e] = make_float4(r[base - 1], r[base], last ? -1.f : r[base + 1], 0.f);
                    std::memcpy(&tbl.a[e].w, &base, sizeof(int));
                    tbl.b[e] = make_uint4(v[base - 1].x, v[base - 1].y, v[base].x, v[base].y);
                    tbl.c[e] = last ? v[base] : v[base + 1];
                }
            }
        }
        if (ok) {
            tbl.dir.push_back(make_uint2((uint32_t)first, (uint32_t)G));
            return true;
        }
        tbl.a.resize(first);
        tbl.b.resize(first);
        tbl.c.resize(first);
    }
    return false;
}
}  // namespace

extern "C" {

// ------------------------------------------------------------------------------------------------------------
// skeleton
// ------------------------------------------------------------------------------------------------------------
dmk_status dmk_skeleton_load_ozz(dmk_context_t ctx, const void* bytes, size_t size, dmk_skeleton_t* out) {
    if (!ctx || !out) return fail(ctx, DMK_ERR_INVALID, "null argument");
    *out = nullptr;
    Building<dmk_skeleton_s, dmk_skeleton_destroy> skel(new dmk_skeleton_s());
    skel->ctx = ctx;
    std::string err;
    if (!parse_skeleton(bytes, size, skel->data, err)) return fail(ctx, DMK_ERR_FORMAT, err);
    if (bind(ctx)) return DMK_ERR_CUDA;
    const int J = skel->data.num_joints;
    // level-order schedule of LocalToModelJob: a joint's model matrix needs its parent's, so joints are processed by
    // depth; within a depth every (joint, column) pair is an independent task, 32 tasks per round (one per lane)
    std::vector<int> depth(J, 0);
    int max_depth = 0;
    for (int j = 0; j < J; ++j) {
        depth[j] = skel->data.parents[j] < 0 ? 0 : depth[skel->data.parents[j]] + 1;
        max_depth = std::max(max_depth, depth[j]);
    }
    // Which 8 joints of a depth share a round is free: lane (joint j, column c) reads and writes shared memory at 13 j + 3 c + r, so two
    // joints of a round collide in a bank iff 13 (j - j') = 3 (c' - c) (mod 32), i.e. j - j' in {0, 2, 13, 15, 17, 19, 30} (mod 32);
    // their parents' matrices (13 p + e) collide iff p = p' (mod 32) for different parents.  Greedy: rounds without such pairs first.
    auto clash = [&](int a, int b) {
        const int d = ((a - b) % 32 + 32) % 32;
        if (d == 0 || d == 2 || d == 13 || d == 15 || d == 17 || d == 19 || d == 30) return true;
        const int pa = skel->data.parents[a], pb = skel->data.parents[b];
        return pa >= 0 && pb >= 0 && pa != pb && pa % 32 == pb % 32;
    };
    std::vector<uint16_t> sched;
    for (int d = 0; d <= max_depth && J > 0; ++d) {
        std::vector<int> todo;
        for (int j = 0; j < J; ++j)
            if (depth[j] == d) todo.push_back(j);
        while (!todo.empty()) {
            std::vector<int> round, rest;
            for (int j : todo) {
                bool ok = round.size() < 8;
                for (size_t i = 0; ok && i < round.size(); ++i) ok = !clash(j, round[i]);
                (ok ? round : rest).push_back(j);
            }
            // joints that clash with everything chosen so far still have to run: when a whole round's worth is left over, or at the
            // end of the depth, they fill the round up (as many rounds as before: ceil(joints of the depth / 8))
            const size_t rounds_left = (rest.size() + round.size() + 7) / 8;
            if (round.size() < 8 && rounds_left * 8 - 8 < rest.size()) {
                while (round.size() < 8 && !rest.empty()) {
                    round.push_back(rest.front());
                    rest.erase(rest.begin());
                }
            }
            size_t in_round = 0;
            for (int j : round)
                for (int c = 0; c < 4; ++c) {
                    sched.push_back(static_cast<uint16_t>(j | (c << 12)));
                    ++in_round;
                }
            while (in_round % 32) {
                sched.push_back(kIdleTask);
                ++in_round;
            }
            todo.swap(rest);
        }
    }
    skel->rounds = static_cast<int>(sched.size() / 32);
    DMK_CUDA(ctx, upload(ctx, &skel->d_parents, skel->data.parents));
    DMK_CUDA(ctx, upload(ctx, &skel->d_schedule, sched));
    {
        const int P = skel->data.num_soa() * 4;
        std::vector<float> rest(static_cast<size_t>(P) * 10);
        for (int j = 0; j < P; ++j)
            for (int k = 0; k < 10; ++k) rest[static_cast<size_t>(j) * 10 + k] = skel->data.rest_soa[static_cast<size_t>(j / 4) * 40 + k * 4 + (j & 3)];
        DMK_CUDA(ctx, upload(ctx, &skel->d_rest, rest));
    }
    *out = skel.release();
    return DMK_OK;
}

void dmk_skeleton_destroy(dmk_skeleton_t skel) {
    if (!skel) return;
    cudaSetDevice(skel->ctx->device);
    dev_free(skel->d_parents);
    dev_free(skel->d_schedule);
    dev_free(skel->d_rest);
    delete skel;
}
int dmk_skeleton_num_joints(dmk_skeleton_t skel) { return skel ? skel->data.num_joints : 0; }
const char* dmk_skeleton_joint_name(dmk_skeleton_t skel, int i) {
    return (skel && i >= 0 && i < skel->data.num_joints) ? skel->data.names[i].c_str() : nullptr;
}
int dmk_skeleton_joint_parent(dmk_skeleton_t skel, int i) {
    return (skel && i >= 0 && i < skel->data.num_joints) ? skel->data.parents[i] : -1;
}
int dmk_skeleton_find_joint(dmk_skeleton_t skel, const char* name) {
    if (!skel || !name) return -1;
    for (int i = 0; i < skel->data.num_joints; ++i)
        if (skel->data.names[i] == name) return i;
    return -1;
}

// ------------------------------------------------------------------------------------------------------------
// clips
// ------------------------------------------------------------------------------------------------------------
dmk_status dmk_clip_load_ozz(dmk_context_t ctx, const void* bytes, size_t size, dmk_clip_t* out) {
    if (!ctx || !out) return fail(ctx, DMK_ERR_INVALID, "null argument");
    *out = nullptr;
    auto clip = std::make_unique<dmk_clip_s>();
    clip->ctx = ctx;
    std::string err;
    if (!parse_animation(bytes, size, clip->data, err)) return fail(ctx, DMK_ERR_FORMAT, err);
    *out = clip.release();
    return DMK_OK;
}
void dmk_clip_destroy(dmk_clip_t clip) { delete clip; }
float dmk_clip_duration(dmk_clip_t clip) { return clip ? clip->data.duration : 0.f; }
const char* dmk_clip_name(dmk_clip_t clip) { return clip ? clip->data.name.c_str() : nullptr; }
int dmk_clip_num_tracks(dmk_clip_t clip) { return clip ? clip->data.num_tracks : 0; }


dmk_status dmk_clipset_create(dmk_context_t ctx, dmk_skeleton_t skel, const dmk_clip_t* clips, int num_clips, dmk_clipset_t* out) {
    if (!ctx || !skel || !clips || !out || num_clips <= 0) return fail(ctx, DMK_ERR_INVALID, "null argument or empty clip list");
    *out = nullptr;
    if (bind(ctx)) return DMK_ERR_CUDA;
    const int padded = skel->data.num_soa() * 4;
    Building<dmk_clipset_s, dmk_clipset_destroy> set(new dmk_clipset_s());
    set->ctx = ctx;
    set->num_clips = num_clips;
    set->padded_tracks = padded;
    std::vector<uint2> dir(static_cast<size_t>(num_clips) * 3 * padded);
    PackedComponent pc[3];
    for (int c = 0; c < num_clips; ++c) {
        if (!clips[c]) return fail(ctx, DMK_ERR_INVALID, "null clip");
        const AnimationData& a = clips[c]->data;
        if (a.num_soa() != skel->data.num_soa())
            return fail(ctx, DMK_ERR_UNSUPPORTED,
                        "animation '" + a.name + "' has " + std::to_string(a.num_tracks) + " tracks, skeleton has " +
                            std::to_string(skel->data.num_joints) + " joints (SoA track counts must match)");
        std::string err;
        uint2* d = dir.data() + static_cast<size_t>(c) * 3 * padded;
        if (!regroup(a.translations, padded, pc[0], d, err) || !regroup(a.rotations, padded, pc[1], d + padded, err) ||
            !regroup(a.scales, padded, pc[2], d + 2 * padded, err))
            return fail(ctx, DMK_ERR_FORMAT, err);
        set->durations.push_back(a.duration);
    }
    // bucket tables need every clip's keys regrouped first: rebuild the per-clip views over the packed arrays
    BucketTable tbl;
    bool buckets = true;
    for (int c = 0; c < num_clips && buckets; ++c)
        buckets = build_buckets(pc, dir.data() + static_cast<size_t>(c) * 3 * padded, padded, tbl);
    if (buckets) {
        DMK_CUDA(ctx, upload(ctx, &set->d_bucket_dir, tbl.dir));
        DMK_CUDA(ctx, upload(ctx, &set->d_bucket_a, tbl.a));
        DMK_CUDA(ctx, upload(ctx, &set->d_bucket_b, tbl.b));
        DMK_CUDA(ctx, upload(ctx, &set->d_bucket_c, tbl.c));
        set->has_buckets = true;
    }
    DMK_CUDA(ctx, upload(ctx, &set->d_dir, dir));
    for (int k = 0; k < 3; ++k) {
        DMK_CUDA(ctx, upload(ctx, &set->d_ratio[k], pc[k].ratio));
        DMK_CUDA(ctx, upload(ctx, &set->d_value[k], pc[k].value));
        DMK_CUDA(ctx, upload(ctx, &set->d_stream[k], pc[k].stream));
    }
    DMK_CUDA(ctx, upload(ctx, &set->d_duration, set->durations));
    *out = set.release();
    return DMK_OK;
}

void dmk_clipset_destroy(dmk_clipset_t set) {
    if (!set) return;
    cudaSetDevice(set->ctx->device);
    dev_free(set->d_dir);
    for (int k = 0; k < 3; ++k) {
        dev_free(set->d_ratio[k]);
        dev_free(set->d_value[k]);
        dev_free(set->d_stream[k]);
    }
    dev_free(set->d_duration);
    dev_free(set->d_bucket_dir);
    dev_free(set->d_bucket_a);
    dev_free(set->d_bucket_b);
    dev_free(set->d_bucket_c);
    delete set;
}

// ------------------------------------------------------------------------------------------------------------
// armature + mesh
// ------------------------------------------------------------------------------------------------------------
dmk_status dmk_armature_create(dmk_context_t ctx, dmk_skeleton_t skel, const char* const* joint_names, const float* inv_bind,
                               int num_joints, dmk_armature_t* out) {
    if (!ctx || !skel || !out || num_joints < 0 || (num_joints > 0 && (!joint_names || !inv_bind)))
        return fail(ctx, DMK_ERR_INVALID, "null argument");
    *out = nullptr;
    if (num_joints + 1 > kMaxPalette)
        return fail(ctx, DMK_ERR_UNSUPPORTED, "armature has more than 255 joints (palette slots are bytes; the reference caps at 64)");
    if (bind(ctx)) return DMK_ERR_CUDA;
    Building<dmk_armature_s, dmk_armature_destroy> arm(new dmk_armature_s());
    arm->ctx = ctx;
    arm->num_joints = num_joints;
    std::vector<float> ib(static_cast<size_t>(num_joints) * 12);
    for (int k = 0; k < num_joints; ++k) {
        const float* m = inv_bind + static_cast<size_t>(k) * 16;
        if (m[3] != 0.f || m[7] != 0.f || m[11] != 0.f || m[15] != 1.f)
            return fail(ctx, DMK_ERR_UNSUPPORTED, "inverse bind pose of joint " + std::to_string(k) + " is not affine (bottom row must be 0,0,0,1)");
        for (int c = 0; c < 4; ++c)
            for (int r = 0; r < 3; ++r) ib[static_cast<size_t>(k) * 12 + c * 3 + r] = m[c * 4 + r];
        arm->remap.push_back(joint_names[k] ? dmk_skeleton_find_joint(skel, joint_names[k]) : -1);
    }
    // Processing order of the palette pass (ArmatureDev): groups of 32 armature joints whose skeleton joints differ mod 32, as far as the
    // armature allows -- a warp reads 32 model matrices at once, 13 floats apart per joint, and equal residues share a bank.
    std::vector<int> order;
    {
        std::vector<char> taken(static_cast<size_t>(num_joints), 0);
        int left = num_joints;
        while (left > 0) {
            bool used[32] = {};
            int in_group = 0;
            for (int pass = 0; pass < 2 && in_group < 32; ++pass)            // pass 0: distinct residues only; pass 1: fill up
                for (int k = 0; k < num_joints && in_group < 32; ++k) {
                    if (taken[static_cast<size_t>(k)]) continue;
                    const int jm = arm->remap[static_cast<size_t>(k)];
                    const int res = jm < 0 ? -1 : jm % 32;                   // unknown names read the identity: no bank of the model array
                    if (pass == 0 && res >= 0 && used[res]) continue;
                    if (res >= 0) used[res] = true;
                    taken[static_cast<size_t>(k)] = 1;
                    order.push_back(k);
                    ++in_group;
                    --left;
                }
        }
    }
    std::vector<int32_t> remap_ordered(static_cast<size_t>(num_joints));
    std::vector<float> ib_ordered(ib.size());
    for (int i = 0; i < num_joints; ++i) {
        const int k = order[static_cast<size_t>(i)];
        remap_ordered[static_cast<size_t>(i)] = (arm->remap[static_cast<size_t>(k)] & 0xffff) | (k << 16);
        std::copy(ib.begin() + static_cast<long>(k) * 12, ib.begin() + static_cast<long>(k) * 12 + 12, ib_ordered.begin() + static_cast<long>(i) * 12);
    }
    DMK_CUDA(ctx, upload(ctx, &arm->d_remap, remap_ordered));
    DMK_CUDA(ctx, upload(ctx, &arm->d_inv_bind, ib_ordered));
    *out = arm.release();
    return DMK_OK;
}
void dmk_armature_destroy(dmk_armature_t arm) {
    if (!arm) return;
    cudaSetDevice(arm->ctx->device);
    dev_free(arm->d_remap);
    dev_free(arm->d_inv_bind);
    delete arm;
}
int dmk_armature_palette_size(dmk_armature_t arm) { return arm ? arm->num_joints + 1 : 0; }

dmk_status dmk_mesh_create(dmk_context_t ctx, const float* position, const float* normal, const float* tangent, const float* indices,
                           const float* weights, int num_vertices, int palette_size, dmk_mesh_t* out) {
    return dmk_mesh_create_ex(ctx, position, normal, tangent, indices, weights, num_vertices, palette_size, 0, out);
}

dmk_status dmk_mesh_vertex_order(dmk_mesh_t mesh, int32_t* out_order) {
    if (!mesh || !out_order) return DMK_ERR_INVALID;
    std::copy(mesh->order.begin(), mesh->order.end(), out_order);
    return DMK_OK;
}

// (slot, weight) pairs per caller vertex: getSkinningMatrixPart reads u_skinning[int(index + 1)]
// (assets/shaders/core/darmok_skinning.inc.slang:5-14); applySkinning passes vertices with indices.x < 0 through
// (= 1 * identity, slot 0).
static bool vertex_influences(const float* indices, const float* weights, int num_vertices, int palette_size, std::vector<ClusterInfluence>& infl,
                              std::string& err) {
    infl.resize(static_cast<size_t>(num_vertices));
    for (int v = 0; v < num_vertices; ++v) {
        ClusterInfluence& x = infl[static_cast<size_t>(v)];
        x = ClusterInfluence{{0, 0, 0, 0}, {1.f, 0.f, 0.f, 0.f}};
        const float* idx = indices + static_cast<size_t>(v) * 4;
        if (!(idx[0] >= 0.f)) continue;
        for (int c = 0; c < 4; ++c) {
            const int slot = static_cast<int>(idx[c] + 1.f);
            if (!(idx[c] >= -1.f) || slot < 0 || slot >= palette_size) {
                err = "vertex " + std::to_string(v) + " references bone " + std::to_string(idx[c]) + " outside the palette (size " +
                      std::to_string(palette_size) + ")";
                return false;
            }
            x.slot[c] = static_cast<uint8_t>(slot);
            x.weight[c] = weights[static_cast<size_t>(v) * 4 + c];
        }
    }
    return true;
}

dmk_status dmk_mesh_cluster_stats(const float* indices, const float* weights, int num_vertices, int palette_size, int32_t* out_order,
                                  int64_t* out_stats4) {
    if (num_vertices < 0 || (num_vertices > 0 && (!indices || !weights)) || !out_stats4 || palette_size < 1 || palette_size > kMaxPalette)
        return fail(nullptr, DMK_ERR_INVALID, "dmk_mesh_cluster_stats: argument out of range");
    std::vector<ClusterInfluence> infl;
    std::string err;
    if (!vertex_influences(indices, weights, num_vertices, palette_size, infl, err)) return fail(nullptr, DMK_ERR_INVALID, err);
    int vpad = (num_vertices + 7) / 8 * 8;
    SkinPlan plan{};
    if (vpad > 0) {
        const int vpad64 = (num_vertices + 63) / 64 * 64;
        plan = plan_skin(vpad64, palette_size, 148, 0);   // the bulk-store output path wants a pitch that is a multiple of 64
        if (plan.strided_store) vpad = vpad64;
        else plan = plan_skin(vpad, palette_size, 148, 0);
        cudaGetLastError();   // without a device the occupancy query fails; the tile shape does not depend on it
    }
    const ClusterResult cl = plan.strided_store ? cluster_mesh_strided(infl, vpad, plan.num_tiles, plan.tile_vertices)
                                                : cluster_mesh(infl, plan.num_tiles, plan.tile_vertices, plan.ws_vpt == 4 ? 16 : 32);
    if (out_order) std::copy(cl.order.begin(), cl.order.begin() + num_vertices, out_order);
    out_stats4[0] = cl.gather_wavefronts;
    out_stats4[1] = cl.gather_wavefronts_unclustered;
    out_stats4[2] = cl.shared_blocks;
    out_stats4[3] = cl.total_blocks;
    return DMK_OK;
}

dmk_status dmk_mesh_cluster_layout(int num_vertices, int palette_size, int64_t* out_layout3) {
    if (num_vertices < 0 || !out_layout3 || palette_size < 1 || palette_size > kMaxPalette)
        return fail(nullptr, DMK_ERR_INVALID, "dmk_mesh_cluster_layout: argument out of range");
    int vpad = (num_vertices + 7) / 8 * 8;
    SkinPlan plan{};
    if (vpad > 0) {
        const int vpad64 = (num_vertices + 63) / 64 * 64;
        plan = plan_skin(vpad64, palette_size, 148, 0);
        if (plan.strided_store) vpad = vpad64;
        else plan = plan_skin(vpad, palette_size, 148, 0);
        cudaGetLastError();
    }
    out_layout3[0] = vpad;
    out_layout3[1] = plan.strided_store ? 1 : 0;
    out_layout3[2] = plan.tile_vertices;
    return DMK_OK;
}

dmk_status dmk_mesh_create_ex(dmk_context_t ctx, const float* position, const float* normal, const float* tangent, const float* indices,
                              const float* weights, int num_vertices, int palette_size, uint32_t flags, dmk_mesh_t* out) {
    if (!ctx || !out || num_vertices < 0 || (num_vertices > 0 && (!position || !normal || !tangent || !indices || !weights)))
        return fail(ctx, DMK_ERR_INVALID, "null argument");
    *out = nullptr;
    if (palette_size < 1 || palette_size > kMaxPalette) return fail(ctx, DMK_ERR_INVALID, "palette_size must be in [1, 256]");
    if (bind(ctx)) return DMK_ERR_CUDA;
    Building<dmk_mesh_s, dmk_mesh_destroy> mesh(new dmk_mesh_s());
    mesh->ctx = ctx;
    mesh->num_vertices = num_vertices;
    mesh->palette_size = palette_size;
    int vpad = (num_vertices + 7) / 8 * 8;
    // clustered by bone: the vertex pitch becomes a multiple of 64 when the skinning plan offers the bulk-store output path for it
    SkinPlan cluster_plan{};
    if ((flags & DMK_MESH_CLUSTER_BONES) && num_vertices > 0) {
        const int vpad64 = (num_vertices + 63) / 64 * 64;
        cluster_plan = plan_skin(vpad64, palette_size, ctx->num_sms, 0);
        if (cluster_plan.strided_store) vpad = vpad64;
        else cluster_plan = plan_skin(vpad, palette_size, ctx->num_sms, 0);
    }
    mesh->vpad = vpad;
    std::vector<float> pos(static_cast<size_t>(vpad) * 3, 0.f), nrm(pos.size(), 0.f), tan(pos.size(), 0.f);
    std::vector<float> w(static_cast<size_t>(vpad) * 4, 0.f);
    std::vector<uint32_t> slots(vpad, 0u);
    // storage order: the caller's, or grouped by the number of non-zero weights (stable, most influences first)
    mesh->order.resize(static_cast<size_t>(num_vertices));
    for (int v = 0; v < num_vertices; ++v) mesh->order[static_cast<size_t>(v)] = v;
    if (flags & DMK_MESH_SORT_BY_INFLUENCES) {
        auto count = [&](int v) {
            const float* idx = indices + static_cast<size_t>(v) * 4;
            if (!(idx[0] >= 0.f)) return 1;
            int n = 1;
            for (int c = 1; c < 4; ++c) n += weights[static_cast<size_t>(v) * 4 + c] != 0.f ? 1 : 0;
            return n;
        };
        std::stable_sort(mesh->order.begin(), mesh->order.end(), [&](int a, int b) { return count(a) > count(b); });
        // Every thread of the skinning kernel owns 8 consecutive stored vertices and a warp executes the k-th vertex of its
        // 32 threads together.  Fill the storage "column major" over those 8-vertex blocks: sorted vertex i goes to block
        // i % B, slot i / B.  Then slot k of (almost) every block holds vertices with the same influence count -- the
        // gathers of an absent influence are skipped by whole warps -- and every thread, warp and tile gets the same mix.
        // (4-vertex blocks for the 4-vertices-per-thread shapes, DMK_MESH_SORT_BLOCKS_OF_4.)
        const size_t per = (flags & DMK_MESH_SORT_BLOCKS_OF_4) ? 4 : 8;
        const size_t V = mesh->order.size(), B = static_cast<size_t>(vpad) / per;
        std::vector<int32_t> dealt(V);
        size_t next = 0;
        for (size_t slot = 0; slot < per; ++slot)
            for (size_t blk = 0; blk < B; ++blk) {
                const size_t pos = blk * per + slot;
                if (pos < V) dealt[pos] = mesh->order[next++];
            }
        mesh->order = dealt;
        mesh->sorted = true;
        mesh->sort_block = static_cast<int>(per);
    }
    std::vector<ClusterInfluence> infl;
    {
        std::string err;
        if (!vertex_influences(indices, weights, num_vertices, palette_size, infl, err)) return fail(ctx, DMK_ERR_INVALID, err);
    }
    std::vector<ClusterInfluence> stored;
    if (flags & DMK_MESH_CLUSTER_BONES) {
        const SkinPlan& plan = cluster_plan;
        ClusterResult cl = plan.strided_store ? cluster_mesh_strided(infl, vpad, plan.num_tiles, plan.tile_vertices)
                                              : cluster_mesh(infl, plan.num_tiles, plan.tile_vertices, plan.ws_vpt == 4 ? 16 : 32);
        mesh->strided = plan.strided_store;
        mesh->order.assign(cl.order.begin(), cl.order.begin() + num_vertices);   // padding sits behind the last real vertex
        stored = std::move(cl.stored);
        mesh->clustered = true;
        mesh->sorted = true;   // absent influences carry weight 0 behind the present ones: the gathers are predicated
        mesh->sort_block = 8;
        mesh->cluster_gather_wavefronts = cl.gather_wavefronts;
        mesh->cluster_shared_blocks = cl.shared_blocks;
    }
    for (int s = 0; s < vpad; ++s) {
        float* wv = &w[static_cast<size_t>(s) * 4];
        wv[0] = 1.f;  // padding: 1 * identity
        if (s >= num_vertices) {
            if (mesh->strided) {   // padding inside a short block: 0 * the block's bone (no gather, no mixed-block path; output 0 either way)
                wv[0] = 0.f;
                slots[s] = stored[static_cast<size_t>(s)].slot[0];
            }
            continue;
        }
        const int v = mesh->order[static_cast<size_t>(s)];
        for (int k = 0; k < 3; ++k) {
            pos[static_cast<size_t>(s) * 3 + k] = position[static_cast<size_t>(v) * 3 + k];
            nrm[static_cast<size_t>(s) * 3 + k] = normal[static_cast<size_t>(v) * 3 + k];
            tan[static_cast<size_t>(s) * 3 + k] = tangent[static_cast<size_t>(v) * 3 + k];
        }
        const ClusterInfluence& x = mesh->clustered ? stored[static_cast<size_t>(s)] : infl[static_cast<size_t>(v)];
        uint32_t packed = 0;
        for (int c = 0; c < 4; ++c) {
            packed |= static_cast<uint32_t>(x.slot[c]) << (8 * c);
            wv[c] = x.weight[c];
        }
        slots[s] = packed;
    }
    DMK_CUDA(ctx, upload(ctx, &mesh->d_pos, pos));
    DMK_CUDA(ctx, upload(ctx, &mesh->d_nrm, nrm));
    DMK_CUDA(ctx, upload(ctx, &mesh->d_tan, tan));
    DMK_CUDA(ctx, upload(ctx, &mesh->d_weights, w));
    DMK_CUDA(ctx, upload(ctx, &mesh->d_slots, slots));
    *out = mesh.release();
    return DMK_OK;
}
void dmk_mesh_destroy(dmk_mesh_t mesh) {
    if (!mesh) return;
    cudaSetDevice(mesh->ctx->device);
    dev_free(mesh->d_pos);
    dev_free(mesh->d_nrm);
    dev_free(mesh->d_tan);
    dev_free(mesh->d_weights);
    dev_free(mesh->d_slots);
    delete mesh;
}
int dmk_mesh_num_vertices(dmk_mesh_t mesh) { return mesh ? mesh->num_vertices : 0; }
int dmk_mesh_vertex_pitch(dmk_mesh_t mesh) { return mesh ? mesh->vpad : 0; }

}  // extern "C"
