/*
 * oracle/crowd.c -- TEST INFRASTRUCTURE (see dmk_oracle.h).
 *
 * The reference's per-character frame looped over a crowd, used (a) as the checker of
 * the batched CUDA step and (b) as the CPU baseline of bench.py: OpenMP `parallel for` over characters.
 *   per clip:   SamplingJob with a persistent Context       src/skeleton_ozz.cpp:429-437
 *   per state:  BlendingJob over the layers (if > 1)         src/skeleton_ozz.cpp:552-608 / :709-724
 *   per char:   LocalToModelJob                              src/skeleton_ozz.cpp:1034-1039
 *   per mesh:   palette (src/skeleton.cpp:236-249) + LBS (darmok_skinning.inc.slang)
 */
#include "dmk_oracle.h"

#include <stdlib.h>
#include <string.h>

struct orc_crowd {
    const orc_skeleton* skel;
    const orc_animation* const* clips;
    int num_clips;
    const int32_t* remap;
    const float* inv_bind;
    int k;
    int64_t n;
    int num_layers;
    orc_sampling_context** ctx; /* [layer][character], as one Context per clip state in the reference */
};

orc_crowd* orc_crowd_create(const orc_skeleton* skel, const orc_animation* const* clips, int num_clips,
                            const int32_t* remap, const float* inv_bind, int k, int64_t n, int num_layers) {
    orc_crowd* c = (orc_crowd*)calloc(1, sizeof *c);
    c->skel = skel;
    c->clips = clips;
    c->num_clips = num_clips;
    c->remap = remap;
    c->inv_bind = inv_bind;
    c->k = k;
    c->n = n;
    c->num_layers = num_layers;
    c->ctx = (orc_sampling_context**)calloc((size_t)(n * num_layers), sizeof(void*));
    for (int64_t i = 0; i < n * num_layers; ++i) c->ctx[i] = orc_sampling_context_create(skel->num_joints);
    return c;
}

void orc_crowd_free(orc_crowd* c) {
    if (!c) return;
    for (int64_t i = 0; i < c->n * c->num_layers; ++i) orc_sampling_context_free(c->ctx[i]);
    free(c->ctx);
    free(c);
}

int64_t orc_crowd_step(orc_crowd* c, const int32_t* clip, const float* ratio, const float* weight,
                       float threshold, float* out_locals, float* out_models, float* out_palette,
                       const float* pos, const float* nrm, const float* tan, const float* indices,
                       const float* weights, int num_vertices, float* out_skinned, int num_threads) {
    const int J = c->skel->num_joints, S = c->skel->num_soa, L = c->num_layers, K = c->k;
    const int64_t n = c->n;
    int64_t first_error = 0;
    if (num_threads < 1) num_threads = 1;
#pragma omp parallel num_threads(num_threads)
    {
        float* layer_buf = (float*)malloc(sizeof(float) * ORC_SOA_FLOATS * (size_t)S * (size_t)(L + 1));
        float* models = (float*)malloc(sizeof(float) * 16 * (size_t)J);
        float* palette = (float*)malloc(sizeof(float) * 16 * (size_t)(K + 1));
        /* pos != NULL with out_skinned == NULL: skin into a per-thread scratch (CPU baseline timing: the result
         * stays cache resident, the most favourable case for the CPU) */
        const int scratch_skin = (!out_skinned && pos && num_vertices > 0);
        float* scratch = scratch_skin ? (float*)malloc(sizeof(float) * 9 * (size_t)num_vertices) : NULL;
        const float** layer_ptr = (const float**)malloc(sizeof(float*) * (size_t)L);
        float* wts = (float*)malloc(sizeof(float) * (size_t)L);
#pragma omp for schedule(static)
        for (int64_t i = 0; i < n; ++i) {
            int ok = 1;
            for (int l = 0; l < L && ok; ++l) {
                const int32_t ci = clip[(int64_t)l * n + i];
                float* dst = layer_buf + (size_t)l * S * ORC_SOA_FLOATS;
                layer_ptr[l] = dst;
                wts[l] = weight[(int64_t)l * n + i];
                if (ci < 0 || ci >= c->num_clips) { ok = 0; break; }
                ok = orc_sample(c->clips[ci], c->ctx[(int64_t)l * n + i], ratio[(int64_t)l * n + i], dst, S, NULL);
            }
            const float* locals = layer_buf;
            if (ok && L > 1) {
                /* transition form: rest pose = layer 0 (src/skeleton_ozz.cpp:719) */
                float* blended = layer_buf + (size_t)L * S * ORC_SOA_FLOATS;
                ok = orc_blend(L, layer_ptr, wts, layer_ptr[0], threshold, S, blended);
                locals = blended;
            }
            if (ok) ok = orc_local_to_model(c->skel, locals, models);
            if (!ok) {
#pragma omp critical
                if (first_error == 0 || i + 1 < first_error) first_error = i + 1;
                continue;
            }
            if (out_locals) memcpy(out_locals + (size_t)i * S * ORC_SOA_FLOATS, locals, sizeof(float) * ORC_SOA_FLOATS * (size_t)S);
            if (out_models) memcpy(out_models + (size_t)i * J * 16, models, sizeof(float) * 16 * (size_t)J);
            if (out_palette || scratch_skin || (out_skinned && num_vertices > 0)) {
                orc_palette(models, J, c->remap, c->inv_bind, K, palette);
                if (out_palette) memcpy(out_palette + (size_t)i * (K + 1) * 16, palette, sizeof(float) * 16 * (size_t)(K + 1));
                if (out_skinned && num_vertices > 0)
                    orc_skin(palette, K + 1, pos, nrm, tan, indices, weights, num_vertices,
                             out_skinned + (size_t)i * (size_t)num_vertices * 9);
                else if (scratch_skin)
                    orc_skin(palette, K + 1, pos, nrm, tan, indices, weights, num_vertices, scratch);
            }
        }
        free(scratch);
        free(layer_buf);
        free(models);
        free(palette);
        free(layer_ptr);
        free(wts);
    }
    return first_error;
}
