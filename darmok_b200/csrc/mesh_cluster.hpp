// mesh_cluster.hpp -- host-side vertex ordering for the skinning kernel (DMK_MESH_CLUSTER_BONES).
//
// The skinning kernel (skin_kernel.cu) is bound by shared-memory wavefronts, not by HBM: every vertex gathers 4 palette
// matrices x 3 rows of 16 bytes (192 B) against 36 B of output.  Two properties of the STORED vertex order cut that:
//   * a thread owns 8 consecutive stored vertices ("block").  When all 8 carry a common bone, the thread loads that
//     bone's 3 rows ONCE per character into registers and the 8 per-vertex gathers of it disappear;
//   * a 128-bit shared-memory load is served a quarter-warp (8 lanes) at a time, so a predicated-off gather only saves
//     its wavefront when the 8 lanes of a quarter-warp agree.  Blocks are therefore laid out so that the k-th vertices of
//     the 8 blocks of a quarter-warp have the same number of influences.
// This file computes such an order for any mesh (it is a permutation of the caller's vertices plus a reordering of every
// vertex's 4 (bone, weight) pairs -- both leave the skinned result unchanged up to the order of a 4-term float sum).
// The caller gets the permutation through dmk_mesh_vertex_order and remaps its index buffer once, as after any
// vertex-cache optimisation.  Reference for the per-vertex input contract: MeshData::exportData src/mesh_core.cpp:336-356.
#pragma once

#include <cstdint>
#include <vector>

namespace dmk {

struct ClusterInfluence {
    uint8_t slot[4];   // palette slots (bone index + 1; 0 = identity)
    float weight[4];
};

struct ClusterResult {
    std::vector<int32_t> order;              // [vpad]: caller's vertex index stored at position i, -1 = padding vertex
    std::vector<ClusterInfluence> stored;    // [vpad]: influences of the stored vertex, entry 0 = the block's bone where it has it
    // what the ordering buys, in 128-bit shared-memory load wavefronts per character (quarter-warp granularity):
    int64_t gather_wavefronts = 0;           // with this order: block bone once per thread + predicated gathers
    int64_t gather_wavefronts_unclustered = 0;   // 4 influences x 3 rows for every vertex (the caller-order kernel)
    int32_t shared_blocks = 0;               // 8-vertex blocks whose vertices all carry the block's bone
    int32_t total_blocks = 0;
};

// `in[v]`: the 4 (slot, weight) pairs of caller vertex v.  `tile_vertices` (multiple of 8) and `num_tiles`: the tile shape
// of the skinning plan the order is dealt for (every tile, warp and quarter-warp gets the same mix of heavy and light
// blocks); any other plan stays correct, only less balanced.  `blocks_per_warp`: 32 (a thread owns a block) or 16 (two lanes
// share a block: the 4-vertices-per-thread shape); the gather statistics count the extra block-bone reads of the latter.
ClusterResult cluster_mesh(const std::vector<ClusterInfluence>& in, int num_tiles, int tile_vertices, int blocks_per_warp = 32);

// The same for the STRIDED ownership of the skinning kernel's bulk-store output path: thread (warp w, lane l) of a tile owns the stored
// vertices tile_begin + 256 w + k nl + l (k = 0..7; nl = 32, or the lanes in use in the tile's last warp), so that round k of a warp
// is nl consecutive vertices.  `vpad` (vertex pitch) and `tile_vertices` are multiples of 64; positions >= in.size() are padding
// (order = -1; they name their block's bone with weight 0: no gather, output 0).
ClusterResult cluster_mesh_strided(const std::vector<ClusterInfluence>& in, int vpad, int num_tiles, int tile_vertices);

}  // namespace dmk
