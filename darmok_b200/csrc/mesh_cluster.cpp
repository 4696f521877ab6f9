// mesh_cluster.cpp -- see mesh_cluster.hpp.  Host code only (no CUDA), deterministic for a given input.
#include "mesh_cluster.hpp"

#include <algorithm>
#include <array>
#include <cstring>

namespace dmk {

namespace {

constexpr int kBlock = 8;      // vertices per thread of the skinning kernel
constexpr int kQuarter = 8;    // lanes served by one 128-bit shared-memory wavefront

struct Vtx {
    int32_t id = -1;           // caller's index, -1 = padding
    int n = 1;                 // influences with a non-zero weight (at least 1: entry 0 is always evaluated)
    int home = 0;              // entry that names the bone this vertex is grouped under
    uint8_t slot[4] = {0, 0, 0, 0};
    float weight[4] = {1.f, 0.f, 0.f, 0.f};
};

struct Block {
    std::array<int, kBlock> v{};   // indices into the vertex table, position 0 carries the block's bone
    int bone = 0;
    std::array<uint8_t, kBlock> gathers{};   // palette matrices vertex k still fetches from shared memory
};

// entries with a non-zero weight first, in the caller's order (a NaN weight counts as non-zero: it must reach the sum)
Vtx normalise(const ClusterInfluence& in, int32_t id) {
    Vtx x;
    x.id = id;
    int n = 0;
    for (int c = 0; c < 4; ++c)
        if (in.weight[c] != 0.f) {
            x.slot[n] = in.slot[c];
            x.weight[n] = in.weight[c];
            ++n;
        }
    if (n == 0) {   // all weights zero: the blended matrix is 0 * palette[slot 0]
        x.slot[0] = in.slot[0];
        x.weight[0] = in.weight[0];
        n = 1;
    }
    for (int c = n; c < 4; ++c) {
        x.slot[c] = 0;
        x.weight[c] = 0.f;
    }
    x.n = n;
    return x;
}

}  // namespace

ClusterResult cluster_mesh(const std::vector<ClusterInfluence>& in, int num_tiles, int tile_vertices, int blocks_per_warp) {
    ClusterResult out;
    const int V = static_cast<int>(in.size());
    const int vpad = (V + kBlock - 1) / kBlock * kBlock;
    const int nb = vpad / kBlock;
    out.total_blocks = nb;
    out.order.assign(static_cast<size_t>(vpad), -1);
    out.stored.resize(static_cast<size_t>(vpad));
    if (vpad == 0) return out;
    if (num_tiles < 1) num_tiles = 1;
    if (tile_vertices < kBlock) tile_vertices = kBlock;
    if (blocks_per_warp != 16) blocks_per_warp = 32;
    const int quarters = blocks_per_warp / kQuarter;   // quarter-warp groups of 8 blocks per warp

    std::vector<Vtx> vt(static_cast<size_t>(vpad));
    for (int v = 0; v < V; ++v) vt[static_cast<size_t>(v)] = normalise(in[static_cast<size_t>(v)], v);
    // padding vertices keep the defaults: 1 * identity

    // ---- 1. a home bone per vertex, chosen so that as many bones as possible hold a multiple of 8 vertices ----
    std::array<int, 256> cnt{};
    for (int v = 0; v < V; ++v) ++cnt[vt[static_cast<size_t>(v)].slot[0]];
    for (int pass = 0; pass < 64; ++pass) {
        bool moved = false;
        for (int v = 0; v < V; ++v) {
            Vtx& x = vt[static_cast<size_t>(v)];
            if (x.n < 2) continue;
            const int b = x.slot[x.home];
            const int r = cnt[b] % kBlock;
            if (r == 0) continue;
            // move to the bone that is closest to completing a block (never to one that is further from it than this one)
            int best = -1, best_r = r - 1;
            for (int e = 0; e < x.n; ++e) {
                const int b2 = x.slot[e];
                if (b2 == b) continue;
                const int r2 = cnt[b2] % kBlock;
                if (r2 > best_r) {
                    best_r = r2;
                    best = e;
                }
            }
            if (best < 0) continue;
            --cnt[b];
            ++cnt[x.slot[best]];
            x.home = best;
            moved = true;
        }
        if (!moved) break;
    }
    // entry 0 = the home bone
    for (int v = 0; v < V; ++v) {
        Vtx& x = vt[static_cast<size_t>(v)];
        if (x.home != 0) {
            std::swap(x.slot[0], x.slot[x.home]);
            std::swap(x.weight[0], x.weight[x.home]);
            x.home = 0;
        }
    }

    // ---- 2. blocks: per bone, vertices by influence count; what does not fill a block goes to mixed blocks ----
    std::vector<std::vector<int>> members(256);
    for (int v = 0; v < V; ++v) members[vt[static_cast<size_t>(v)].slot[0]].push_back(v);
    std::vector<Block> blocks;
    blocks.reserve(static_cast<size_t>(nb));
    std::vector<int> rest;
    for (int b = 0; b < 256; ++b) {
        auto& m = members[static_cast<size_t>(b)];
        std::stable_sort(m.begin(), m.end(), [&](int a, int c) { return vt[static_cast<size_t>(a)].n > vt[static_cast<size_t>(c)].n; });
        const size_t full = m.size() / kBlock * kBlock;
        for (size_t i = 0; i < full; i += kBlock) {
            Block blk;
            blk.bone = b;
            for (int k = 0; k < kBlock; ++k) {
                blk.v[static_cast<size_t>(k)] = m[i + static_cast<size_t>(k)];
                blk.gathers[static_cast<size_t>(k)] = static_cast<uint8_t>(vt[static_cast<size_t>(m[i + static_cast<size_t>(k)])].n - 1);
            }
            blocks.push_back(blk);
        }
        for (size_t i = full; i < m.size(); ++i) rest.push_back(m[i]);
    }
    out.shared_blocks = static_cast<int32_t>(blocks.size());
    for (int v = V; v < vpad; ++v) rest.push_back(v);   // padding
    // mixed blocks: leftovers of the same bone stay together; the block's bone is its most frequent one
    for (size_t i = 0; i < rest.size(); i += kBlock) {
        Block blk;
        std::array<int, 256> seen{};
        int bone = vt[static_cast<size_t>(rest[i])].slot[0], bone_n = 0;
        for (int k = 0; k < kBlock; ++k) {
            if (rest[i + static_cast<size_t>(k)] >= V) continue;   // padding names no bone
            const int s = vt[static_cast<size_t>(rest[i + static_cast<size_t>(k)])].slot[0];
            if (++seen[static_cast<size_t>(s)] > bone_n) {
                bone_n = seen[static_cast<size_t>(s)];
                bone = s;
            }
        }
        blk.bone = bone;
        std::array<int, kBlock> idx{};
        for (int k = 0; k < kBlock; ++k) idx[static_cast<size_t>(k)] = rest[i + static_cast<size_t>(k)];
        auto gathers = [&](int v) {
            const Vtx& x = vt[static_cast<size_t>(v)];
            return x.n - (x.slot[0] == bone ? 1 : 0);
        };
        // (padding vertices stay behind the real ones: stored positions 0 .. V-1 must be the caller's vertices)
        std::stable_sort(idx.begin(), idx.end(), [&](int a, int c) { return (a >= V) != (c >= V) ? c >= V : gathers(a) > gathers(c); });
        // position 0 names the block's bone for the kernel: the carrier with the most gathers goes first
        for (int k = 0; k < kBlock; ++k)
            if (idx[static_cast<size_t>(k)] < V && vt[static_cast<size_t>(idx[static_cast<size_t>(k)])].slot[0] == bone) {
                std::rotate(idx.begin(), idx.begin() + k, idx.begin() + k + 1);
                break;
            }
        for (int k = 0; k < kBlock; ++k) {
            blk.v[static_cast<size_t>(k)] = idx[static_cast<size_t>(k)];
            blk.gathers[static_cast<size_t>(k)] = static_cast<uint8_t>(gathers(idx[static_cast<size_t>(k)]));
        }
        blocks.push_back(blk);
    }

    // ---- 3. heavy blocks first, equal profiles next to each other; the block that holds the padding stays last ----
    const bool has_padding = vpad > V;
    std::vector<int> seq(blocks.size() - (has_padding ? 1 : 0));
    for (size_t i = 0; i < seq.size(); ++i) seq[i] = static_cast<int>(i);
    std::stable_sort(seq.begin(), seq.end(), [&](int a, int c) { return blocks[static_cast<size_t>(a)].gathers > blocks[static_cast<size_t>(c)].gathers; });

    // ---- 4. deal quarter-warps of 8 blocks over tiles, and inside a tile over its warps, heavy to light ----
    const int tb = tile_vertices / kBlock;                         // blocks per tile
    std::vector<int> place(static_cast<size_t>(nb), -1);           // stored block position -> block
    const int dealt_end = has_padding ? nb - 1 : nb;               // positions the sorted sequence is dealt over
    if (has_padding) place[static_cast<size_t>(nb - 1)] = static_cast<int>(blocks.size()) - 1;
    struct Slot { int first, capacity; };
    std::vector<std::vector<Slot>> tile_slots(static_cast<size_t>(num_tiles));
    size_t max_slots = 0;
    for (int t = 0; t < num_tiles; ++t) {
        const int begin = t * tb, end = std::min(dealt_end, (t + 1) * tb);
        if (begin >= end) continue;
        const int len = end - begin, warps = (len + blocks_per_warp - 1) / blocks_per_warp;
        for (int quarter = 0; quarter < quarters; ++quarter)
            for (int w = 0; w < warps; ++w) {
                const int pos = (w * quarters + quarter) * kQuarter;
                if (pos < len) tile_slots[static_cast<size_t>(t)].push_back({begin + pos, std::min(kQuarter, len - pos)});
            }
        max_slots = std::max(max_slots, tile_slots[static_cast<size_t>(t)].size());
    }
    size_t next = 0;
    for (size_t s = 0; s < max_slots; ++s)
        for (int t = 0; t < num_tiles; ++t) {
            const auto& slots = tile_slots[static_cast<size_t>(t)];
            if (s >= slots.size()) continue;
            for (int k = 0; k < slots[s].capacity && next < seq.size(); ++k) place[static_cast<size_t>(slots[s].first + k)] = seq[next++];
        }
    // a tile shape that does not cover every block (never the plan's own): the rest in order
    for (int pos = 0; pos < dealt_end; ++pos)
        if (place[static_cast<size_t>(pos)] < 0 && next < seq.size()) place[static_cast<size_t>(pos)] = seq[next++];

    // ---- 5. stored vertices + the wavefront count this order gives ----
    for (int pos = 0; pos < nb; ++pos) {
        const Block& blk = blocks[static_cast<size_t>(place[static_cast<size_t>(pos)])];
        for (int k = 0; k < kBlock; ++k) {
            const Vtx& x = vt[static_cast<size_t>(blk.v[static_cast<size_t>(k)])];
            const size_t s = static_cast<size_t>(pos) * kBlock + static_cast<size_t>(k);
            out.order[s] = x.id;
            std::memcpy(out.stored[s].slot, x.slot, 4);
            std::memcpy(out.stored[s].weight, x.weight, sizeof(float) * 4);
        }
    }
    for (int t = 0; t * tb < nb; ++t) {
        const int begin = t * tb, end = std::min(nb, (t + 1) * tb);
        for (int q0 = begin; q0 < end; q0 += kQuarter) {
            const int q1 = std::min(end, q0 + kQuarter);
            out.gather_wavefronts += blocks_per_warp == 16 ? 6 : 3;   // the block bones of the 8 lanes (read by both lanes of a pair)
            for (int k = 0; k < kBlock; ++k) {
                uint32_t any = 0;
                for (int pos = q0; pos < q1; ++pos) {
                    const size_t s = static_cast<size_t>(pos) * kBlock + static_cast<size_t>(k);
                    const ClusterInfluence& x = out.stored[s];
                    const uint8_t bone = out.stored[static_cast<size_t>(pos) * kBlock].slot[0];
                    uint32_t used = x.slot[0] != bone ? 1u : 0u;
                    for (int c = 1; c < 4; ++c) used |= x.weight[c] != 0.f ? (1u << c) : 0u;
                    any |= used;
                }
                for (int c = 0; c < 4; ++c) out.gather_wavefronts += (any >> c & 1u) ? 3 : 0;
            }
            out.gather_wavefronts_unclustered += 3 * 4 * kBlock;
        }
    }
    return out;
}

}  // namespace dmk
