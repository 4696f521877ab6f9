// mesh_cluster.cpp -- see mesh_cluster.hpp.  Host code only (no CUDA), deterministic for a given input.
#include "mesh_cluster.hpp"

#include <algorithm>
#include <array>
#include <cstdlib>
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


// A home bone per vertex (entry 0 afterwards), chosen so that as many bones as possible hold a multiple of 8 vertices
void assign_home_bones(std::vector<Vtx>& vt, int V) {
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

}

// Orders the vertices of one block: the carrier of the block's bone with the most gathers first (position 0 names the bone for the
// kernel), then by gathers, heavy to light; returns the gathers per position.  `idx` may hold fewer than 8 vertices.
std::array<uint8_t, kBlock> order_block(const std::vector<Vtx>& vt, std::vector<int>& idx, int bone) {
    auto gathers = [&](int v) {
        const Vtx& x = vt[static_cast<size_t>(v)];
        return x.n - (x.slot[0] == bone ? 1 : 0);
    };
    std::stable_sort(idx.begin(), idx.end(), [&](int a, int c) { return gathers(a) > gathers(c); });
    for (size_t k = 0; k < idx.size(); ++k)
        if (vt[static_cast<size_t>(idx[k])].slot[0] == bone) {
            std::rotate(idx.begin(), idx.begin() + static_cast<long>(k), idx.begin() + static_cast<long>(k) + 1);
            break;
        }
    std::array<uint8_t, kBlock> g{};
    for (size_t k = 0; k < idx.size(); ++k) g[k] = static_cast<uint8_t>(gathers(idx[k]));
    return g;
}

int most_frequent_bone(const std::vector<Vtx>& vt, const std::vector<int>& idx) {
    std::array<int, 256> seen{};
    int bone = 0, bone_n = 0;
    for (int v : idx) {
        const int b = vt[static_cast<size_t>(v)].slot[0];
        if (++seen[static_cast<size_t>(b)] > bone_n) {
            bone_n = seen[static_cast<size_t>(b)];
            bone = b;
        }
    }
    return bone;
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

    assign_home_bones(vt, V);

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
        // (serpentine: odd rounds run over the warps backwards, so that no warp always gets the heavier group of its round -- a block's
        // characters advance at the pace of its slowest warp)
        for (int quarter = 0; quarter < quarters; ++quarter)
            for (int wi = 0; wi < warps; ++wi) {
                const int w = (quarter & 1) ? warps - 1 - wi : wi;
                const int pos = (w * quarters + quarter) * kQuarter;
                if (pos < len) tile_slots[static_cast<size_t>(t)].push_back({begin + pos, std::min(kQuarter, len - pos)});
            }
        max_slots = std::max(max_slots, tile_slots[static_cast<size_t>(t)].size());
    }
    size_t next = 0;
    for (size_t s = 0; s < max_slots; ++s)
        for (int ti = 0; ti < num_tiles; ++ti) {   // (serpentine over the tiles as well: every tile gets the same share of heavy groups)
            const int t = (s & 1) ? num_tiles - 1 - ti : ti;
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


// ---- strided ownership (round 2: the skinning kernel's bulk-store output path) -----------------------------------------------------
// Thread (warp w, lane l) of a tile owns the stored vertices  tile_begin + 256 w + k nl + l,  k = 0..7, nl = lanes of warp w that have
// vertices (32, fewer in a tile's last warp): round k of a warp is then nl consecutive vertices = ONE contiguous piece of the output.
// A "block" is still what a thread owns; stored positions >= V are padding, which makes the blocks of the last warp of the last tile
// shorter (capacity < 8) instead of one block holding all the padding.
namespace {
struct StridedGeometry {
    int tile_vertices, vpad, V;
    int blocks_per_tile() const { return tile_vertices / kBlock; }
    int num_blocks() const { return vpad / kBlock; }
    int tile_of(int g) const { return g / blocks_per_tile(); }
    int tile_blocks(int t) const { return std::min(blocks_per_tile(), num_blocks() - t * blocks_per_tile()); }
    // stored position of vertex k of block g
    int position(int g, int k) const {
        const int t = tile_of(g), p = g - t * blocks_per_tile(), w = p / 32, l = p % 32;
        const int nl = std::min(32, tile_blocks(t) - 32 * w);
        return t * tile_vertices + w * 256 + k * nl + l;
    }
    int capacity(int g) const {
        int c = 0;
        for (int k = 0; k < kBlock; ++k) c += position(g, k) < V ? 1 : 0;
        return c;
    }
};
struct SBlock {
    std::vector<int> v;   // vertices (indices into the vertex table), position 0 carries the block's bone
    int bone = 0;
    std::array<uint8_t, kBlock> gathers{};
};
}  // namespace

ClusterResult cluster_mesh_strided(const std::vector<ClusterInfluence>& in, int vpad, int num_tiles, int tile_vertices) {
    ClusterResult out;
    const int V = static_cast<int>(in.size());
    out.order.assign(static_cast<size_t>(vpad), -1);
    out.stored.resize(static_cast<size_t>(vpad));
    for (auto& x : out.stored) {
        std::memset(x.slot, 0, 4);
        x.weight[0] = x.weight[1] = x.weight[2] = x.weight[3] = 0.f;
    }
    if (vpad <= 0 || vpad % 64 || tile_vertices % 64 || tile_vertices <= 0 || V > vpad) return out;   // (the caller checked: not this layout)
    (void)num_tiles;
    const StridedGeometry geo{tile_vertices, vpad, V};
    const int nb = geo.num_blocks();
    out.total_blocks = nb;

    std::vector<Vtx> vt(static_cast<size_t>(V));
    for (int v = 0; v < V; ++v) vt[static_cast<size_t>(v)] = normalise(in[static_cast<size_t>(v)], v);
    assign_home_bones(vt, V);

    // capacities: full blocks and the short ones (their positions are fixed by the geometry)
    std::vector<int> special;   // block indices with capacity < 8, in position order
    int full_needed = 0;
    for (int g = 0; g < nb; ++g) {
        if (geo.capacity(g) == kBlock) ++full_needed;
        else special.push_back(g);
    }

    // ---- blocks per bone, vertices by influence count.  What a bone has beyond the full blocks needed stays in its pool; the short
    // blocks are cut from the largest pools (so they share a bone too), the rest -- bone by bone -- fills mixed blocks ----
    std::vector<std::vector<int>> members(256);
    for (int v = 0; v < V; ++v) members[vt[static_cast<size_t>(v)].slot[0]].push_back(v);
    // short blocks first take their share out of the bones with the most vertices (largest capacity first)
    std::vector<SBlock> shorts(special.size());
    {
        std::vector<size_t> by_capacity(special.size());
        for (size_t i = 0; i < by_capacity.size(); ++i) by_capacity[i] = i;
        std::stable_sort(by_capacity.begin(), by_capacity.end(), [&](size_t a, size_t c) { return geo.capacity(special[a]) > geo.capacity(special[c]); });
        for (size_t i : by_capacity) {
            const int c = geo.capacity(special[i]);
            if (c == 0) continue;
            // the bone whose count is furthest above a multiple of 8 by at least c, else the largest bone that has c vertices
            int best = -1;
            for (int b = 0; b < 256; ++b) {
                const int m = static_cast<int>(members[static_cast<size_t>(b)].size());
                if (m < c) continue;
                if (best < 0) { best = b; continue; }
                const int mb = static_cast<int>(members[static_cast<size_t>(best)].size());
                const bool fits = m % kBlock >= c, best_fits = mb % kBlock >= c;
                if (fits != best_fits ? fits : m > mb) best = b;
            }
            if (best < 0) continue;   // no bone has that many vertices: filled from the rest below
            auto& m = members[static_cast<size_t>(best)];
            // the lightest vertices of the bone: the block sits in the lightest corner of the mesh (last warp of the last tile)
            std::stable_sort(m.begin(), m.end(), [&](int a, int d) { return vt[static_cast<size_t>(a)].n > vt[static_cast<size_t>(d)].n; });
            SBlock& blk = shorts[i];
            blk.bone = best;
            blk.v.assign(m.end() - c, m.end());
            m.resize(m.size() - static_cast<size_t>(c));
            blk.gathers = order_block(vt, blk.v, best);
            ++out.shared_blocks;
        }
    }
    std::vector<SBlock> full;
    std::vector<int> rest;
    for (int b = 0; b < 256; ++b) {
        auto& m = members[static_cast<size_t>(b)];
        std::stable_sort(m.begin(), m.end(), [&](int a, int c) { return vt[static_cast<size_t>(a)].n > vt[static_cast<size_t>(c)].n; });
        size_t i = 0;
        for (; i + kBlock <= m.size() && static_cast<int>(full.size()) < full_needed; i += kBlock) {
            SBlock blk;
            blk.bone = b;
            blk.v.assign(m.begin() + static_cast<long>(i), m.begin() + static_cast<long>(i) + kBlock);
            blk.gathers = order_block(vt, blk.v, b);
            full.push_back(std::move(blk));
        }
        for (; i < m.size(); ++i) rest.push_back(m[i]);
    }
    out.shared_blocks += static_cast<int32_t>(full.size());
    // mixed full blocks, then the short blocks no bone could fill, cut from `rest` in order (leftovers of one bone stay together)
    size_t next_rest = 0;
    auto cut = [&](int count) {
        SBlock blk;
        for (int k = 0; k < count && next_rest < rest.size(); ++k) blk.v.push_back(rest[next_rest++]);
        blk.bone = blk.v.empty() ? 0 : most_frequent_bone(vt, blk.v);
        blk.gathers = order_block(vt, blk.v, blk.bone);
        return blk;
    };
    while (static_cast<int>(full.size()) < full_needed) full.push_back(cut(kBlock));
    for (size_t i = 0; i < special.size(); ++i)
        if (shorts[i].v.empty() && geo.capacity(special[i]) > 0) shorts[i] = cut(geo.capacity(special[i]));

    // ---- heavy blocks first, equal profiles next to each other; deal quarter-warps of 8 blocks over tiles and warps ----
    std::vector<int> seq(full.size());
    for (size_t i = 0; i < seq.size(); ++i) seq[i] = static_cast<int>(i);
    std::stable_sort(seq.begin(), seq.end(), [&](int a, int c) { return full[static_cast<size_t>(a)].gathers > full[static_cast<size_t>(c)].gathers; });
    std::vector<const SBlock*> place(static_cast<size_t>(nb), nullptr);
    for (size_t i = 0; i < special.size(); ++i) place[static_cast<size_t>(special[i])] = &shorts[i];
    const int tb = geo.blocks_per_tile();
    const int tiles = (nb + tb - 1) / tb;
    struct Slot { int first, capacity; };
    std::vector<std::vector<Slot>> tile_slots(static_cast<size_t>(tiles));
    size_t max_slots = 0;
    for (int t = 0; t < tiles; ++t) {
        const int begin = t * tb, len = geo.tile_blocks(t), warps = (len + 31) / 32;
        for (int quarter = 0; quarter < 4; ++quarter)
            for (int wi = 0; wi < warps; ++wi) {
                const int w = (quarter & 1) ? warps - 1 - wi : wi;   // serpentine, as in cluster_mesh
                const int pos = (w * 4 + quarter) * kQuarter;
                if (pos < len) tile_slots[static_cast<size_t>(t)].push_back({begin + pos, std::min(kQuarter, len - pos)});
            }
        max_slots = std::max(max_slots, tile_slots[static_cast<size_t>(t)].size());
    }
    size_t next = 0;
    for (size_t sl = 0; sl < max_slots; ++sl)
        for (int ti = 0; ti < tiles; ++ti) {
            const int t = (sl & 1) ? tiles - 1 - ti : ti;
            const auto& slots = tile_slots[static_cast<size_t>(t)];
            if (sl >= slots.size()) continue;
            for (int k = 0; k < slots[sl].capacity; ++k) {
                const size_t g = static_cast<size_t>(slots[sl].first + k);
                if (!place[g] && next < seq.size()) place[g] = &full[static_cast<size_t>(seq[next++])];
            }
        }

    // ---- stored vertices; padding positions of a short block name the block's bone with weight 0 (no gather, output 0) ----
    for (int g = 0; g < nb; ++g) {
        const SBlock* blk = place[static_cast<size_t>(g)];
        const int bone = blk ? blk->bone : 0;
        for (int k = 0; k < kBlock; ++k) {
            const size_t s = static_cast<size_t>(geo.position(g, k));
            ClusterInfluence& st = out.stored[s];
            if (blk && k < static_cast<int>(blk->v.size())) {
                const Vtx& x = vt[static_cast<size_t>(blk->v[static_cast<size_t>(k)])];
                out.order[s] = x.id;
                std::memcpy(st.slot, x.slot, 4);
                std::memcpy(st.weight, x.weight, sizeof(float) * 4);
            } else {
                st.slot[0] = static_cast<uint8_t>(bone);
            }
        }
    }
    // ---- the wavefront count this order gives (quarter-warp = 8 consecutive blocks of a tile) ----
    for (int t = 0; t < tiles; ++t) {
        const int begin = t * tb, end = begin + geo.tile_blocks(t);
        for (int q0 = begin; q0 < end; q0 += kQuarter) {
            const int q1 = std::min(end, q0 + kQuarter);
            out.gather_wavefronts += 3;   // the block bones of the 8 lanes
            for (int k = 0; k < kBlock; ++k) {
                uint32_t any = 0;
                for (int g = q0; g < q1; ++g) {
                    const ClusterInfluence& x = out.stored[static_cast<size_t>(geo.position(g, k))];
                    const uint8_t bone = out.stored[static_cast<size_t>(geo.position(g, 0))].slot[0];
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
