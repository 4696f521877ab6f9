// ozz_io_fuzz.cpp -- the product's ozz archive reader (darmok_b200/csrc/ozz_io.cpp) on damaged archives, built with
// AddressSanitizer + UBSan: every prefix of a valid skeleton / animation archive and seeded random corruptions of them must
// be answered with true or false, never with a crash or an out-of-bounds access, and whatever is accepted must satisfy the
// invariants the device upload relies on.
//
//   ozz_io_fuzz <skeleton.ozz> <animation.ozz> [mutations] [seed]
#include "ozz_io.hpp"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iterator>

static std::vector<uint8_t> readFile(const char* path) {
    std::ifstream f(path, std::ios::binary);
    return std::vector<uint8_t>(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
}

static int g_failures = 0;
#define CHECK(cond, ...) do { if (!(cond)) { std::printf("FAIL: "); std::printf(__VA_ARGS__); std::printf("\n"); ++g_failures; } } while (0)

static bool trySkeleton(const std::vector<uint8_t>& bytes) {
    // an exactly sized heap copy so that ASan sees any read past the end
    std::vector<uint8_t> copy(bytes);
    dmk::SkeletonData s;
    std::string err;
    const bool ok = dmk::parse_skeleton(copy.empty() ? nullptr : copy.data(), copy.size(), s, err);
    if (!ok) { CHECK(!err.empty(), "rejected skeleton without a message"); return false; }
    CHECK(s.num_joints >= 0 && s.num_joints <= 1024, "joint count %d", s.num_joints);
    CHECK((int)s.names.size() == s.num_joints && (int)s.parents.size() == s.num_joints, "skeleton table sizes");
    CHECK(s.rest_soa.size() == (size_t)s.num_soa() * 40, "rest pose size");
    for (int i = 0; i < s.num_joints; ++i) CHECK(s.parents[i] >= -1 && s.parents[i] < i, "parent %d of joint %d", s.parents[i], i);
    return true;
}

static bool tryAnimation(const std::vector<uint8_t>& bytes) {
    std::vector<uint8_t> copy(bytes);
    dmk::AnimationData a;
    std::string err;
    const bool ok = dmk::parse_animation(copy.empty() ? nullptr : copy.data(), copy.size(), a, err);
    if (!ok) { CHECK(!err.empty(), "rejected animation without a message"); return false; }
    const int padded = a.num_soa() * 4;
    CHECK(a.num_tracks >= 0 && a.num_tracks <= 1024, "track count %d", a.num_tracks);
    for (const auto& k : a.translations) CHECK(k.track < padded, "translation track %d", k.track);
    for (const auto& k : a.scales) CHECK(k.track < padded, "scale track %d", k.track);
    for (const auto& k : a.rotations) CHECK(k.track < padded && k.largest < 4 && k.sign < 2, "rotation key fields");
    if (a.num_tracks > 0)
        CHECK((int)a.translations.size() >= 2 * padded && (int)a.rotations.size() >= 2 * padded && (int)a.scales.size() >= 2 * padded, "fewer than two keys per track");
    return true;
}

int main(int argc, char** argv) {
    if (argc < 3) { std::printf("usage: ozz_io_fuzz <skeleton.ozz> <animation.ozz> [mutations] [seed]\n"); return 2; }
    const std::vector<uint8_t> skel = readFile(argv[1]), anim = readFile(argv[2]);
    const int mutations = argc > 3 ? std::atoi(argv[3]) : 20000;
    uint64_t state = (argc > 4 ? std::strtoull(argv[4], nullptr, 10) : 1) * 0x9E3779B97F4A7C15ull + 99;
    auto next = [&]() { state ^= state >> 12; state ^= state << 25; state ^= state >> 27; return state * 0x2545F4914F6CDD1Dull; };
    CHECK(trySkeleton(skel) && tryAnimation(anim), "the undamaged archives must load");
    CHECK(!trySkeleton(anim) && !tryAnimation(skel), "an archive of the other type must be rejected");
    int accepted = 0, rejected = 0;
    // every proper prefix is a truncated file
    for (size_t n = 0; n < skel.size(); ++n) (trySkeleton(std::vector<uint8_t>(skel.begin(), skel.begin() + n)) ? accepted : rejected)++;
    const size_t step = anim.size() > 4096 ? anim.size() / 4096 : 1;   // 4096 cut points, plus every byte of the header
    for (size_t n = 0; n < anim.size(); n += (n < 256 ? 1 : step)) (tryAnimation(std::vector<uint8_t>(anim.begin(), anim.begin() + n)) ? accepted : rejected)++;
    CHECK(accepted == 0, "%d truncated archives were accepted", accepted);
    // random corruptions: byte flips anywhere, 32-bit overwrites in the header (counts, sizes, version), tail cut + flip
    for (int i = 0; i < mutations; ++i) {
        const bool which = next() & 1;
        std::vector<uint8_t> m = which ? skel : anim;
        const int edits = 1 + (int)(next() % 4);
        for (int e = 0; e < edits; ++e) {
            const uint64_t r = next();
            const size_t header = m.size() < 64 ? m.size() : 64;
            switch (r % 4) {
            case 0: m[(r >> 8) % m.size()] ^= (uint8_t)(1u << ((r >> 40) % 8)); break;
            case 1: m[(r >> 8) % m.size()] = (uint8_t)(r >> 48); break;
            case 2: {
                const uint32_t vals[] = {0u, 1u, 0x7fffffffu, 0x80000000u, 0xffffffffu, 1024u, 1025u, 65536u, (uint32_t)(r >> 32)};
                const uint32_t v = vals[(r >> 4) % 9];
                if (header >= 4) std::memcpy(m.data() + (r >> 16) % (header - 3), &v, 4);
                break;
            }
            default: m.resize(1 + (r >> 8) % m.size()); break;
            }
        }
        ((which ? trySkeleton(m) : tryAnimation(m)) ? accepted : rejected)++;
    }
    std::printf("ozz_io fuzz: %d damaged archives rejected, %d accepted with valid invariants, %d failures\n", rejected, accepted, g_failures);
    return g_failures ? 1 : 0;
}
