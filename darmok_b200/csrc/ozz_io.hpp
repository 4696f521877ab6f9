// ozz_io.hpp -- reader for the ozz-animation binary archive container (runtime Skeleton / Animation objects).
//
// Replaces `ozz::io::IArchive >> ozz::animation::{Skeleton,Animation}` as driven by OzzUtils::loadFromData
// (darmok src/skeleton_ozz.cpp:27-44) over DataOzzStream (:121-184).  Byte layout: SURVEY.md Appendix A.1-A.3
// (Skeleton version 2; Animation version 6 = ozz-animation 0.14, and version 7 = ozz-animation >= 0.15, which is what a darmok
// build of today links: CMake/ozzAnimationConfig.cmake fetches `GIT_TAG master`).  Both animation versions decode into the
// same key lists (ratio, track, value); everything downstream of the reader -- SoA re-layout, bucket tables, sampling -- is
// version independent except for the quantisation of a rotation key (QuaternionKey::format).
// NOTE: ozz-animation is not in the build container.  The v6 layout follows SURVEY.md Appendix A.3 (medium confidence in the
// field order); the v7 layout (A.3', parse_animation_v7 in ozz_io.cpp) is written FROM MEMORY of upstream 0.15 and has never
// met a file written by ozz: it is validated against this repository's own writer only (darmok_b200/synth.py).  The version
// switch is one function per version so that a real file can correct either without touching anything else.
#pragma once

#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace dmk {

struct SkeletonData {
    int num_joints = 0;
    std::vector<std::string> names;
    std::vector<int16_t> parents;
    std::vector<float> rest_soa;  // SoaTransform[(J + 3) / 4], 40 floats each
    int num_soa() const { return (num_joints + 3) / 4; }
};

struct Float3Key {
    float ratio;
    uint16_t track;
    uint16_t value[3];  // IEEE half
};
struct QuaternionKey {
    float ratio;
    uint16_t track;
    uint8_t largest;
    uint8_t sign;
    int16_t value[3];    // format 0 (archive v6): signed, c = value / (32767 sqrt 2); format 1 (v7): 0 .. 32767, c = value sqrt 2 / 32767 - 1 / sqrt 2
    uint8_t format = 0;
};

struct AnimationData {
    float duration = 0.f;
    int num_tracks = 0;
    std::string name;
    std::vector<Float3Key> translations;
    std::vector<QuaternionKey> rotations;
    std::vector<Float3Key> scales;
    int num_soa() const { return (num_tracks + 3) / 4; }
};

// false + err on failure; err uses the reference's strings where it has one
bool parse_skeleton(const void* bytes, size_t size, SkeletonData& out, std::string& err);
bool parse_animation(const void* bytes, size_t size, AnimationData& out, std::string& err);

}  // namespace dmk
