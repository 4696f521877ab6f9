// ozz_io.hpp -- reader for the ozz-animation binary archive container (runtime Skeleton / Animation objects).
//
// Replaces `ozz::io::IArchive >> ozz::animation::{Skeleton,Animation}` as driven by OzzUtils::loadFromData
// (darmok src/skeleton_ozz.cpp:27-44) over DataOzzStream (:121-184).  Byte layout: SURVEY.md Appendix A.1-A.3
// (ozz-animation 0.14: Skeleton version 2, Animation version 6).
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
    int16_t value[3];
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
