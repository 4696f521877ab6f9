// ozz_io.cpp -- see ozz_io.hpp
#include "ozz_io.hpp"

#include <cstring>

namespace dmk {

namespace {

class Reader {
public:
    Reader(const void* bytes, size_t size) : _p(static_cast<const uint8_t*>(bytes)), _n(size) {}

    // IArchive constructor: one endianness byte; then TestTag<T>() / operator>>: tag string + uint32 version
    bool open(const char* tag, uint32_t& version) {
        uint8_t endian = 0;
        if (!raw(&endian, 1)) return false;
        const uint16_t probe = 1;
        const bool host_little = *reinterpret_cast<const uint8_t*>(&probe) == 1;
        _swap = (endian != 0) != host_little;
        const size_t len = std::strlen(tag) + 1;
        if (_pos + len > _n || std::memcmp(_p + _pos, tag, len) != 0) return false;
        _pos += len;
        return get(version);
    }

    template <typename T>
    bool get(T& v) {
        if (!raw(&v, sizeof(T))) return false;
        if (_swap && sizeof(T) > 1) {
            auto* b = reinterpret_cast<uint8_t*>(&v);
            for (size_t i = 0; i < sizeof(T) / 2; ++i) std::swap(b[i], b[sizeof(T) - 1 - i]);
        }
        return true;
    }
    template <typename T>
    bool get_array(T* v, size_t count) {
        for (size_t i = 0; i < count; ++i)
            if (!get(v[i])) return false;
        return true;
    }
    bool raw(void* dst, size_t n) {
        if (_pos + n > _n) return false;
        std::memcpy(dst, _p + _pos, n);
        _pos += n;
        return true;
    }
    size_t remaining() const { return _n - _pos; }

private:
    const uint8_t* _p;
    size_t _n;
    size_t _pos = 0;
    bool _swap = false;
};

constexpr int kMaxJoints = 1024;  // ozz::animation::Skeleton::kMaxJoints

}  // namespace

bool parse_skeleton(const void* bytes, size_t size, SkeletonData& out, std::string& err) {
    Reader rd(bytes, size);
    uint32_t version = 0;
    if (!bytes || !rd.open("ozz-skeleton", version)) {
        err = "archive does not contain the type";  // darmok src/skeleton_ozz.cpp:39
        return false;
    }
    if (version != 2) {
        err = "unsupported ozz skeleton archive version " + std::to_string(version) + " (supported: 2)";
        return false;
    }
    int32_t num_joints = 0;
    if (!rd.get(num_joints) || num_joints < 0 || num_joints > kMaxJoints) {
        err = "corrupt ozz skeleton archive";
        return false;
    }
    out = SkeletonData{};
    out.num_joints = num_joints;
    if (num_joints == 0) return true;
    int32_t chars = 0;
    if (!rd.get(chars) || chars < num_joints || static_cast<size_t>(chars) > rd.remaining()) {
        err = "corrupt ozz skeleton archive";
        return false;
    }
    std::vector<char> buf(static_cast<size_t>(chars) + 1, '\0');
    rd.raw(buf.data(), static_cast<size_t>(chars));
    const char* c = buf.data();
    const char* end = buf.data() + chars;
    for (int i = 0; i < num_joints; ++i) {
        if (c >= end) {
            err = "corrupt ozz skeleton archive";
            return false;
        }
        out.names.emplace_back(c);
        c += out.names.back().size() + 1;
    }
    out.parents.resize(num_joints);
    out.rest_soa.resize(static_cast<size_t>(out.num_soa()) * 40);
    if (!rd.get_array(out.parents.data(), out.parents.size()) || !rd.get_array(out.rest_soa.data(), out.rest_soa.size())) {
        err = "corrupt ozz skeleton archive";
        return false;
    }
    for (int i = 0; i < num_joints; ++i) {
        // LocalToModelJob needs parents before children (depth-first storage)
        if (out.parents[i] >= i || out.parents[i] < -1) {
            err = "corrupt ozz skeleton archive";
            return false;
        }
    }
    return true;
}

bool parse_animation(const void* bytes, size_t size, AnimationData& out, std::string& err) {
    Reader rd(bytes, size);
    uint32_t version = 0;
    if (!bytes || !rd.open("ozz-animation", version)) {
        err = "archive does not contain the type";
        return false;
    }
    if (version != 6) {
        err = "unsupported ozz animation archive version " + std::to_string(version) + " (supported: 6)";
        return false;
    }
    out = AnimationData{};
    int32_t num_tracks = 0, name_len = 0, nt = 0, nr = 0, ns = 0;
    bool ok = rd.get(out.duration) && rd.get(num_tracks) && rd.get(name_len) && rd.get(nt) && rd.get(nr) && rd.get(ns);
    ok = ok && num_tracks >= 0 && num_tracks <= kMaxJoints && name_len >= 0 && nt >= 0 && nr >= 0 && ns >= 0;
    ok = ok && static_cast<size_t>(name_len) + 12u * static_cast<size_t>(nt) + 14u * static_cast<size_t>(nr) +
                       12u * static_cast<size_t>(ns) <= rd.remaining();
    if (!ok) {
        err = "corrupt ozz animation archive";
        return false;
    }
    out.num_tracks = num_tracks;
    out.name.resize(static_cast<size_t>(name_len));
    rd.raw(out.name.data(), out.name.size());
    out.translations.resize(nt);
    out.rotations.resize(nr);
    out.scales.resize(ns);
    auto read_f3 = [&](std::vector<Float3Key>& keys) {
        for (auto& k : keys) ok = ok && rd.get(k.ratio) && rd.get(k.track) && rd.get_array(k.value, 3);
    };
    read_f3(out.translations);
    for (auto& k : out.rotations)
        ok = ok && rd.get(k.ratio) && rd.get(k.track) && rd.get(k.largest) && rd.get(k.sign) && rd.get_array(k.value, 3);
    read_f3(out.scales);
    const int padded = out.num_soa() * 4;
    if (num_tracks > 0) ok = ok && nt >= 2 * padded && nr >= 2 * padded && ns >= 2 * padded;
    for (auto& k : out.translations) ok = ok && k.track < padded;
    for (auto& k : out.rotations) ok = ok && k.track < padded && k.largest < 4 && k.sign < 2;
    for (auto& k : out.scales) ok = ok && k.track < padded;
    if (!ok) {
        err = "corrupt ozz animation archive";
        return false;
    }
    return true;
}

}  // namespace dmk
