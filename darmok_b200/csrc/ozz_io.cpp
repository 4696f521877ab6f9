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
        if (n) std::memcpy(dst, _p + _pos, n);   // (an empty destination may be a null pointer)
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

namespace {

// ---- Animation archive version 6 (ozz-animation <= 0.14): every key carries its ratio and track ----
bool parse_animation_v6(Reader& rd, AnimationData& out) {
    int32_t num_tracks = 0, name_len = 0, nt = 0, nr = 0, ns = 0;
    bool ok = rd.get(out.duration) && rd.get(num_tracks) && rd.get(name_len) && rd.get(nt) && rd.get(nr) && rd.get(ns);
    ok = ok && num_tracks >= 0 && num_tracks <= kMaxJoints && name_len >= 0 && nt >= 0 && nr >= 0 && ns >= 0;
    ok = ok && static_cast<size_t>(name_len) + 12u * static_cast<size_t>(nt) + 14u * static_cast<size_t>(nr) +
                       12u * static_cast<size_t>(ns) <= rd.remaining();
    if (!ok) return false;
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
    return ok;
}

// ---- Animation archive version 7 (ozz-animation >= 0.15) -- layout from memory, see ozz_io.hpp ----
//   f32 duration; u32 num_tracks; u32 name_len; u32 timepoints_count; u32 translation_count, rotation_count, scale_count;
//   u32 {iframe_entries_count, iframe_desc_count} x {translations, rotations, scales};
//   char name[name_len]; f32 timepoints[timepoints_count];
//   then per component (translations, rotations, scales) its KeyframesCtrl and values:
//     ratios      u8[count] when timepoints_count <= 256, else u16[count]: index of the key's ratio in timepoints
//     previouses  u16[count]: distance back to the previous key of the same track (the first key of track i is key i)
//     iframe_entries u8[], iframe_desc u32[], f32 iframe_interval: cached cursor snapshots for seeking -- read and ignored,
//                 the stateless key search here needs none
//     values      u16[3] per key: three IEEE halves (Float3Key), or for a rotation 48 bits
//                 [0:1] largest component, [2] its sign, [3:17] [18:32] [33:47] the other three as 15-bit magnitudes
//                 mapped from [-1/sqrt 2, 1/sqrt 2]
struct CtrlCounts { uint32_t keys = 0, iframe_entries = 0, iframe_desc = 0; };

template <typename Key, typename SetValue>
bool read_component_v7(Reader& rd, const CtrlCounts& n, const std::vector<float>& timepoints, int padded, std::vector<Key>& keys, SetValue set_value) {
    const bool wide = timepoints.size() > 256;
    const size_t need = static_cast<size_t>(n.keys) * ((wide ? 2u : 1u) + 2u + 6u) + n.iframe_entries + 4u * static_cast<size_t>(n.iframe_desc) + 4u;
    if (need > rd.remaining()) return false;
    keys.resize(n.keys);
    bool ok = true;
    for (auto& k : keys) {
        uint32_t index = 0;
        if (wide) {
            uint16_t v = 0;
            ok = ok && rd.get(v);
            index = v;
        } else {
            uint8_t v = 0;
            ok = ok && rd.get(v);
            index = v;
        }
        ok = ok && index < timepoints.size();
        k.ratio = ok ? timepoints[index] : 0.f;
    }
    for (size_t i = 0; ok && i < keys.size(); ++i) {
        uint16_t back = 0;
        ok = rd.get(back);
        if (i < static_cast<size_t>(padded)) {
            keys[i].track = static_cast<uint16_t>(i);                  // key 0 of every track opens the stream, in track order
        } else {
            ok = ok && back >= 1 && back <= i;
            if (ok) keys[i].track = keys[i - back].track;
        }
    }
    std::vector<uint8_t> skip(static_cast<size_t>(n.iframe_entries) + 4u * static_cast<size_t>(n.iframe_desc));
    float interval = 0.f;
    ok = ok && rd.raw(skip.data(), skip.size()) && rd.get(interval);
    for (auto& k : keys) {
        uint16_t v[3] = {0, 0, 0};
        ok = ok && rd.get_array(v, 3);
        set_value(k, v);
    }
    return ok;
}

bool parse_animation_v7(Reader& rd, AnimationData& out) {
    uint32_t num_tracks = 0, name_len = 0, timepoints_count = 0;
    CtrlCounts t, r, s;
    bool ok = rd.get(out.duration) && rd.get(num_tracks) && rd.get(name_len) && rd.get(timepoints_count) && rd.get(t.keys) && rd.get(r.keys) &&
              rd.get(s.keys) && rd.get(t.iframe_entries) && rd.get(t.iframe_desc) && rd.get(r.iframe_entries) && rd.get(r.iframe_desc) &&
              rd.get(s.iframe_entries) && rd.get(s.iframe_desc);
    ok = ok && num_tracks <= static_cast<uint32_t>(kMaxJoints) && timepoints_count <= 65536u;
    ok = ok && static_cast<size_t>(name_len) + 4u * static_cast<size_t>(timepoints_count) <= rd.remaining();
    if (!ok) return false;
    out.num_tracks = static_cast<int>(num_tracks);
    out.name.resize(name_len);
    rd.raw(out.name.data(), out.name.size());
    std::vector<float> timepoints(timepoints_count);
    if (!rd.get_array(timepoints.data(), timepoints.size())) return false;
    const int padded = out.num_soa() * 4;
    auto f3 = [](Float3Key& k, const uint16_t v[3]) { k.value[0] = v[0]; k.value[1] = v[1]; k.value[2] = v[2]; };
    auto quat = [](QuaternionKey& k, const uint16_t v[3]) {
        const uint64_t packed = static_cast<uint64_t>(v[0]) | (static_cast<uint64_t>(v[1]) << 16) | (static_cast<uint64_t>(v[2]) << 32);
        k.largest = static_cast<uint8_t>(packed & 3u);
        k.sign = static_cast<uint8_t>((packed >> 2) & 1u);
        k.value[0] = static_cast<int16_t>((packed >> 3) & 0x7fffu);
        k.value[1] = static_cast<int16_t>((packed >> 18) & 0x7fffu);
        k.value[2] = static_cast<int16_t>((packed >> 33) & 0x7fffu);
        k.format = 1;
    };
    return read_component_v7(rd, t, timepoints, padded, out.translations, f3) && read_component_v7(rd, r, timepoints, padded, out.rotations, quat) &&
           read_component_v7(rd, s, timepoints, padded, out.scales, f3);
}

}  // namespace

bool parse_animation(const void* bytes, size_t size, AnimationData& out, std::string& err) {
    Reader rd(bytes, size);
    uint32_t version = 0;
    if (!bytes || !rd.open("ozz-animation", version)) {
        err = "archive does not contain the type";
        return false;
    }
    if (version != 6 && version != 7) {
        err = "unsupported ozz animation archive version " + std::to_string(version) + " (supported: 6, 7)";
        return false;
    }
    out = AnimationData{};
    bool ok = version == 6 ? parse_animation_v6(rd, out) : parse_animation_v7(rd, out);
    // what the sampling cursor relies on, whatever the archive version
    const int padded = out.num_soa() * 4;
    const size_t nt = out.translations.size(), nr = out.rotations.size(), ns = out.scales.size();
    if (ok && out.num_tracks > 0) ok = nt >= 2u * static_cast<size_t>(padded) && nr >= 2u * static_cast<size_t>(padded) && ns >= 2u * static_cast<size_t>(padded);
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
