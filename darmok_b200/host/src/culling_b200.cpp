// culling_b200.cpp -- see darmok_b200/culling.hpp
#include <darmok_b200/culling.hpp>
#include <darmok_b200/skeleton.hpp>

#include <dmk.h>

#include <cstring>

namespace darmok {

FrustumCuller::FrustumCuller(std::shared_ptr<DeviceContext> ctx) noexcept : _ctx{std::move(ctx)} {}
FrustumCuller::~FrustumCuller() noexcept = default;

expected<void, std::string> FrustumCuller::init() noexcept {
    if (!_ctx) return unexpected<std::string>{"FrustumCuller: no device context"};
    return {};
}
expected<void, std::string> FrustumCuller::load(const Definition&) noexcept { return {}; }
expected<void, std::string> FrustumCuller::shutdown() noexcept {
    _index.clear();
    _worldInverse.clear();
    _bounds.clear();
    _visible.clear();
    _culledCount = 0;
    _dirty = true;
    return {};
}

expected<void, std::string> FrustumCuller::setEntities(const Entity* entities, const float* worldInverse, const float* bounds, size_t count) noexcept {
    if (count && (!entities || !bounds)) return unexpected<std::string>{"FrustumCuller: null entity or bounds array"};
    _index.clear();
    _index.reserve(count);
    for (size_t i = 0; i < count; ++i) _index[entities[i]] = i;
    if (_index.size() != count) return unexpected<std::string>{"FrustumCuller: an entity is listed twice"};
    _bounds.assign(bounds, bounds + count * 6);
    _worldInverse.resize(count * 16);
    if (worldInverse) {
        std::memcpy(_worldInverse.data(), worldInverse, count * 16 * sizeof(float));
    } else {   // no Transform: the camera frustum as it is (src/culling.cpp:271-275)
        static const float identity[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
        for (size_t i = 0; i < count; ++i) std::memcpy(&_worldInverse[i * 16], identity, sizeof identity);
    }
    _visible.assign((count + 31) / 32, 0u);
    _culledCount = 0;
    _dirty = true;
    return {};
}

void FrustumCuller::setViewProjection(const glm::mat4& viewProj, float normalizedNearDepth) noexcept {
    _viewProj = viewProj;
    _nearDepth = normalizedNearDepth;
    _dirty = true;
}

expected<void, std::string> FrustumCuller::update(float) noexcept {
    const size_t count = _index.size();
    _culledCount = 0;
    if (count == 0) return {};
    if (!_ctx) return unexpected<std::string>{"FrustumCuller: no device context"};
    if (dmk_cull_host(_ctx->handle(), &_viewProj[0][0], _nearDepth, _worldInverse.data(), _bounds.data(), static_cast<int64_t>(count), _visible.data()) != DMK_OK)
        return unexpected<std::string>{_ctx->lastError()};
    size_t visible = 0;
    for (size_t w = 0; w < _visible.size(); ++w) {
        uint32_t bits = _visible[w];
        if (w == _visible.size() - 1 && count % 32) bits &= (1u << (count % 32)) - 1u;
        visible += static_cast<size_t>(__builtin_popcount(bits));
    }
    _culledCount = count - visible;
    _dirty = false;
    return {};
}

bool FrustumCuller::shouldEntityBeCulled(Entity entity) noexcept {
    // an entity the filter did not yield is not in _culled: darmok renders it (src/culling.cpp:287-290)
    const auto it = _index.find(entity);
    if (it == _index.end() || _dirty) return false;
    return ((_visible[it->second >> 5] >> (it->second & 31)) & 1u) == 0u;
}

}  // namespace darmok
