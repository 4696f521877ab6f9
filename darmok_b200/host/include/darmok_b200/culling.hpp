// darmok_b200/culling.hpp -- darmok's FrustumCuller over the B200 C ABI (include/dmk.h dmk_cull_*).
//
// Mirrors FrustumCuller of include/darmok/culling.hpp:53-71 / src/culling.cpp:222-290: an ICameraComponent that, every
// update(), tests EVERY entity the camera sees that has a Renderable and a BoundingBox (filter src/culling.cpp:36-44) --
// animated or not -- against the camera frustum taken into the entity's local space (`frust *= trans->getWorldInverse()`,
// :272-275; an entity without a Transform uses the camera frustum as it is, i.e. an identity world inverse), and answers
// shouldEntityBeCulled(entity) from the set it built (:287-290).  Here the per-entity loop is ONE batched kernel launch
// (K4 cull_kernel, visibility bit-exact with the reference arithmetic: tests/test_gpu_cull.py).
//
// What darmok's class pulls out of Camera / Scene itself -- the entity list of the filter, each entity's world inverse and
// bounds, the camera's view-projection matrix -- is handed in by the caller (setEntities / setViewProjection): the darmok-side
// glue that fills them from EnTT is in INTEGRATION.md.  SkeletalAnimationSceneComponent (skeleton.hpp) keeps its own fused
// cull for the animated crowd; this class is for everything else that is Renderable.
#pragma once

#include "expected.hpp"
#include "glm_lite.hpp"

#include <cstddef>
#include <cstdint>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

namespace darmok {

class DeviceContext;
using Entity = uint32_t;   // entt::entity in darmok (include/darmok/scene_fwd.hpp)

class FrustumCuller final {
public:
    struct Definition {};   // protobuf::FrustumCuller has no fields (assets/protobuf/camera.proto:32-33)
    static Definition createDefinition() noexcept { return {}; }

    explicit FrustumCuller(std::shared_ptr<DeviceContext> ctx) noexcept;
    ~FrustumCuller() noexcept;

    expected<void, std::string> init() noexcept;                    // FrustumCuller::init(cam, scene, app): nothing to acquire here
    expected<void, std::string> load(const Definition& def) noexcept;
    expected<void, std::string> shutdown() noexcept;                // forgets the entities and the culled set

    // the entities of the camera's filter, in any order: world inverse = Transform::getWorldInverse() as column-major mat4
    // ([n][16]; pass nullptr when no entity has a Transform), bounds = BoundingBox min.xyz, max.xyz ([n][6])
    expected<void, std::string> setEntities(const Entity* entities, const float* worldInverse, const float* bounds, size_t count) noexcept;
    void setViewProjection(const glm::mat4& viewProj, float normalizedNearDepth = 0.f) noexcept;   // _cam->getViewProjectionMatrix()

    expected<void, std::string> update(float deltaTime) noexcept;   // updateCulled(): one batched launch
    bool shouldEntityBeCulled(Entity entity) noexcept;              // _culled.contains(entity)
    size_t culledCount() const noexcept { return _culledCount; }

private:
    std::shared_ptr<DeviceContext> _ctx;
    std::unordered_map<Entity, size_t> _index;
    std::vector<float> _worldInverse, _bounds;
    std::vector<uint32_t> _visible;   // bit i: entity i is inside the frustum
    glm::mat4 _viewProj{1.f};
    float _nearDepth = 0.f;
    size_t _culledCount = 0;
    bool _dirty = true;
};

}  // namespace darmok
