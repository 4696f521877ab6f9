// darmok_b200/protobuf_adapter.hpp -- protobuf::SkeletalAnimator / protobuf::Armature -> the shim's plain definitions.
//
// darmok's SkeletalAnimator::Definition is the protoc-generated class darmok::protobuf::SkeletalAnimator
// (include/darmok/skeleton.hpp:262-263, assets/protobuf/skeleton.proto:19-57); the shim keeps plain structs with the same field
// names so that it does not depend on protobuf.  These templates are the whole adapter: they only use the accessors protoc
// generates for that .proto (name(), animations(), blend_position().x(), tween().easing(), ...), so they compile against the
// real generated header inside darmok and against any type with the same accessors (tests/cpp/boundary_test.cpp).
#pragma once

#include "skeleton.hpp"

namespace darmok {

template <class ProtoTween>
SkeletalAnimatorTweenDefinition tweenFromProtobuf(const ProtoTween& t) {
    SkeletalAnimatorTweenDefinition out;
    out.easing = static_cast<EasingType>(static_cast<int>(t.easing()));   // protobuf::Easing::Type has EasingType's values (easing.proto)
    out.duration = t.duration();
    return out;
}

template <class ProtoAnimator>
SkeletalAnimatorDefinition fromProtobuf(const ProtoAnimator& def) {
    SkeletalAnimatorDefinition out;
    out.skeleton_path = def.skeleton_path();
    out.animation_pattern = def.animation_pattern();
    for (const auto& s : def.states()) {
        SkeletalAnimatorStateDefinition state;
        state.name = s.name();
        for (const auto& a : s.animations()) {
            SkeletalAnimatorAnimationDefinition anim;
            anim.name = a.name();
            anim.blend_position = glm::vec2{a.blend_position().x(), a.blend_position().y()};
            anim.speed = a.speed();
            anim.loop = a.loop();
            state.animations.push_back(std::move(anim));
        }
        state.blend = static_cast<SkeletalAnimatorStateDefinition::BlendType>(static_cast<int>(s.blend()));
        state.threshold = s.threshold();
        state.tween = tweenFromProtobuf(s.tween());
        state.next_state = s.next_state();
        state.speed = s.speed();
        out.states.push_back(std::move(state));
    }
    for (const auto& t : def.transitions()) {
        SkeletalAnimatorTransitionDefinition tr;
        tr.src_state = t.src_state();
        tr.dst_state = t.dst_state();
        tr.tween = tweenFromProtobuf(t.tween());
        tr.offset = t.offset();
        out.transitions.push_back(std::move(tr));
    }
    return out;
}

// protobuf::Armature (assets/protobuf/skeleton.proto:8-16): inverse_bind_pose is a Mat4 of four column vectors col1..col4
// (assets/protobuf/math.proto:41-46)
template <class ProtoArmature>
ArmatureDefinition armatureFromProtobuf(const ProtoArmature& def) {
    ArmatureDefinition out;
    out.name = def.name();
    for (const auto& j : def.joints()) {
        ArmatureJointDefinition joint;
        joint.name = j.name();
        const auto& m = j.inverse_bind_pose();
        const auto col = [&](int c, const auto& v) {
            joint.inverse_bind_pose[c][0] = v.x();
            joint.inverse_bind_pose[c][1] = v.y();
            joint.inverse_bind_pose[c][2] = v.z();
            joint.inverse_bind_pose[c][3] = v.w();
        };
        col(0, m.col1());
        col(1, m.col2());
        col(2, m.col3());
        col(3, m.col4());
        out.joints.push_back(std::move(joint));
    }
    return out;
}

}  // namespace darmok
