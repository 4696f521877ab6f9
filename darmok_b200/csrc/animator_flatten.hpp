// animator_flatten.hpp -- host-only: a flattened SkeletalAnimatorDefinition (dmk.h dmk_anim_*_def) turned into the tables the
// animator kernel reads (device_types.cuh): clip durations resolved, OzzSkeletalAnimatorState::calcDuration per state, the rest
// clip and, for Directional states, the clip-pair constants of calcBlendWeight.  Used by dmk_crowd_set_animator (api_crowd.cu) and
// by tests/cpp/animator_kernel_host.cpp, which runs the kernel's per-character code on the CPU.
#pragma once

#include "../../include/dmk.h"
#include "device_types.cuh"

#include <algorithm>
#include <cmath>
#include <numeric>
#include <string>
#include <vector>

namespace dmk {

// OzzSkeletalAnimatorState::calcDuration (src/skeleton_ozz.cpp:487-518): least common multiple of the clip durations in
// decimal fixed point.  For a duration without a fractional part the reference converts ceil(-log10(0)) = +inf to int,
// which x86 turns into INT_MIN so that pow(10, INT_MIN) = 0 and the factor stays 1; that outcome is kept here.
inline float anim_state_duration(const std::vector<float>& clip_durations) {
    if (clip_durations.empty()) return 0.f;
    int f = 1;
    for (float d : clip_durations) {
        const float frac = d - std::floor(d);
        if (!(frac > 0.f)) continue;
        const int places = static_cast<int>(std::ceil(-std::log10(frac)));
        f = std::max(f, static_cast<int>(std::pow(10, places)));
    }
    int d = static_cast<int>(std::round(clip_durations[0] * f));
    for (size_t i = 1; i < clip_durations.size(); ++i) d = std::lcm(d, static_cast<int>(std::round(clip_durations[i] * f)));
    return static_cast<float>(d) / f;
}

// glm::length / glm::angle as calcBlendWeight calls them (src/skeleton.cpp:312-325): no normalisation before the acos
inline float anim_length(float x, float y) { return std::sqrt(x * x + y * y); }
inline float anim_angle(float ax, float ay, float bx, float by) { return std::acos(std::min(std::max(ax * bx + ay * by, -1.f), 1.f)); }


struct FlatAnimator {
    std::vector<AnimStateDef> states;
    std::vector<AnimClipDef> clips;
    std::vector<AnimTransitionDef> transitions;
    std::vector<float> tables;
    int max_clips = 0;   // largest number of clips of a state: a transition needs 2 * max_clips layer rows
};

// "" on success, else the message dmk_crowd_set_animator fails with.  clip_durations = durations of the clip set's clips.
inline std::string flatten_animator(const dmk_anim_state_def* states, int num_states, const dmk_anim_clip_def* clips, int num_clips,
                                    const dmk_anim_transition_def* transitions, int num_transitions, const std::vector<float>& clip_durations,
                                    FlatAnimator& out) {
    const int num_set_clips = static_cast<int>(clip_durations.size());
    std::vector<AnimClipDef> hc(static_cast<size_t>(num_clips));
    for (int i = 0; i < num_clips; ++i) {
        if (clips[i].clip < 0 || clips[i].clip >= num_set_clips)
            return "animation " + std::to_string(i) + " references a clip outside the clip set";
        hc[i] = AnimClipDef{clips[i].clip, clips[i].blend_x, clips[i].blend_y, clips[i].speed, clips[i].loop ? 1 : 0,
                            clip_durations[static_cast<size_t>(clips[i].clip)]};
    }
    std::vector<AnimStateDef> hs(static_cast<size_t>(num_states));
    std::vector<float> tables(1, 0.f);   // never empty, so the upload always yields a pointer
    int max_clips = 0;
    for (int i = 0; i < num_states; ++i) {
        const dmk_anim_state_def& d = states[i];
        if (d.num_clips < 1 || d.num_clips > kMaxStateClips || d.first_clip < 0 || d.first_clip + d.num_clips > num_clips)
            return "state " + std::to_string(i) + ": clip range must hold 1.." + std::to_string(kMaxStateClips) + " animations";
        if (d.next_state < -1 || d.next_state >= num_states || d.blend_type < 0 || d.blend_type > 1 || d.tween_easing < 0 || d.tween_easing > 31)
            return "state " + std::to_string(i) + ": next_state, blend_type or easing out of range";
        max_clips = std::max(max_clips, d.num_clips);
        std::vector<float> durations;
        int rest = -1;
        for (int k = 0; k < d.num_clips; ++k) {
            const AnimClipDef& cd = hc[static_cast<size_t>(d.first_clip + k)];
            const float speed = cd.speed == 0.f ? 1.f : cd.speed;
            durations.push_back(cd.duration / speed);
            if (cd.blend_x == 0.f && cd.blend_y == 0.f) rest = k;
        }
        int table = -1;
        if (d.blend_type == 1) {   // Directional: the clip-pair constants of the gradient bands
            table = static_cast<int>(tables.size());
            const AnimClipDef* sc = &hc[static_cast<size_t>(d.first_clip)];
            for (int k = 0; k < d.num_clips; ++k) tables.push_back(anim_length(sc[k].blend_x, sc[k].blend_y));
            for (int a = 0; a < d.num_clips; ++a)
                for (int b = 0; b < d.num_clips; ++b) tables.push_back(anim_angle(sc[b].blend_x, sc[b].blend_y, sc[a].blend_x, sc[a].blend_y));
        }
        hs[i] = AnimStateDef{d.first_clip, d.num_clips, d.blend_type, d.threshold, d.tween_easing, d.tween_duration, d.next_state, d.speed,
                             anim_state_duration(durations), rest < 0 ? 0 : rest, table};
    }
    std::vector<AnimTransitionDef> ht(static_cast<size_t>(num_transitions));
    for (int i = 0; i < num_transitions; ++i) {
        const dmk_anim_transition_def& t = transitions[i];
        if (t.src_state < -1 || t.src_state >= num_states || t.dst_state < -1 || t.dst_state >= num_states || t.easing < 0 || t.easing > 31)
            return "transition " + std::to_string(i) + ": state or easing out of range";
        ht[i] = AnimTransitionDef{t.src_state, t.dst_state, t.easing, t.duration, t.offset};
    }
    out.states = std::move(hs);
    out.clips = std::move(hc);
    out.transitions = std::move(ht);
    out.tables = std::move(tables);
    out.max_clips = max_clips;
    return std::string();
}

}  // namespace dmk
