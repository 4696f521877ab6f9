// animator_kernel_host.hpp -- the device animator's per-character code (darmok_b200/csrc/animator_kernel.cu) compiled for the
// CPU, behind the same calls the C ABI offers (dmk_crowd_set_animator / _animator_commands / _animator_set_inputs /
// dmk_crowd_step(DMK_STEP_ANIMATOR) / the getters).  TEST INFRASTRUCTURE: lets tests/cpp/animator_test.cpp run thousands of
// scripted characters through the kernel's state machine without a GPU and compare them with darmok's animator.
#pragma once

#include "dmk.h"

#include <cstdint>
#include <memory>
#include <string>
#include <vector>

namespace devsim {

class Sim {
public:
    // the arguments of dmk_crowd_set_animator plus the clip set's durations; throws std::runtime_error with its message
    Sim(const dmk_anim_state_def* states, int numStates, const dmk_anim_clip_def* clips, int numClips, const dmk_anim_transition_def* transitions,
        int numTransitions, const std::vector<float>& clipDurations, int64_t n);
    ~Sim();
    void commands(const int32_t* state, const float* speed);          // dmk_crowd_animator_commands
    void setInputs(const float* blendXY, const float* speed);         // dmk_crowd_animator_set_inputs
    void step(float dt);                                              // the animator kernel's tick for every character
    int layers() const;                                               // rows of the layer arrays (2 * largest clip count)
    void getEvents(dmk_anim_events* events, int32_t* status) const;   // dmk_crowd_get_animator_events
    void getState(dmk_anim_state* out) const;                         // dmk_crowd_get_animator_state
    void getPrograms(int numLayers, int32_t* clip, float* ratio, float* weight, dmk_pose_program* programs) const;   // dmk_crowd_get_programs
private:
    struct Impl;
    std::unique_ptr<Impl> _impl;
};

}  // namespace devsim
