// animator_kernel_host.cpp -- see animator_kernel_host.hpp.  The kernel source is compiled as it is for the CPU lockstep
// simulator of tests/devsim (DMK_DEVSIM: one fiber per CUDA thread, blocks one after the other) and its three kernels are
// LAUNCHED through it -- the block-level regrouping of animator_kernel (shared counters, barriers) runs as well as the
// per-character tick.  Built with -ffp-contract=off, the device intrinsics have their IEEE meaning.
#define DMK_DEVSIM 1
#include <dmk_devsim_intrinsics.hpp>

#include <stdexcept>

#include "../../darmok_b200/csrc/animator_flatten.hpp"
#include "../../darmok_b200/csrc/animator_kernel.cu"

#include "animator_kernel_host.hpp"

namespace devsim {

static_assert(sizeof(dmk::AnimEvents) == sizeof(dmk_anim_events) && sizeof(dmk::AnimStateInfo) == sizeof(dmk_anim_state) &&
                  sizeof(dmk::PoseProgram) == sizeof(dmk_pose_program),
              "the C ABI structs are the device structs");

struct Sim::Impl {
    dmk::FlatAnimator flat;
    int64_t n = 0;
    int layers = 0;
    std::vector<uint32_t> slots;
    std::vector<int32_t> transitionDef, cmd, status, clip;
    std::vector<float> transitionTime, cmdSpeed, blend, speed, ratio, weight;
    std::vector<dmk::AnimEvents> events;
    std::vector<dmk::PoseProgram> program;
    dmk::AnimatorParams params(float dt) {
        dmk::AnimatorParams a{};
        a.states = flat.states.data();
        a.clips = flat.clips.data();
        a.transitions = flat.transitions.data();
        a.num_states = static_cast<int>(flat.states.size());
        a.num_transitions = static_cast<int>(flat.transitions.size());
        a.tables = flat.tables.data();
        a.slots = slots.data();
        a.transition_def = transitionDef.data();
        a.transition_time = transitionTime.data();
        a.cmd_state = cmd.data();
        a.cmd_speed = cmdSpeed.data();
        a.in_blend = blend.data();
        a.in_speed = speed.data();
        a.events = events.data();
        a.status = status.data();
        a.n = n;
        a.num_layers = layers;
        a.dt = dt;
        a.clip = clip.data();
        a.ratio = ratio.data();
        a.weight = weight.data();
        a.program = program.data();
        return a;
    }
};

Sim::Sim(const dmk_anim_state_def* states, int numStates, const dmk_anim_clip_def* clips, int numClips, const dmk_anim_transition_def* transitions,
         int numTransitions, const std::vector<float>& clipDurations, int64_t n)
    : _impl(std::make_unique<Impl>()) {
    const std::string msg = dmk::flatten_animator(states, numStates, clips, numClips, transitions, numTransitions, clipDurations, _impl->flat);
    if (!msg.empty()) throw std::runtime_error(msg);
    Impl& s = *_impl;
    s.n = n;
    s.layers = 2 * s.flat.max_clips;
    const size_t N = static_cast<size_t>(n);
    s.slots.assign(static_cast<size_t>(dmk::kNumSlots) * dmk::kSlotWords * N, 0u);
    s.transitionDef.assign(N, 0); s.cmd.assign(N, 0); s.status.assign(N, 0);
    s.transitionTime.assign(N, 0.f); s.cmdSpeed.assign(N, 0.f); s.blend.assign(2 * N, 0.f); s.speed.assign(N, 0.f);
    s.clip.assign(static_cast<size_t>(s.layers) * N, 0); s.ratio.assign(s.clip.size(), 0.f); s.weight.assign(s.clip.size(), 0.f);
    s.events.resize(N);
    std::memset(s.events.data(), 0xff, N * sizeof(dmk::AnimEvents));
    s.program.assign(N, dmk::PoseProgram{});
    dmk::launch_animator_init(s.params(0.f), s.cmdSpeed.data(), s.blend.data(), s.speed.data(), nullptr);
}
Sim::~Sim() = default;

void Sim::commands(const int32_t* state, const float* speed) {
    const size_t N = static_cast<size_t>(_impl->n);
    std::copy(state, state + N, _impl->cmd.begin());
    if (speed) std::copy(speed, speed + N, _impl->cmdSpeed.begin());
    else std::fill(_impl->cmdSpeed.begin(), _impl->cmdSpeed.end(), 1.f);
}
void Sim::setInputs(const float* blendXY, const float* speed) {
    const size_t N = static_cast<size_t>(_impl->n);
    if (blendXY) std::copy(blendXY, blendXY + 2 * N, _impl->blend.begin());
    if (speed) std::copy(speed, speed + N, _impl->speed.begin());
}
void Sim::step(float dt) {
    dmk::launch_animator(_impl->params(dt), nullptr);
}
int Sim::layers() const { return _impl->layers; }
void Sim::getEvents(dmk_anim_events* events, int32_t* status) const {
    const size_t N = static_cast<size_t>(_impl->n);
    if (events) std::memcpy(events, _impl->events.data(), N * sizeof(dmk_anim_events));
    if (status) std::copy(_impl->status.begin(), _impl->status.end(), status);
}
void Sim::getState(dmk_anim_state* out) const {
    dmk::launch_animator_query(_impl->params(0.f), reinterpret_cast<dmk::AnimStateInfo*>(out), nullptr);
}
void Sim::getPrograms(int numLayers, int32_t* clip, float* ratio, float* weight, dmk_pose_program* programs) const {
    const size_t N = static_cast<size_t>(_impl->n), rows = static_cast<size_t>(std::min(numLayers, _impl->layers));
    if (clip) std::copy(_impl->clip.begin(), _impl->clip.begin() + rows * N, clip);
    if (ratio) std::copy(_impl->ratio.begin(), _impl->ratio.begin() + rows * N, ratio);
    if (weight) std::copy(_impl->weight.begin(), _impl->weight.begin() + rows * N, weight);
    if (programs) std::memcpy(programs, _impl->program.data(), N * sizeof(dmk_pose_program));
}

}  // namespace devsim
