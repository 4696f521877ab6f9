// animator_kernel.cu -- SURVEY.md 8(f) rank 2: darmok's animator state machine for a whole crowd, one thread per character.
//
// Device port of the host shim's animator_core.cpp, which mirrors (paths relative to the darmok repository):
//   OzzSkeletalAnimatorAnimationState::update   src/skeleton_ozz.cpp:411-428   clip clocks
//   OzzSkeletalAnimatorState::update            src/skeleton_ozz.cpp:535-610   state clock, tweened blend position, gradient-band weights
//   calcBlendWeight(s)                          src/skeleton.cpp:306-364
//   OzzSkeletalAnimatorTransition::update       src/skeleton_ozz.cpp:682-726
//   SkeletalAnimatorImpl::play / update / afterUpdate   src/skeleton_ozz.cpp:783-856, 1006-1087
// A tick resolves every character into its layer rows (clip, ratio, weight) and its PoseProgram, which the sampling and
// model kernels then execute; listener events come back as one AnimEvents record per character.
//
// State is structure-of-arrays (device_types.cuh): lane c of a warp reads word w of its character at slots[.. * N + c],
// so loads and stores coalesce and a character touches only the words its situation needs.
// Compiled with -fmad=false: clocks, tweens and Cartesian weights are bit-identical to the host; the Directional
// weights' per-definition angles come precomputed from the host, the per-tick angle(position, clip) uses CUDA's acosf
// and the transcendental easing curves CUDA's sinf/cosf/powf (last-ulp differences from glibc are possible).
#include "device_intrinsics.cuh"
#include "device_types.cuh"
#include "kernels.cuh"

#include "../host/include/darmok_b200/easing.hpp"

#include <cfloat>

namespace dmk {

namespace {

using darmok::EasingType;
using darmok::ease01;

// one slot of one character
struct Slot {
    uint32_t* w;   // &slots[(slot * kSlotWords) * n + c]
    int64_t n;
    __device__ __forceinline__ float getf(int word) const { return __uint_as_float(w[(int64_t)word * n]); }
    __device__ __forceinline__ void setf(int word, float v) const { w[(int64_t)word * n] = __float_as_uint(v); }
    __device__ __forceinline__ int geti(int word) const { return (int)w[(int64_t)word * n]; }
    __device__ __forceinline__ void seti(int word, int v) const { w[(int64_t)word * n] = (uint32_t)v; }
};
__device__ __forceinline__ Slot slot_of(const AnimatorParams& p, int slot, int64_t c) {
    return Slot{p.slots + (int64_t)slot * kSlotWords * p.n + c, p.n};
}

__device__ void state_create(const Slot& s, int def, const AnimStateDef* states) {
    s.seti(kWDef, def);
    s.seti(kWClipLooped, 0);
    s.setf(kWBlendX, 0.f); s.setf(kWBlendY, 0.f); s.setf(kWOldBlendX, 0.f); s.setf(kWOldBlendY, 0.f);
    s.setf(kWNormalizedTime, 0.f);
    s.setf(kWTweenTime, 0.f);
    s.setf(kWSpeed, states[def].speed);
    s.seti(kWLooped, 0);
    const int m = states[def].num_clips;
    for (int k = 0; k < m; ++k) s.setf(kWClipTime + k, 0.f);
}
__device__ void state_copy(const Slot& dst, const Slot& src, const AnimStateDef* states) {
    const int def = src.geti(kWDef);
    const int words = kWClipTime + (def >= 0 ? states[def].num_clips : 0);
    for (int k = 0; k < words; ++k) dst.w[(int64_t)k * dst.n] = src.w[(int64_t)k * src.n];
}

__device__ bool state_finished(const Slot& s, const AnimStateDef& d, const AnimClipDef* clips) {
    const uint32_t looped = (uint32_t)s.geti(kWClipLooped);
    for (int k = 0; k < d.num_clips; ++k)
        if (!(((looped >> k) & 1u) && !clips[d.first_clip + k].loop)) return false;
    return true;
}

__device__ __forceinline__ float vec2_length(float x, float y) { return __fsqrt_rn(x * x + y * y); }
__device__ __forceinline__ float vec2_angle(float ax, float ay, float bx, float by) {   // glm::angle without normalisation, as called
    float d = ax * bx + ay * by;
    d = fminf(fmaxf(d, -1.f), 1.f);
    return acosf(d);
}

// calcBlendWeight (src/skeleton.cpp:306-344) of clip i for the blend position (px, py)
__device__ float blend_weight(const AnimStateDef& d, const AnimClipDef* clips, const float* tables, float px, float py, float len, int i) {
    const int m = d.num_clips;
    const float ax = clips[d.first_clip + i].blend_x, ay = clips[d.first_clip + i].blend_y;
    float w = FLT_MAX;
    if (d.blend_type == 1) {
        const float* lens = tables + d.table;
        const float* angles = lens + m + i * m;
        const float len_anim = lens[i];
        const float a1 = vec2_angle(px, py, ax, ay);
        for (int j = 0; j < m; ++j) {
            const float len_anim2 = lens[j];
            const float p = len_anim + len_anim2;
            const float a0 = (len - len_anim) * p;
            const float b0 = (len_anim2 - len_anim) * p, b1 = angles[j];
            const float f = 1.F - __fdiv_rn(a0 * b0 + a1 * b1, b0 * b0 + b1 * b1);
            if (f < w) w = f;
        }
    } else {
        const float a0 = px - ax, a1 = py - ay;
        for (int j = 0; j < m; ++j) {
            const float b0 = clips[d.first_clip + j].blend_x - ax, b1 = clips[d.first_clip + j].blend_y - ay;
            const float f = 1.F - __fdiv_rn(a0 * b0 + a1 * b1, b0 * b0 + b1 * b1);
            if (f < w) w = f;   // the j == i term is 0/0 = NaN and drops out
        }
    }
    return w;
}

struct GroupOut {
    int count, rest, passthrough;
    int zero;   // passthrough of a clip that has never been sampled: darmok's zero-initialised locals (kGroupZero)
    float threshold;
};

// OzzSkeletalAnimatorState::update; writes the group's layer rows [first, first + count).  false = "error in the blending job".
// commit = false evaluates without touching the stored state (the frozen pose of a finished transition).
__device__ bool state_update(const AnimatorParams& p, int64_t c, const Slot& s, float dt, float target_x, float target_y, int first,
                             GroupOut& g, bool commit = true) {
    const AnimStateDef& d = p.states[s.geti(kWDef)];
    dt *= s.getf(kWSpeed);
    uint32_t clip_looped = (uint32_t)s.geti(kWClipLooped);
    for (int k = 0; k < d.num_clips; ++k) {   // OzzSkeletalAnimatorAnimationState::update, minus the sampling
        const AnimClipDef& cd = p.clips[d.first_clip + k];
        float t = s.getf(kWClipTime + k);
        const bool looped = (clip_looped >> k) & 1u;
        if (!(looped && !cd.loop)) {
            float speed = cd.speed;
            speed = speed == 0.f ? 1.f : speed;
            const float duration = __fdiv_rn(cd.duration, speed);
            const float norm_time = t + __fdiv_rn(dt, duration);
            const bool now_looped = norm_time > 1.f;
            clip_looped = (clip_looped & ~(1u << k)) | ((now_looped ? 1u : 0u) << k);
            if (!(now_looped && !cd.loop)) {
                t = fmodf(norm_time, 1.f);
                if (commit) s.setf(kWClipTime + k, t);
                if (commit) clip_looped |= 0x10000u << k;   // SamplingJob has run for this clip at least once (bits 16.. of the word)
            }
        }
        const int64_t at = (int64_t)(first + k) * p.n + c;
        p.clip[at] = cd.clip;
        p.ratio[at] = t;
        p.weight[at] = 0.f;
    }
    if (commit) {
        s.seti(kWClipLooped, (int)clip_looped);
        float nt = s.getf(kWNormalizedTime) + __fdiv_rn(dt, d.duration);
        s.seti(kWLooped, nt > 1.f ? 1 : 0);
        s.setf(kWNormalizedTime, fmodf(nt, 1.f));
    }

    g.count = d.num_clips;
    g.threshold = d.threshold;
    g.rest = 0;
    g.passthrough = 1;
    g.zero = ((clip_looped >> 16) & 1u) ? 0 : 1;
    if (d.num_clips <= 1) return true;
    const EasingType easing = static_cast<EasingType>(d.tween_easing);
    float blend_x = s.getf(kWBlendX), blend_y = s.getf(kWBlendY), old_x = s.getf(kWOldBlendX), old_y = s.getf(kWOldBlendY);
    float tween = s.getf(kWTweenTime);
    if (blend_x != target_x || blend_y != target_y) {
        const float f = ease01(easing, tween);
        old_x = old_x + f * (blend_x - old_x);
        old_y = old_y + f * (blend_y - old_y);
        tween = 0.f;
        blend_x = target_x;
        blend_y = target_y;
        if (commit) {
            s.setf(kWOldBlendX, old_x); s.setf(kWOldBlendY, old_y);
            s.setf(kWBlendX, blend_x); s.setf(kWBlendY, blend_y);
        }
    }
    tween += __fdiv_rn(dt, d.tween_duration);
    if (tween >= 1.f) tween = 1.f;
    if (commit) s.setf(kWTweenTime, tween);
    const float f = ease01(easing, tween);
    const float px = old_x + f * (blend_x - old_x), py = old_y + f * (blend_y - old_y);
    const float len = d.blend_type == 1 ? vec2_length(px, py) : 0.f;
    for (int k = 0; k < d.num_clips; ++k) {
        const AnimClipDef& cd = p.clips[d.first_clip + k];
        if (cd.blend_x == px && cd.blend_y == py) {   // early exit: a clip sits exactly at the blend position
            g.rest = k;
            g.zero = ((clip_looped >> (16 + k)) & 1u) ? 0 : 1;
            return true;
        }
        p.weight[(int64_t)(first + k) * p.n + c] = blend_weight(d, p.clips, p.tables, px, py, len, k);
    }
    g.passthrough = 0;
    g.zero = 0;
    g.rest = d.rest_clip;
    return d.threshold > 0.f;   // BlendingJob::Validate
}

__device__ int find_transition(const AnimatorParams& p, int src, int dst) {
    // src/skeleton.cpp:575-599: exact, src>*, *>dst, *>*
    for (int i = 0; i < p.num_transitions; ++i) if (p.transitions[i].src_state == src && p.transitions[i].dst_state == dst) return i;
    for (int i = 0; i < p.num_transitions; ++i) if (p.transitions[i].src_state == src && p.transitions[i].dst_state < 0) return i;
    for (int i = 0; i < p.num_transitions; ++i) if (p.transitions[i].src_state < 0 && p.transitions[i].dst_state == dst) return i;
    for (int i = 0; i < p.num_transitions; ++i) if (p.transitions[i].src_state < 0 && p.transitions[i].dst_state < 0) return i;
    return -1;
}

struct Character {   // SkeletalAnimatorImpl's scalars, in registers for the tick
    int state_def;        // _state's definition (-1 = none); mirrors slot kSlotState's kWDef
    int transition_def;   // -1 = no transition
    float transition_time;
};

// SkeletalAnimatorImpl::play (src/skeleton_ozz.cpp:783-856)
__device__ void play(const AnimatorParams& p, int64_t c, Character& a, int state, float speed, int16_t& started, int16_t& transition_started) {
    const Slot st = slot_of(p, kSlotState, c), prev = slot_of(p, kSlotPrev, c), cur = slot_of(p, kSlotCur, c);
    if (a.transition_def >= 0) {
        if (cur.geti(kWDef) == state) { cur.setf(kWSpeed, speed); return; }
    } else if (a.state_def >= 0 && a.state_def == state) {
        st.setf(kWSpeed, speed);
        return;
    }
    if (state < 0 || state >= p.num_states) return;
    // the source of a possible transition: _state, overridden by a running transition's current state
    int prev_def = -1;
    bool prev_in_cur = false;
    if (a.state_def >= 0) prev_def = a.state_def;
    if (a.transition_def >= 0) { prev_def = cur.geti(kWDef); prev_in_cur = true; }
    const int t = prev_def >= 0 ? find_transition(p, prev_def, state) : -1;
    if (t >= 0) state_copy(prev, prev_in_cur ? cur : st, p.states);
    a.state_def = -1;
    st.seti(kWDef, -1);
    a.transition_def = -1;
    if (t >= 0) {
        state_create(cur, state, p.states);
        a.transition_def = t;
        a.transition_time = 0.f;
        started = (int16_t)state;
        transition_started = 1;
        return;
    }
    state_create(st, state, p.states);
    a.state_def = state;
    started = (int16_t)state;
}

__device__ __forceinline__ void set_group(PoseProgram& prog, int k, const GroupOut& g) {
    const uint8_t mode = g.zero ? kGroupZero : g.passthrough ? kGroupPassthrough : kGroupBlend;
    if (k == 0) { prog.n0 = (uint8_t)g.count; prog.rest0 = (uint8_t)g.rest; prog.mode0 = mode; prog.threshold0 = g.threshold; }
    else { prog.n1 = (uint8_t)g.count; prog.rest1 = (uint8_t)g.rest; prog.mode1 = mode; prog.threshold1 = g.threshold; }
}

// Characters of one block are regrouped by situation (state, or transition into a state, or a pending command) before the
// tick so that a warp's lanes run the same control flow: with one thread per character in id order a warp executed the
// union of ~4 situations at 4.9 active lanes per instruction (ncu).  The regrouping stays inside the block's window of
// consecutive ids, so the structure-of-arrays accesses still fall into the same few cache lines.
constexpr int kAnimBlock = 256;
constexpr int kAnimKeys = 64;

__device__ __forceinline__ void tick_character(const AnimatorParams& p, int64_t c);

__global__ void __launch_bounds__(kAnimBlock) animator_kernel(const AnimatorParams p) {
    __shared__ int key_count[kAnimKeys];
    __shared__ uint16_t order[kAnimBlock];
    const int64_t block_base = (int64_t)blockIdx.x * kAnimBlock;
    {
        if (threadIdx.x < kAnimKeys) key_count[threadIdx.x] = 0;
        __syncthreads();
        const int64_t mine = block_base + threadIdx.x;
        int key = kAnimKeys - 1;   // past the end of the crowd: sorted last, exits below
        if (mine < p.n) {
            const int t = p.transition_def[mine];
            const int def = slot_of(p, t >= 0 ? kSlotCur : kSlotState, mine).geti(kWDef);
            if (p.cmd_state[mine] != -2) key = kAnimKeys - 2;                      // commands restructure: keep them together
            else if (def < 0) key = 0;                                             // nothing playing
            else key = 1 + min(2 * def + (t >= 0 ? 1 : 0), kAnimKeys - 4);         // (state, in transition)
        }
        const int rank = atomicAdd(&key_count[key], 1);
        __syncthreads();
        if (threadIdx.x == 0) {   // exclusive scan of 64 counters
            int sum = 0;
            for (int k = 0; k < kAnimKeys; ++k) { const int v = key_count[k]; key_count[k] = sum; sum += v; }
        }
        __syncthreads();
        order[key_count[key] + rank] = (uint16_t)threadIdx.x;
        __syncthreads();
    }
    const int64_t c = block_base + order[threadIdx.x];
    if (c >= p.n) return;
    tick_character(p, c);
}

// One SkeletalAnimatorImpl tick: pending command, update, afterUpdate.  (A function of its own so that
// tests/cpp/animator_kernel_host.cpp can compile this file for the CPU and run it next to the host state machine.)
__device__ __forceinline__ void tick_character(const AnimatorParams& p, int64_t c) {
    const Slot st = slot_of(p, kSlotState, c), prev = slot_of(p, kSlotPrev, c), cur = slot_of(p, kSlotCur, c);
    Character a;
    a.state_def = st.geti(kWDef);
    a.transition_def = p.transition_def[c];
    a.transition_time = a.transition_def >= 0 ? p.transition_time[c] : 0.f;
    const float blend_x = p.in_blend[2 * c], blend_y = p.in_blend[2 * c + 1];
    AnimEvents ev;
    ev.cmd_started = -1; ev.cmd_transition_started = 0; ev.transition_finished = 0; ev.prev_finished = -1;
    ev.looped = -1; ev.finished = -1; ev.auto_started = -1; ev.auto_transition_started = 0;

    // pending command (SkeletalAnimator::play / stop issued since the last tick)
    const int cmd = p.cmd_state[c];
    if (cmd != -2) {
        if (cmd >= 0) play(p, c, a, cmd, p.cmd_speed[c], ev.cmd_started, ev.cmd_transition_started);
        else if (cmd == -1) { a.state_def = -1; st.seti(kWDef, -1); }   // stop(): clears _state only; a running transition keeps going (C10)
        else if (cmd == -3) { a.state_def = -1; st.seti(kWDef, -1); a.transition_def = -1; }   // play() of a state that cannot be created (:793-845)
        p.cmd_state[c] = -2;
    }

    // SkeletalAnimatorImpl::update
    const float dt = p.dt * p.in_speed[c];
    PoseProgram prog;
    prog.n0 = prog.n1 = prog.rest0 = prog.rest1 = 0;
    prog.mode0 = prog.mode1 = kGroupPassthrough;
    prog.top = kTopSkip;
    prog.reserved = 0;
    prog.threshold0 = prog.threshold1 = 0.f;
    prog.w0 = 1.f; prog.w1 = 0.f;
    int status = 0;
    bool updated = false;
    if (a.transition_def >= 0) {
        const AnimTransitionDef& td = p.transitions[a.transition_def];
        updated = true;
        GroupOut g0, g1;
        if (!(a.transition_time >= 1.f)) {
            if (!state_update(p, c, prev, dt, blend_x, blend_y, 0, g0)) {
                status = 1;   // the previous state's failure is returned (:688-692)
                updated = false;
            } else {
                float cur_dt = dt;
                if (a.transition_time == 0.f) cur_dt += td.offset;
                if (state_update(p, c, cur, cur_dt, blend_x, blend_y, g0.count, g1)) {
                    a.transition_time += __fdiv_rn(dt, td.duration);
                    const float v = ease01(static_cast<EasingType>(td.easing), a.transition_time);
                    prog.top = kTopTransition;
                    set_group(prog, 0, g0);
                    set_group(prog, 1, g1);
                    prog.w0 = 1.f - v;
                    prog.w1 = v;
                }   // else: the current state's failure is swallowed (C8): no job runs, matrices stay
            }
        } else if (state_update(p, c, cur, 0.f, blend_x, blend_y, 0, g0, /*commit=*/false)) {
            // already finished at entry (afterUpdate retires a finished transition, so this needs a failed tick in
            // between): the current state's pose, its clocks no longer ticking
            prog.top = kTopGroup0;
            set_group(prog, 0, g0);
        }
    } else if (a.state_def >= 0) {
        GroupOut g0;
        if (state_update(p, c, st, dt, blend_x, blend_y, 0, g0)) {
            updated = true;
            prog.top = kTopGroup0;
            set_group(prog, 0, g0);
        } else {
            status = 1;
        }
    }

    // afterUpdate (src/skeleton_ozz.cpp:1044-1087): only after a successful update of a transition or a state
    if (updated) {
        if (a.transition_def >= 0 && a.transition_time >= 1.f) {
            ev.transition_finished = 1;
            ev.prev_finished = (int16_t)prev.geti(kWDef);
            state_copy(st, cur, p.states);   // (_state, if any, would first get afterFinished(); it is overwritten right away)
            a.state_def = cur.geti(kWDef);
            a.transition_def = -1;
        }
        if (a.state_def >= 0) {
            const AnimStateDef& d = p.states[a.state_def];
            if (st.geti(kWLooped)) ev.looped = (int16_t)a.state_def;
            if (state_finished(st, d, p.clips)) {
                ev.finished = (int16_t)a.state_def;
                st.setf(kWSpeed, d.speed);   // afterFinished
                if (d.next_state >= 0) play(p, c, a, d.next_state, 1.f, ev.auto_started, ev.auto_transition_started);
            }
        }
    }
    p.transition_def[c] = a.transition_def;
    if (a.transition_def >= 0) p.transition_time[c] = a.transition_time;
    p.program[c] = prog;
    p.events[c] = ev;
    p.status[c] = status;
}

__device__ __forceinline__ void query_character(const AnimatorParams& p, int64_t c, AnimStateInfo* out);
__global__ void animator_query_kernel(const AnimatorParams p, AnimStateInfo* out) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= p.n) return;
    query_character(p, c, out);
}
__device__ __forceinline__ void query_character(const AnimatorParams& p, int64_t c, AnimStateInfo* out) {
    const int t = p.transition_def[c];
    const Slot s = slot_of(p, t >= 0 ? kSlotCur : kSlotState, c);   // getCurrentState (src/skeleton_ozz.cpp:867-878)
    const Slot pv = slot_of(p, kSlotPrev, c);
    AnimStateInfo o;
    o.state = s.geti(kWDef);
    o.transition = t;
    o.previous_state = t >= 0 ? pv.geti(kWDef) : -1;
    o.state_time = o.state >= 0 ? s.getf(kWNormalizedTime) : 0.f;
    o.transition_time = t >= 0 ? p.transition_time[c] : 0.f;
    o.previous_time = t >= 0 ? pv.getf(kWNormalizedTime) : 0.f;
    o.state_speed = o.state >= 0 ? s.getf(kWSpeed) : 0.f;
    o.previous_speed = t >= 0 ? pv.getf(kWSpeed) : 0.f;
    out[c] = o;
}

__device__ __forceinline__ void init_character(const AnimatorParams& p, int64_t c, float* cmd_speed, float* in_blend, float* in_speed);
__global__ void animator_init_kernel(const AnimatorParams p, float* cmd_speed, float* in_blend, float* in_speed) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= p.n) return;
    init_character(p, c, cmd_speed, in_blend, in_speed);
}
__device__ __forceinline__ void init_character(const AnimatorParams& p, int64_t c, float* cmd_speed, float* in_blend, float* in_speed) {
    for (int s = 0; s < kNumSlots; ++s) slot_of(p, s, c).seti(kWDef, -1);
    p.transition_def[c] = -1;
    p.transition_time[c] = 0.f;
    p.cmd_state[c] = -2;
    cmd_speed[c] = 1.f;
    in_blend[2 * c] = in_blend[2 * c + 1] = 0.f;
    in_speed[c] = 1.f;
}

}  // namespace

cudaError_t launch_animator_init(const AnimatorParams& p, float* cmd_speed, float* in_blend, float* in_speed, cudaStream_t stream) {
    if (p.n <= 0) return cudaSuccess;
    DMK_LAUNCH(animator_init_kernel, (unsigned)((p.n + 255) / 256), 256, 0, stream)(p, cmd_speed, in_blend, in_speed);
    return cudaGetLastError();
}

cudaError_t launch_animator_query(const AnimatorParams& p, AnimStateInfo* out, cudaStream_t stream) {
    if (p.n <= 0) return cudaSuccess;
    DMK_LAUNCH(animator_query_kernel, (unsigned)((p.n + 255) / 256), 256, 0, stream)(p, out);
    return cudaGetLastError();
}

cudaError_t launch_animator(const AnimatorParams& p, cudaStream_t stream) {
    if (p.n <= 0) return cudaSuccess;
    DMK_LAUNCH(animator_kernel, (unsigned)((p.n + kAnimBlock - 1) / kAnimBlock), kAnimBlock, 0, stream)(p);
    return cudaGetLastError();
}

}  // namespace dmk
