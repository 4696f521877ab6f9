// dmk_devsim_intrinsics.hpp -- TEST INFRASTRUCTURE.  The names darmok_b200/csrc/device_intrinsics.cuh gives the kernels,
// re-stated for a host compiler, plus the handful of CUDA built-ins the kernels use.  With -DDMK_DEVSIM the kernel sources
// of darmok_b200/csrc compile with g++ as they are and run on the CPU: one fiber per CUDA thread, one block at a time,
// blocking calls (__syncthreads, __syncwarp, warp collectives, mbarrier waits, __nanosleep) switch fibers
// (tests/devsim/devsim_runtime.cpp).  What this checks: index arithmetic (device buffers and dynamic shared memory are
// bounds-checked / ASan-visible heap blocks), hand-off protocols (a deadlock or a spin bound aborts with a message) and the
// host-side launch logic of api_*.cu -- not performance and not the memory model (fibers interleave only at blocking calls).
// Nothing under darmok_b200/ or include/ uses this; the product library is the nvcc build.
#pragma once

#include <cuda_runtime.h>   // the real header: vector types, dim3, error codes and the runtime's prototypes (no CUDA library is linked)

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <unordered_map>
#include <vector>
#include <functional>
#include <tuple>
#include <type_traits>
#include <utility>

#undef __shared__
#undef __launch_bounds__
#define __shared__ static        // blocks run one after the other: block-shared == static
#define __launch_bounds__(...)

namespace devsim {

extern uint3 g_thread_idx, g_block_idx;
extern dim3 g_block_dim, g_grid_dim;

[[noreturn]] void trap(const char* why);
void yield_blocked();                 // a wait that only another fiber's action can end
void yield_sleep();                   // a wait that time can end (__nanosleep in a bounded loop)
void note_progress();                 // some fiber changed state another may be waiting on
void sync_threads();
void sync_warp();
uint32_t warp_exchange(uint32_t mine, int src_lane);   // value of src_lane (own value when it is not a live lane of the warp)
uint32_t warp_ballot(bool pred);
unsigned char* dynamic_smem();        // this block's dynamic shared memory
size_t dynamic_smem_bytes();
void run_grid(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()>& thread_body);
uint64_t now_ns();
struct MBar;
void bulk_copy(void* dst, const void* src, uint32_t bytes, MBar* bar);   // lands at once, or some rounds later (DMK_DEVSIM_SEED)

template <class... P>
struct Launcher {
    void (*kernel)(P...);
    dim3 grid, block;
    size_t smem;
    template <class... A>
    void operator()(A&&... a) const {
        std::tuple<std::decay_t<P>...> args(std::forward<A>(a)...);
        auto k = kernel;
        run_grid(grid, block, smem, [&]() { std::apply(k, args); });
    }
};
template <class... P>
Launcher<P...> launcher(void (*kernel)(P...), dim3 grid, dim3 block, size_t smem) {
    return Launcher<P...>{kernel, grid, block, smem};
}

// shared-memory "addresses" as the kernels see them: byte offsets into the block's dynamic shared memory, biased so that 0
// is never valid
constexpr uint32_t kSmemBias = 1024;
inline uint32_t smem_offset(const void* p) {
    const unsigned char* base = dynamic_smem();
    const unsigned char* q = static_cast<const unsigned char*>(p);
    if (q < base || q >= base + dynamic_smem_bytes()) trap("shared-memory address outside the block's dynamic shared memory");
    return static_cast<uint32_t>(q - base) + kSmemBias;
}
inline unsigned char* smem_pointer(uint32_t addr, size_t bytes) {
    if (addr < kSmemBias || addr - kSmemBias + bytes > dynamic_smem_bytes()) trap("shared-memory access outside the block's dynamic shared memory");
    return dynamic_smem() + (addr - kSmemBias);
}

// mbarrier state in its 8 bytes of shared memory
struct MBar {
    uint32_t phase : 1;
    uint32_t expected : 15;
    uint32_t pending : 16;
    int32_t tx;
};
static_assert(sizeof(MBar) == 8, "mbarrier is a 64-bit object");
inline void mbar_check_complete(MBar* b) {
    if (b->pending == 0 && b->tx == 0) {
        b->phase ^= 1u;
        b->pending = b->expected;
        note_progress();
    }
}

}  // namespace devsim

#define threadIdx (::devsim::g_thread_idx)
#define blockIdx (::devsim::g_block_idx)
#define blockDim (::devsim::g_block_dim)
#define gridDim (::devsim::g_grid_dim)

#define DMK_LAUNCH(kernel, grid, block, smem_bytes, stream) ::devsim::launcher(kernel, (grid), (block), (smem_bytes))
#define DMK_LAUNCH_CLUSTER_OF_ONE(kernel, grid, block, smem_bytes, stream) ::devsim::launcher(kernel, (grid), (block), (smem_bytes))
#define DMK_DYNAMIC_SMEM(name, alignment) unsigned char* const name = ::devsim::dynamic_smem()

// cuda_runtime.h only has the kernel-pointer overload of this one under nvcc
template <class... P>
inline cudaError_t cudaFuncSetAttribute(void (*kernel)(P...), cudaFuncAttribute attr, int value) {
    return ::cudaFuncSetAttribute(reinterpret_cast<const void*>(kernel), attr, value);
}

// ---- CUDA built-ins used by the kernels ------------------------------------------------------------------------------
using std::max;
using std::min;
template <class T>
inline T __ldg(const T* p) { return *p; }
inline float __frcp_rn(float x) { return 1.0f / x; }            // built with -ffp-contract=off: correctly rounded, like the intrinsics
inline float __fsqrt_rn(float x) { return std::sqrt(x); }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline int __float_as_int(float f) { int u; std::memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
inline float __int_as_float(int u) { float f; std::memcpy(&f, &u, 4); return f; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline void __syncthreads() { ::devsim::sync_threads(); }
inline void __syncwarp(unsigned = 0xffffffffu) { ::devsim::sync_warp(); }
inline void __threadfence() {}
inline void __threadfence_system() {}
inline void __nanosleep(unsigned) { ::devsim::yield_sleep(); }
[[noreturn]] inline void __trap() { ::devsim::trap("__trap() reached in kernel code"); }
inline unsigned __ballot_sync(unsigned, int pred) { return ::devsim::warp_ballot(pred != 0); }
template <class T>
inline T __shfl_sync_impl(T v, int src) {
    static_assert(sizeof(T) == 4, "32-bit shuffles only");
    uint32_t bits;
    std::memcpy(&bits, &v, 4);
    bits = ::devsim::warp_exchange(bits, src);
    std::memcpy(&v, &bits, 4);
    return v;
}
template <class T>
inline T __shfl_up_sync(unsigned, T v, unsigned delta) {
    const int lane = static_cast<int>(threadIdx.x & 31u);
    return __shfl_sync_impl(v, lane >= static_cast<int>(delta) ? lane - static_cast<int>(delta) : lane);
}
template <class T>
inline T __shfl_down_sync(unsigned, T v, unsigned delta) {
    const int lane = static_cast<int>(threadIdx.x & 31u);
    return __shfl_sync_impl(v, lane + static_cast<int>(delta) < 32 ? lane + static_cast<int>(delta) : lane);
}
template <class T>
inline T __shfl_xor_sync(unsigned, T v, int lane_mask) {
    const int lane = static_cast<int>(threadIdx.x & 31u);
    return __shfl_sync_impl(v, lane ^ lane_mask);
}
template <class T, class U>
inline T atomicAdd(T* p, U v) { const T o = *p; *p = static_cast<T>(o + static_cast<T>(v)); ::devsim::note_progress(); return o; }
template <class T, class U>
inline T atomicExch(T* p, U v) { const T o = *p; *p = static_cast<T>(v); ::devsim::note_progress(); return o; }
template <class T, class U, class V>
inline T atomicCAS(T* p, U expected, V desired) {
    const T o = *p;
    if (o == static_cast<T>(expected)) *p = static_cast<T>(desired);
    ::devsim::note_progress();
    return o;
}

// ---- the names of device_intrinsics.cuh ------------------------------------------------------------------------------
namespace dmk {
namespace {

inline uint32_t smem_u32(const void* p) { return ::devsim::smem_offset(p); }

inline void mbar_init(uint64_t* bar, int count) {
    auto* b = reinterpret_cast<::devsim::MBar*>(bar);
    b->phase = 0;
    b->expected = static_cast<uint32_t>(count);
    b->pending = static_cast<uint32_t>(count);
    b->tx = 0;
}
inline void mbar_init_fence() {}
inline void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {   // arrive.expect_tx: one arrival + bytes to come
    auto* b = reinterpret_cast<::devsim::MBar*>(bar);
    if (b->pending == 0) ::devsim::trap("mbarrier: more arrivals than the barrier was initialised for");
    b->tx += static_cast<int32_t>(bytes);
    b->pending -= 1;
    ::devsim::mbar_check_complete(b);
}
inline void mbar_arrive(uint64_t* bar) {
    auto* b = reinterpret_cast<::devsim::MBar*>(bar);
    if (b->pending == 0) ::devsim::trap("mbarrier: more arrivals than the barrier was initialised for");
    b->pending -= 1;
    ::devsim::mbar_check_complete(b);
}
inline void mbar_wait(uint64_t* bar, uint32_t parity) {   // waits for the completion of the phase with this parity
    auto* b = reinterpret_cast<::devsim::MBar*>(bar);
    uint32_t spins = 0;
    while (b->phase == (parity & 1u)) {
        if (++spins > (1u << 22)) ::devsim::trap("mbarrier wait: spin bound reached");
        ::devsim::yield_blocked();
    }
}
inline void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    if (bytes % 16 || reinterpret_cast<uintptr_t>(src_gmem) % 16 || reinterpret_cast<uintptr_t>(dst_smem) % 16)
        ::devsim::trap("cp.async.bulk: size and both addresses must be multiples of 16");
    ::devsim::bulk_copy(::devsim::smem_pointer(::devsim::smem_offset(dst_smem), bytes), src_gmem, bytes, reinterpret_cast<::devsim::MBar*>(bar));
}

inline void bulk_s2s(void* dst_smem, const void* src_smem, uint32_t bytes, uint64_t* bar) {
    if (bytes % 16 || reinterpret_cast<uintptr_t>(src_smem) % 16 || reinterpret_cast<uintptr_t>(dst_smem) % 16)
        ::devsim::trap("cp.async.bulk (shared -> shared): size and both addresses must be multiples of 16");
    ::devsim::bulk_copy(::devsim::smem_pointer(::devsim::smem_offset(dst_smem), bytes), ::devsim::smem_pointer(::devsim::smem_offset(src_smem), bytes),
                        bytes, reinterpret_cast<::devsim::MBar*>(bar));
}

// bulk copies shared -> global, per issuing thread: an open group, committed groups, and the copy itself performed LAZILY -- when a
// wait says its source has been read -- so that a buffer overwritten before the matching wait shows as a wrong result
struct PendingStore { void* dst; uint32_t src; uint32_t bytes; };
struct StoreQueue { std::vector<PendingStore> open; std::deque<std::vector<PendingStore>> committed; };
inline std::unordered_map<unsigned, StoreQueue>& store_queues() { static std::unordered_map<unsigned, StoreQueue> q; return q; }
inline void bulk_s2g(void* dst_gmem, uint32_t src_smem_addr, uint32_t bytes) {
    if (bytes % 16 || src_smem_addr % 16 || reinterpret_cast<uintptr_t>(dst_gmem) % 16)
        ::devsim::trap("cp.async.bulk (shared -> global): size and both addresses must be multiples of 16");
    ::devsim::smem_pointer(src_smem_addr, bytes);   // bounds
    store_queues()[threadIdx.x].open.push_back({dst_gmem, src_smem_addr, bytes});
}
inline void bulk_commit() {
    StoreQueue& q = store_queues()[threadIdx.x];
    q.committed.push_back(std::move(q.open));
    q.open.clear();
}
inline void bulk_drain(size_t keep) {
    StoreQueue& q = store_queues()[threadIdx.x];
    while (q.committed.size() > keep) {
        for (const PendingStore& c : q.committed.front()) std::memcpy(c.dst, ::devsim::smem_pointer(c.src, c.bytes), c.bytes);
        q.committed.pop_front();
        ::devsim::note_progress();
    }
}
template <int N>
inline void bulk_wait_read() { bulk_drain(N); }
inline void bulk_wait_all() { bulk_drain(0); }
inline void fence_proxy_async_smem() {}
inline void sts32(uint32_t addr, float v) {
    if (addr % 4) ::devsim::trap("st.shared.f32: address not 4-byte aligned");
    std::memcpy(::devsim::smem_pointer(addr, 4), &v, 4);
}

inline void check_aligned(const void* p, size_t a, const char* what) {
    if (reinterpret_cast<uintptr_t>(p) % a) ::devsim::trap(what);
}
inline void st_v8(float* p, const float* v) {
    check_aligned(p, 32, "st.global.v8.f32: address not 32-byte aligned");
    std::memcpy(p, v, 32);
}
inline void st_v8(float* p, float a, float b, float c, float d, float e, float f, float g, float h) {
    const float v[8] = {a, b, c, d, e, f, g, h};
    st_v8(p, v);
}
inline void st_v8_if(float* p, const float* v, uint32_t take) {
    if (take) st_v8(p, v);
}
inline void ld_v8(const float* p, float* v) {
    check_aligned(p, 32, "ld.global.v8.f32: address not 32-byte aligned");
    std::memcpy(v, p, 32);
}
inline void st_v4(float* p, float a, float b, float c, float d) {
    check_aligned(p, 16, "st.global.v4.f32: address not 16-byte aligned");
    const float v[4] = {a, b, c, d};
    std::memcpy(p, v, 16);
}
inline float4 lds128(uint32_t addr) {
    if (addr % 16) ::devsim::trap("ld.shared.v4.f32: address not 16-byte aligned");
    float4 v;
    std::memcpy(&v, ::devsim::smem_pointer(addr, 16), 16);
    return v;
}
inline float4 lds128_if(uint32_t addr, uint32_t take) { return take ? lds128(addr) : make_float4(0.f, 0.f, 0.f, 0.f); }
inline void lds128_keep(float4& v, uint32_t addr, uint32_t take) {
    if (take) v = lds128(addr);
}
inline void warp_copy_palette(const float4* src, float4* dst, int palette_vec4, int lane) {
    for (int k = lane; k < palette_vec4; k += 32) dst[k] = src[k];
}
inline void warp_copy_palette34_as_mat4(const float* src34, float4* dst, int palette_vec4, int lane) {
    for (int k = lane; k < palette_vec4; k += 32) {
        const float* s = src34 + (size_t)(k >> 2) * 16 + (k & 3);
        dst[k] = make_float4(s[0], s[4], s[8], (k & 3) == 3 ? 1.f : 0.f);
    }
}
inline float4 lds128_or(uint32_t addr, uint32_t take, const float4& fallback) { return take ? lds128(addr) : fallback; }

inline uint32_t ld_acquire_sys(const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); }
inline void st_release_sys(uint32_t* p, uint32_t v) {
    *reinterpret_cast<volatile uint32_t*>(p) = v;
    ::devsim::note_progress();
}
inline uint64_t global_timer_ns() { return ::devsim::now_ns(); }

}  // namespace
}  // namespace dmk
