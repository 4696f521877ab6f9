// devsim_runtime.cpp -- TEST INFRASTRUCTURE (see dmk_devsim_intrinsics.hpp): the lockstep executor behind DMK_LAUNCH and a
// stand-in for the few CUDA runtime calls darmok_b200/csrc/api_*.cu make, with "device" memory on the host heap (so that
// AddressSanitizer sees every kernel access).  Kernels run synchronously inside the launch call, blocks one after the
// other, the threads of a block as fibers switched at blocking calls only.
//
// The library built from this refuses to create a context unless DMK_DEVSIM=1 is set in the environment: it can never be
// mistaken for, or silently stand in for, the CUDA build.
#include "dmk_devsim_intrinsics.hpp"

#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>
#if !defined(__x86_64__)
#include <ucontext.h>
#endif

#include <chrono>
#include <map>
#include <vector>

#if defined(__SANITIZE_ADDRESS__)
#include <sanitizer/common_interface_defs.h>
#define DEVSIM_ASAN 1
#else
#define DEVSIM_ASAN 0
#endif

namespace devsim {

uint3 g_thread_idx, g_block_idx;
dim3 g_block_dim, g_grid_dim;

namespace {

constexpr size_t kStackBytes = 256 * 1024;
constexpr int kMaxThreads = 1024;

// Fiber switch.  x86-64: six callee-saved registers + the floating-point control words and the stack pointer, no system call
// (swapcontext saves the signal mask with one per switch, which dominated the run time); elsewhere: ucontext.
#if defined(__x86_64__)
struct Context {
    void* sp = nullptr;
};
extern "C" void devsim_switch_context(void** save_sp, void* load_sp);
asm(R"(
    .text
    .globl devsim_switch_context
    .hidden devsim_switch_context
    .type devsim_switch_context,@function
devsim_switch_context:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    subq $8, %rsp
    stmxcsr (%rsp)
    fnstcw 4(%rsp)
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    ldmxcsr (%rsp)
    fldcw 4(%rsp)
    addq $8, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
    .size devsim_switch_context,.-devsim_switch_context
)");
void prepare_context(Context& c, unsigned char* stack, size_t bytes, void (*entry)()) {
    // stack as devsim_switch_context leaves it: {mxcsr, x87 cw}, r15, r14, r13, r12, rbx, rbp, return address (= entry), and
    // above it the slot a caller's return address would occupy: entry starts with rsp % 16 == 8 as the ABI has it after a call
    uintptr_t top = (reinterpret_cast<uintptr_t>(stack) + bytes) & ~uintptr_t(15);
    uint64_t* sp = reinterpret_cast<uint64_t*>(top);
    *--sp = 0;                                        // return address of entry: never used, entry does not return
    *--sp = reinterpret_cast<uint64_t>(entry);        // popped by ret
    for (int k = 0; k < 6; ++k) *--sp = 0;            // rbp, rbx, r12..r15
    *--sp = 0x0000037f00001f80ull;                    // mxcsr = 0x1f80 (low word), x87 control word = 0x037f
    c.sp = sp;
}
inline void switch_context(Context& from, Context& to) { devsim_switch_context(&from.sp, to.sp); }
#else
struct Context {
    ucontext_t uc;
};
void prepare_context(Context& c, unsigned char* stack, size_t bytes, void (*entry)()) {
    getcontext(&c.uc);
    c.uc.uc_stack.ss_sp = stack;
    c.uc.uc_stack.ss_size = bytes;
    c.uc.uc_link = nullptr;
    makecontext(&c.uc, entry, 0);
}
inline void switch_context(Context& from, Context& to) { swapcontext(&from.uc, &to.uc); }
#endif

struct Fiber {
    Context ctx;
    unsigned char* stack = nullptr;
    uint3 tid;
    int linear = 0;
    bool done = false;
    bool blocked = false;   // last yield was a wait only another fiber can end
};

struct Warp {
    int live = 0, arrived = 0;
    unsigned generation = 0;
    uint32_t slot[32];
    bool alive[32];
};

struct PendingCopy {   // a bulk copy in flight (DMK_DEVSIM_SEED: it lands some scheduler rounds after its issue)
    void* dst;
    const void* src;
    uint32_t bytes;
    MBar* bar;
    int rounds_left;
};

struct Block {
    std::vector<PendingCopy> copies;
    std::vector<Fiber> fibers;
    std::vector<Warp> warps;
    int live = 0, arrived = 0;
    unsigned generation = 0;
    unsigned char* smem = nullptr;
    size_t smem_bytes = 0;
    uint64_t progress = 0;
};

Context g_scheduler;
#if DEVSIM_ASAN
const void* g_main_stack_bottom = nullptr;
size_t g_main_stack_size = 0;
#endif
Fiber* g_cur = nullptr;
Block* g_block = nullptr;
const std::function<void()>* g_body = nullptr;
unsigned char* g_stacks = nullptr;

// DMK_DEVSIM_SEED=<n>, n != 0: instead of the fixed round-robin, every scheduler round runs the warps in a random order and leaves
// each one out with probability 1/2 (warps drift apart by many rounds), runs the lanes of a warp in a permuted order (lane ^ r), and
// lets bulk copies land 0..3 rounds after their issue
uint64_t g_rng = 0;
bool g_random_schedule = false, g_skip_warp = false;
unsigned g_lane_xor = 0;
uint32_t next_random() {   // xorshift64*
    g_rng ^= g_rng >> 12;
    g_rng ^= g_rng << 25;
    g_rng ^= g_rng >> 27;
    return static_cast<uint32_t>((g_rng * 0x2545F4914F6CDD1Dull) >> 32);
}
void read_schedule_options() {
    static bool done = false;
    if (done) return;
    done = true;
    const char* e = std::getenv("DMK_DEVSIM_SEED");
    const uint64_t seed = e ? std::strtoull(e, nullptr, 10) : 0;
    g_random_schedule = seed != 0;
    g_rng = seed * 0x9E3779B97F4A7C15ull + 1;
}
void land(const PendingCopy& c) {
    std::memcpy(c.dst, c.src, c.bytes);
    c.bar->tx -= static_cast<int32_t>(c.bytes);
    mbar_check_complete(c.bar);
}

void switch_to_scheduler() {
#if DEVSIM_ASAN
    void* fake = nullptr;
    __sanitizer_start_switch_fiber(&fake, g_main_stack_bottom, g_main_stack_size);   // the scheduler runs on the thread's own stack
    switch_context(g_cur->ctx, g_scheduler);
    __sanitizer_finish_switch_fiber(fake, nullptr, nullptr);
#else
    switch_context(g_cur->ctx, g_scheduler);
#endif
}

void release_block_barrier_if_complete(Block& b) {
    if (b.live > 0 && b.arrived == b.live) {
        b.arrived = 0;
        ++b.generation;
        ++b.progress;
    }
}
void release_warp_barrier_if_complete(Block& b, Warp& w) {
    if (w.live > 0 && w.arrived == w.live) {
        w.arrived = 0;
        ++w.generation;
        ++b.progress;
    }
}

void fiber_entry() {
#if DEVSIM_ASAN
    __sanitizer_finish_switch_fiber(nullptr, &g_main_stack_bottom, &g_main_stack_size);
#endif
    (*g_body)();
    Fiber* f = g_cur;
    Block& b = *g_block;
    f->done = true;
    Warp& w = b.warps[static_cast<size_t>(f->linear >> 5)];
    w.alive[f->linear & 31] = false;
    --w.live;
    --b.live;
    ++b.progress;
    // a thread that has exited no longer takes part in barriers
    release_warp_barrier_if_complete(b, w);
    release_block_barrier_if_complete(b);
#if DEVSIM_ASAN
    __sanitizer_start_switch_fiber(nullptr, g_main_stack_bottom, g_main_stack_size);   // null save slot: this fiber never resumes
#endif
    switch_context(f->ctx, g_scheduler);
    std::abort();
}

}  // namespace

[[noreturn]] void trap(const char* why) {
    std::fprintf(stderr, "devsim: %s (block %u,%u thread %u of %u)\n", why, g_block_idx.x, g_block_idx.y, g_thread_idx.x, g_block_dim.x);
    std::fflush(stderr);
    std::abort();
}

uint64_t now_ns() {
    return static_cast<uint64_t>(std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count());
}

void note_progress() {
    if (g_block) ++g_block->progress;
}
void yield_blocked() {
    if (!g_cur) trap("blocking call outside a kernel");
    g_cur->blocked = true;
    switch_to_scheduler();
}
void yield_sleep() {
    if (!g_cur) trap("blocking call outside a kernel");
    g_cur->blocked = false;
    ++g_block->progress;
    switch_to_scheduler();
}

void sync_threads() {
    Block& b = *g_block;
    const unsigned gen = b.generation;
    ++b.arrived;
    release_block_barrier_if_complete(b);
    while (b.generation == gen) yield_blocked();
}
void sync_warp() {
    Block& b = *g_block;
    Warp& w = b.warps[static_cast<size_t>(g_cur->linear >> 5)];
    const unsigned gen = w.generation;
    ++w.arrived;
    release_warp_barrier_if_complete(b, w);
    while (w.generation == gen) yield_blocked();
}
uint32_t warp_exchange(uint32_t mine, int src_lane) {
    Warp& w = g_block->warps[static_cast<size_t>(g_cur->linear >> 5)];
    const int lane = g_cur->linear & 31;
    w.slot[lane] = mine;
    sync_warp();
    const uint32_t got = (src_lane >= 0 && src_lane < 32 && w.alive[src_lane]) ? w.slot[src_lane] : mine;
    sync_warp();   // nobody overwrites a slot before everybody has read
    return got;
}
uint32_t warp_ballot(bool pred) {
    Warp& w = g_block->warps[static_cast<size_t>(g_cur->linear >> 5)];
    const int lane = g_cur->linear & 31;
    w.slot[lane] = pred ? 1u : 0u;
    sync_warp();
    uint32_t bits = 0;
    for (int l = 0; l < 32; ++l)
        if (w.alive[l] && w.slot[l]) bits |= 1u << l;
    sync_warp();
    return bits;
}

void bulk_copy(void* dst, const void* src, uint32_t bytes, MBar* bar) {
    PendingCopy c{dst, src, bytes, bar, g_random_schedule ? static_cast<int>(next_random() % 4) : 0};
    if (c.rounds_left == 0) land(c);   // at once: the earliest possible overwrite of the destination
    else g_block->copies.push_back(c);
}

unsigned char* dynamic_smem() { return g_block ? g_block->smem : nullptr; }
size_t dynamic_smem_bytes() { return g_block ? g_block->smem_bytes : 0; }

void run_grid(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()>& thread_body) {
    if (g_block) trap("nested kernel launch");
    read_schedule_options();
    const size_t nthreads = static_cast<size_t>(block.x) * block.y * block.z;
    if (nthreads == 0 || nthreads > kMaxThreads) trap("launch: block size outside [1, 1024]");
    if (grid.x == 0 || grid.y == 0 || grid.z == 0) trap("launch: empty grid");
    if (smem_bytes > 227 * 1024) trap("launch: more than 227 KB of dynamic shared memory");
    if (!g_stacks) {
        g_stacks = static_cast<unsigned char*>(mmap(nullptr, kStackBytes * kMaxThreads, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0));
        if (g_stacks == MAP_FAILED) trap("cannot map fiber stacks");
    }
    g_body = &thread_body;
    g_grid_dim = grid;
    g_block_dim = block;
    Block b;
    g_block = &b;
    b.fibers.resize(nthreads);
    b.warps.resize((nthreads + 31) / 32);
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                g_block_idx = uint3{bx, by, bz};
                // a fresh heap block per CTA: AddressSanitizer sees accesses past the dynamic shared memory the launch asked for
                b.smem_bytes = smem_bytes;
                b.smem = nullptr;
                if (smem_bytes) {
                    void* m = nullptr;
                    if (posix_memalign(&m, 1024, smem_bytes) != 0) trap("out of memory for a block's shared memory");
                    b.smem = static_cast<unsigned char*>(m);
                    std::memset(b.smem, 0xcd, smem_bytes);
                }
                b.live = static_cast<int>(nthreads);
                b.arrived = 0;
                for (auto& w : b.warps) {
                    w.live = 0;
                    w.arrived = 0;
                    for (int l = 0; l < 32; ++l) { w.alive[l] = false; w.slot[l] = 0; }
                }
                for (size_t t = 0; t < nthreads; ++t) {
                    Fiber& f = b.fibers[t];
                    f.linear = static_cast<int>(t);
                    f.tid = uint3{static_cast<unsigned>(t % block.x), static_cast<unsigned>((t / block.x) % block.y),
                                  static_cast<unsigned>(t / (static_cast<size_t>(block.x) * block.y))};
                    f.done = false;
                    f.blocked = false;
                    f.stack = g_stacks + t * kStackBytes;
                    prepare_context(f.ctx, f.stack, kStackBytes, fiber_entry);
                    Warp& w = b.warps[t >> 5];
                    w.alive[t & 31] = true;
                    ++w.live;
                }
                int idle_rounds = 0;
                b.copies.clear();
                std::vector<int> warp_order(b.warps.size());
                for (size_t w = 0; w < warp_order.size(); ++w) warp_order[w] = static_cast<int>(w);
                while (b.live > 0) {
                    const uint64_t before = b.progress;
                    bool all_blocked = true;
                    for (size_t k = 0; k < b.copies.size();) {   // bulk copies in flight
                        if (--b.copies[k].rounds_left <= 0) {
                            land(b.copies[k]);
                            b.copies[k] = b.copies.back();
                            b.copies.pop_back();
                        } else {
                            ++k;
                        }
                    }
                    if (!b.copies.empty()) all_blocked = false;
                    if (g_random_schedule)
                        for (size_t w = warp_order.size(); w > 1; --w) std::swap(warp_order[w - 1], warp_order[next_random() % w]);
                    for (size_t slot = 0; slot < b.warps.size() * 32; ++slot) {   // whole warps: the lane permutation maps into them
                        // random mode also runs the lanes of a warp in a rotated, possibly reversed order: code that leans on lanes
                        // running in step without a __syncwarp() sees its neighbours' data too early or too late
                        if ((slot & 31) == 0) {
                            g_skip_warp = g_random_schedule && (next_random() & 1u) != 0;
                            g_lane_xor = g_random_schedule ? next_random() & 31u : 0u;
                        }
                        const size_t t = static_cast<size_t>(warp_order[slot >> 5]) * 32 + ((slot & 31) ^ g_lane_xor);
                        if (t >= nthreads) continue;
                        Fiber& f = b.fibers[t];
                        if (f.done) continue;
                        if (g_skip_warp) continue;   // (its fibers keep the blocked / not blocked state of their last run)
                        g_cur = &f;
                        g_thread_idx = f.tid;
                        f.blocked = false;
#if DEVSIM_ASAN
                        void* fake = nullptr;
                        __sanitizer_start_switch_fiber(&fake, f.stack, kStackBytes);
                        switch_context(g_scheduler, f.ctx);
                        __sanitizer_finish_switch_fiber(fake, nullptr, nullptr);
#else
                        switch_context(g_scheduler, f.ctx);
#endif
                    }
                    g_cur = nullptr;
                    // deadlock: every live fiber sits in a wait that only another fiber can end, and round after round nothing changes
                    // (64 rounds: under the random schedule a fiber only runs in about every other one)
                    for (size_t t = 0; t < nthreads && all_blocked; ++t)
                        if (!b.fibers[t].done && !b.fibers[t].blocked) all_blocked = false;
                    if (b.live > 0 && all_blocked && b.progress == before) {
                        if (++idle_rounds > 64) trap("deadlock: every live thread of the block waits and nothing changes");
                    } else {
                        idle_rounds = 0;
                    }
                }
                std::free(b.smem);
                b.smem = nullptr;
            }
    g_block = nullptr;
    g_body = nullptr;
}

}  // namespace devsim

// ======================================================================================================================
// CUDA runtime stand-in: exactly the calls api_*.cu make.  One device ("B200", 148 SMs, sm_100), memory on the heap,
// streams and events are tokens (everything is synchronous).
// ======================================================================================================================
namespace {
bool devsim_enabled() {
    const char* e = std::getenv("DMK_DEVSIM");
    return e && e[0] == '1';
}
struct EventImpl { uint64_t t_ns = 0; };
int g_stream_token;
long g_live_allocations = 0, g_allocation_count = 0, g_fail_countdown = 0;
struct SharedBlock { int fd; size_t size; };
std::map<void*, SharedBlock> g_shared;   // memfd-backed "device" blocks of this process
std::map<void*, size_t> g_opened;        // blocks of other processes mapped here
struct IpcHandle { void* pointer; int64_t pid; int32_t fd; uint64_t size; };
size_t shared_min_bytes() {
    const char* e = std::getenv("DMK_DEVSIM_SHARED");
    return e && e[0] == '1' ? size_t(65536) : ~size_t(0);
}
bool allocation_fails() {
    ++g_allocation_count;
    return g_fail_countdown > 0 && --g_fail_countdown == 0;
}
}  // namespace

extern "C" {

static int device_count() {   // DMK_DEVSIM_DEVICES: how many (identical, independent) devices to announce; one process per device in the tests
    if (!devsim_enabled()) return 0;   // without the switch this library has no device: dmk_context_create fails
    const char* e = std::getenv("DMK_DEVSIM_DEVICES");
    const int n = e ? std::atoi(e) : 1;
    return n < 1 ? 1 : n;
}
cudaError_t cudaGetDeviceCount(int* count) {
    *count = device_count();
    return cudaSuccess;
}
cudaError_t cudaSetDevice(int device) { return device >= 0 && device < device_count() ? cudaSuccess : cudaErrorInvalidDevice; }
cudaError_t cudaGetDeviceProperties_v2(cudaDeviceProp* prop, int device) {
    if (device < 0 || device >= device_count()) return cudaErrorInvalidDevice;
    std::memset(prop, 0, sizeof *prop);
    std::snprintf(prop->name, sizeof prop->name, "devsim (CPU lockstep simulator, not a GPU)");
    prop->major = 10;
    prop->minor = 0;
    prop->multiProcessorCount = 148;
    prop->sharedMemPerBlockOptin = 227 * 1024;
    prop->warpSize = 32;
    return cudaSuccess;
}
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t e) {
    return e == cudaSuccess ? "no error" : e == cudaErrorMemoryAllocation ? "out of memory" : "devsim runtime error";
}
cudaError_t cudaDeviceSynchronize(void) { return cudaSuccess; }

// Allocation bookkeeping for the tests: live device / pinned blocks, and an injected failure -- devsim_fail_allocation_after(k)
// makes the k-th allocation from now (device or pinned, counted together) fail once with cudaErrorMemoryAllocation.
__attribute__((visibility("default"))) long devsim_live_allocations(void) { return g_live_allocations; }
__attribute__((visibility("default"))) long devsim_allocation_count(void) { return g_allocation_count; }
__attribute__((visibility("default"))) void devsim_fail_allocation_after(long k) { g_fail_countdown = k; }

cudaError_t cudaMalloc(void** p, size_t bytes) {
    *p = nullptr;
    if (allocation_fails()) return cudaErrorMemoryAllocation;
    if (bytes >= shared_min_bytes()) {
        // DMK_DEVSIM_SHARED=1: large blocks are memfd mappings, so that cudaIpcGetMemHandle / cudaIpcOpenMemHandle can map them into
        // another PROCESS (the two-process check of the peer-memory exchange); ASan does not see past their end
        const size_t size = (bytes + 4095) & ~size_t(4095);
        const int fd = memfd_create("devsim", 0);
        if (fd < 0 || ftruncate(fd, static_cast<off_t>(size)) != 0) return cudaErrorMemoryAllocation;
        void* m = mmap(nullptr, size, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        if (m == MAP_FAILED) { close(fd); return cudaErrorMemoryAllocation; }
        std::memset(m, 0xcd, bytes);
        g_shared[m] = SharedBlock{fd, size};
        *p = m;
        ++g_live_allocations;
        return cudaSuccess;
    }
    // cudaMalloc's 256-byte alignment, but the exact size: ASan flags any access past the end
    if (posix_memalign(p, 256, bytes ? bytes : 1) != 0) { *p = nullptr; return cudaErrorMemoryAllocation; }
    std::memset(*p, 0xcd, bytes);
    ++g_live_allocations;
    return cudaSuccess;
}
cudaError_t cudaFree(void* p) {
    if (p) --g_live_allocations;
    auto it = g_shared.find(p);
    if (it != g_shared.end()) {
        munmap(p, it->second.size);
        close(it->second.fd);
        g_shared.erase(it);
        return cudaSuccess;
    }
    std::free(p);
    return cudaSuccess;
}
cudaError_t cudaHostAlloc(void** p, size_t bytes, unsigned int) {
    *p = nullptr;
    if (allocation_fails()) return cudaErrorMemoryAllocation;
    *p = std::malloc(bytes ? bytes : 1);
    if (*p) ++g_live_allocations;
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
cudaError_t cudaFreeHost(void* p) {
    if (p) --g_live_allocations;
    std::free(p);
    return cudaSuccess;
}
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t bytes, cudaMemcpyKind, cudaStream_t) {
    if (bytes) std::memmove(dst, src, bytes);
    return cudaSuccess;
}
cudaError_t cudaMemcpy2DAsync(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height, cudaMemcpyKind, cudaStream_t) {
    for (size_t r = 0; r < height; ++r) std::memmove(static_cast<char*>(dst) + r * dpitch, static_cast<const char*>(src) + r * spitch, width);
    return cudaSuccess;
}
cudaError_t cudaMemsetAsync(void* p, int value, size_t bytes, cudaStream_t) {
    if (bytes) std::memset(p, value, bytes);
    return cudaSuccess;
}

cudaError_t cudaStreamCreate(cudaStream_t* s) {
    *s = reinterpret_cast<cudaStream_t>(&g_stream_token);
    return cudaSuccess;
}
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned int) { return cudaStreamCreate(s); }
cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned int) { return cudaSuccess; }   // (everything here completes at the call)

cudaError_t cudaEventCreate(cudaEvent_t* e) {
    *e = reinterpret_cast<cudaEvent_t>(new EventImpl());
    return cudaSuccess;
}
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned int) { return cudaEventCreate(e); }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) {
    reinterpret_cast<EventImpl*>(e)->t_ns = devsim::now_ns();
    return cudaSuccess;
}
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
    *ms = static_cast<float>(static_cast<double>(reinterpret_cast<EventImpl*>(b)->t_ns - reinterpret_cast<EventImpl*>(a)->t_ns) * 1e-6);
    return cudaSuccess;
}
cudaError_t cudaEventDestroy(cudaEvent_t e) {
    delete reinterpret_cast<EventImpl*>(e);
    return cudaSuccess;
}

cudaError_t cudaFuncSetAttribute(const void*, cudaFuncAttribute, int) { return cudaSuccess; }
cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* blocks, const void*, int, size_t) {
    *blocks = 1;
    return cudaSuccess;
}
cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessorWithFlags(int* blocks, const void*, int, size_t, unsigned int) {
    *blocks = 1;
    return cudaSuccess;
}

// IPC: the handle carries the pointer (same process) or {pid, fd, size} of a memfd block (DMK_DEVSIM_SHARED=1), which another
// process maps through /proc/<pid>/fd/<fd>
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* handle, void* p) {
    IpcHandle h{};
    h.pointer = p;
    h.pid = static_cast<int64_t>(getpid());
    auto it = g_shared.find(p);
    if (it != g_shared.end()) {
        h.fd = it->second.fd;
        h.size = static_cast<uint64_t>(it->second.size);
    } else {
        h.fd = -1;
    }
    static_assert(sizeof(IpcHandle) <= sizeof(cudaIpcMemHandle_t), "handle payload");
    std::memset(handle, 0, sizeof *handle);
    std::memcpy(handle, &h, sizeof h);
    return cudaSuccess;
}
cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t handle, unsigned int) {
    IpcHandle h{};
    std::memcpy(&h, &handle, sizeof h);
    *p = nullptr;
    if (h.pid == static_cast<int64_t>(getpid())) {
        *p = h.pointer;
        return cudaSuccess;
    }
    if (h.fd < 0) return cudaErrorInvalidValue;   // a heap block of another process: run both sides with DMK_DEVSIM_SHARED=1
    char path[64];
    std::snprintf(path, sizeof path, "/proc/%lld/fd/%d", static_cast<long long>(h.pid), h.fd);
    const int fd = open(path, O_RDWR);
    if (fd < 0) return cudaErrorInvalidValue;
    void* m = mmap(nullptr, h.size, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (m == MAP_FAILED) return cudaErrorMemoryAllocation;
    g_opened[m] = static_cast<size_t>(h.size);
    *p = m;
    return cudaSuccess;
}
cudaError_t cudaIpcCloseMemHandle(void* p) {
    auto it = g_opened.find(p);
    if (it != g_opened.end()) {
        munmap(p, it->second);
        g_opened.erase(it);
    }
    return cudaSuccess;
}
cudaError_t cudaDeviceCanAccessPeer(int* can, int, int) {
    *can = 1;
    return cudaSuccess;
}
cudaError_t cudaDeviceEnablePeerAccess(int, unsigned int) { return cudaSuccess; }

}  // extern "C"
