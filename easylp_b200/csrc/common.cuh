// common.cuh — error handling, device buffers, launch accounting, small device helpers.
// Part of libeasylp_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <cstring>
#include <string>
#include <stdexcept>
#include <atomic>
#include <chrono>
#include <vector>

namespace elp {

constexpr int kNumSMs = 148;   // B200: 2 dies x 74 SMs; grids are sized in multiples of this

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

inline std::string format(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    return std::string(buf);
}

#define ELP_CUDA(call)                                                                          \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess)                                                                 \
            throw elp::Error(elp::format("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,  \
                                         cudaGetErrorString(e__)));                             \
    } while (0)

#define ELP_REQUIRE(cond, ...)                                    \
    do {                                                          \
        if (!(cond)) throw elp::Error(elp::format(__VA_ARGS__)); \
    } while (0)

extern std::atomic<int64_t> g_launches;        // every kernel launch of the library is counted here
extern thread_local std::string g_last_error;

// Launch wrapper: counts, launches, checks the launch error (not the execution).
#define ELP_LAUNCH(kernel, grid, block, smem, stream, ...)                         \
    do {                                                                           \
        elp::g_launches.fetch_add(1, std::memory_order_relaxed);                   \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                \
        ELP_CUDA(cudaGetLastError());                                              \
    } while (0)

// One device allocation handed out in pieces.  cudaMalloc / cudaFree cost 2-13 ms PER CALL on a B200 box whatever the
// size (scripts/malloc_probe.cu: 1.2 GB in one piece 2 + 6 ms), and a PDLP handle used to make ~60 of them.  While an
// ArenaScope is active on the calling thread, DevBuf::alloc takes its memory from the arena (256-byte aligned, never
// returned piecewise; the arena is freed as a whole by its owner) and falls back to cudaMalloc when the arena is full.
// Released arenas are parked in a small per-thread cache instead of going back to the driver: cudaFree of a ~1 GB block
// was measured at 0.8 / 83 / 661 ms on three consecutive solves of config 5 (scripts/gpu_close_probe.py), more than the
// solve itself.  elp_release_workspace() (and the end of the thread) empties the cache.
struct ArenaCache {
    struct Block { char* p; size_t cap; int dev; };
    std::vector<Block> blocks;
    ~ArenaCache() { clear(); }
    void clear() {
        for (Block& b : blocks) cudaFree(b.p);
        blocks.clear();
        cudaGetLastError();
    }
    char* take(size_t bytes, size_t* cap_out) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
        int best = -1;
        for (int i = 0; i < (int)blocks.size(); ++i)
            if (blocks[i].dev == dev && blocks[i].cap >= bytes && blocks[i].cap <= 4 * bytes + (64u << 20) &&
                (best < 0 || blocks[i].cap < blocks[best].cap)) best = i;
        if (best < 0) return nullptr;
        char* p = blocks[best].p;
        *cap_out = blocks[best].cap;
        blocks.erase(blocks.begin() + best);
        return p;
    }
    void park(char* p, size_t cap) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || blocks.size() >= 4) { cudaFree(p); return; }
        blocks.push_back(Block{p, cap, dev});
    }
};
inline thread_local ArenaCache g_arena_cache;

struct Arena {
    char* base = nullptr;
    size_t cap = 0, used = 0;
    Arena() = default;
    Arena(const Arena&) = delete;
    Arena& operator=(const Arena&) = delete;
    ~Arena() { release(); }
    void reserve(size_t bytes) {
        release();
        if (!bytes) return;
        if (char* p = g_arena_cache.take(bytes, &cap)) { base = p; return; }
        if (cudaMalloc(&base, bytes) == cudaSuccess) cap = bytes;
        else { base = nullptr; cap = 0; cudaGetLastError(); }      // no arena: every buffer gets its own allocation
    }
    void* take(size_t bytes) {
        const size_t at = (used + 255) & ~(size_t)255;
        if (!base || at + bytes > cap) return nullptr;
        used = at + bytes;
        return base + at;
    }
    void release() {
        if (base) g_arena_cache.park(base, cap);
        base = nullptr; cap = used = 0;
    }
};
inline thread_local Arena* g_arena = nullptr;
struct ArenaScope {
    Arena* prev;
    explicit ArenaScope(Arena* a) : prev(g_arena) { g_arena = a; }
    ~ArenaScope() { g_arena = prev; }
    ArenaScope(const ArenaScope&) = delete;
    ArenaScope& operator=(const ArenaScope&) = delete;
};

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    bool owned = true;            // false: a piece of an Arena (freed with it)
    DevBuf() = default;
    explicit DevBuf(size_t count) { alloc(count); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n), owned(o.owned) { o.p = nullptr; o.n = 0; o.owned = true; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) { release(); p = o.p; n = o.n; owned = o.owned; o.p = nullptr; o.n = 0; o.owned = true; }
        return *this;
    }
    ~DevBuf() { release(); }
    void alloc(size_t count) {
        release();
        n = count;
        if (!count) return;
        if (g_arena) {
            p = static_cast<T*>(g_arena->take(count * sizeof(T)));
            if (p) { owned = false; return; }
        }
        ELP_CUDA(cudaMalloc(&p, count * sizeof(T)));
        owned = true;
    }
    void release() {
        if (p && owned) cudaFree(p);
        p = nullptr; n = 0; owned = true;
    }
    void upload(const T* host, size_t count, cudaStream_t s = 0) {
        if (count) ELP_CUDA(cudaMemcpyAsync(p, host, count * sizeof(T), cudaMemcpyHostToDevice, s));
    }
    void download(T* host, size_t count, cudaStream_t s = 0) const {
        if (count) ELP_CUDA(cudaMemcpyAsync(host, p, count * sizeof(T), cudaMemcpyDeviceToHost, s));
    }
    void zero(cudaStream_t s = 0) {
        if (n) ELP_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s));
    }
};

struct WallTimer {
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    double ms() const {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
};

inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---- device helpers ------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <int L>
__device__ __forceinline__ double group_sum(double v) {   // sum over aligned groups of L lanes
#pragma unroll
    for (int o = L / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
template <int L>
__device__ __forceinline__ double group_max(double v) {
#pragma unroll
    for (int o = L / 2; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// streaming (read-once) loads of the matrix arrays: read-only path, do not allocate in L1
__device__ __forceinline__ double ld_stream(const double* p) {
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ int ld_stream(const int* p) {
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

}  // namespace elp
