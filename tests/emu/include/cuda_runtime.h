// TEST INFRASTRUCTURE - never shipped, never loaded by the blobs_b200 package.
//
// A stand-in for <cuda_runtime.h> that lets blobs_b200/csrc/{world,capi}.cu and kernels.cuh be compiled by g++ as plain
// C++ (-DBLOBS_EMU) into tests/emu/libblobs_b200_emu.so. The container this project is developed in has no GPU; this build
// runs the very same kernel SOURCE on the CPU, one cooperative fiber per CUDA thread, so that the parity tests can check the
// kernels' LOGIC (indexing, ordering, warp collectives, shared-memory protocols, the host orchestration) against the oracle
// before any GPU time is spent. It says nothing about performance and is not a fallback: the product loader
// (blobs_b200/_lib.py) only ever opens libblobs_b200.so and fails without it.
//
// Model: kernels of a launch run CTA after CTA on the calling OS thread; inside a CTA every thread is a ucontext fiber,
// scheduled round-robin and switched only inside the collectives below (__syncthreads, __syncwarp, *_sync). `__shared__`
// becomes `static` (one CTA at a time), atomics are plain read-modify-writes, device memory is host memory (malloc, filled
// with 0xCD so that reads of never-written device memory show up).
// f32 arithmetic: __fadd_rn & co are the C++ operators (the TU is built with -ffp-contract=off -frounding-math, no fast-math),
// directed roundings go through fesetround. sincosf is glibc's (CUDA's differs by ulps: only rotating bodies see it).
#pragma once
#include <cfenv>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>

// ---------------------------------------------------------------------------------------------- qualifiers
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __restrict__ __restrict
#define __shared__ static
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

// ---------------------------------------------------------------------------------------------- vector types
struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct uint2 { unsigned int x, y; };
struct alignas(16) uint4 { unsigned int x, y, z, w; };
struct int2 { int x, y; };
struct uint3 { unsigned int x, y, z; };
struct dim3 {
    unsigned int x, y, z;
    dim3(unsigned int x_ = 1, unsigned int y_ = 1, unsigned int z_ = 1) : x(x_), y(y_), z(z_) {}
};
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline int2 make_int2(int x, int y) { return int2{x, y}; }

// ---------------------------------------------------------------------------------------------- execution engine (emu_runtime.cpp)
namespace emu {
struct ThreadCtx;
extern uint3 g_threadIdx, g_blockIdx;
extern dim3 g_blockDim, g_gridDim;
void run_grid(unsigned grid, unsigned block, size_t dyn_smem, const std::function<void()>& body);
void* dynamic_smem();
void cta_barrier();
void warp_barrier(unsigned mask);
unsigned lane_id();
unsigned long long* warp_slots();   // 32 x u64 exchange slots of the current warp
unsigned long long launches();

template <class... A>
struct Launch {
    unsigned grid, block;
    size_t smem;
    void (*fn)(A...);
    void operator()(A... args) const {
        run_grid(grid, block, smem, [&] { fn(args...); });
    }
};
template <class... A>
Launch<A...> make_launch(dim3 g, dim3 b, size_t smem, void (*fn)(A...)) {
    return Launch<A...>{g.x * g.y * g.z, b.x * b.y * b.z, smem, fn};
}
}  // namespace emu

#define threadIdx (::emu::g_threadIdx)
#define blockIdx (::emu::g_blockIdx)
#define blockDim (::emu::g_blockDim)
#define gridDim (::emu::g_gridDim)

// ---------------------------------------------------------------------------------------------- scalar intrinsics
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float emu_round_op(int mode, float a, float b, bool sub) {
    const int old = fegetround();
    fesetround(mode);
    volatile float va = a, vb = b;
    volatile float r = sub ? va - vb : va + vb;
    fesetround(old);
    return r;
}
static inline float __fadd_ru(float a, float b) { return emu_round_op(FE_UPWARD, a, b, false); }
static inline float __fadd_rd(float a, float b) { return emu_round_op(FE_DOWNWARD, a, b, false); }
static inline float __fsub_ru(float a, float b) { return emu_round_op(FE_UPWARD, a, b, true); }
static inline float __fsub_rd(float a, float b) { return emu_round_op(FE_DOWNWARD, a, b, true); }
static inline int __float2int_rd(float v) {   // floor, saturating, NaN -> 0 (PTX cvt.rmi.s32.f32)
    if (v != v) return 0;
    const float f = floorf(v);
    if (f >= 2147483648.0f) return INT32_MAX;
    if (f < -2147483648.0f) return INT32_MIN;
    return (int)f;
}
static inline unsigned int __float_as_uint(float f) { unsigned int u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned int u) { float f; memcpy(&f, &u, 4); return f; }
static inline int __popc(unsigned int x) { return __builtin_popcount(x); }
static inline int __ffs(unsigned int x) { return __builtin_ffs((int)x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) {
    return (unsigned long long)(((unsigned __int128)a * (unsigned __int128)b) >> 64);
}
template <class T> static inline T __ldg(const T* p) { return *p; }

static inline unsigned int min(unsigned int a, unsigned int b) { return a < b ? a : b; }
static inline unsigned int max(unsigned int a, unsigned int b) { return a > b ? a : b; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned int min(unsigned int a, int b) { return a < (unsigned int)b ? a : (unsigned int)b; }
static inline unsigned int min(int a, unsigned int b) { return (unsigned int)a < b ? (unsigned int)a : b; }
static inline unsigned int max(unsigned int a, int b) { return a > (unsigned int)b ? a : (unsigned int)b; }
static inline unsigned int max(int a, unsigned int b) { return (unsigned int)a > b ? (unsigned int)a : b; }
static inline unsigned long long min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
static inline unsigned long long max(unsigned long long a, unsigned long long b) { return a > b ? a : b; }
static inline float min(float a, float b) { return fminf(a, b); }
static inline float max(float a, float b) { return fmaxf(a, b); }

// ---------------------------------------------------------------------------------------------- atomics (one OS thread)
template <class T> static inline T emu_atomic_add(T* p, T v) { T o = *p; *p = (T)(o + v); return o; }
static inline unsigned int atomicAdd(unsigned int* p, unsigned int v) { return emu_atomic_add(p, v); }
static inline int atomicAdd(int* p, int v) { return emu_atomic_add(p, v); }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return emu_atomic_add(p, v); }
static inline float atomicAdd(float* p, float v) { return emu_atomic_add(p, v); }
static inline unsigned int atomicOr(unsigned int* p, unsigned int v) { unsigned int o = *p; *p = o | v; return o; }
static inline int atomicMax(int* p, int v) { int o = *p; if (v > o) *p = v; return o; }
static inline int atomicMin(int* p, int v) { int o = *p; if (v < o) *p = v; return o; }
static inline unsigned int atomicMax(unsigned int* p, unsigned int v) { unsigned int o = *p; if (v > o) *p = v; return o; }
static inline unsigned int atomicMin(unsigned int* p, unsigned int v) { unsigned int o = *p; if (v < o) *p = v; return o; }

// ---------------------------------------------------------------------------------------------- barriers and warp collectives
static inline void __syncthreads() { ::emu::cta_barrier(); }
static inline void __syncwarp(unsigned mask = 0xffffffffu) { ::emu::warp_barrier(mask); }
// Divergence is not modelled: a lane only knows about itself. Callers use this for warp-aggregated atomics, which degrade
// to one atomic per lane (same result).
static inline unsigned __activemask() { return 1u << ::emu::lane_id(); }

template <class T> static inline unsigned long long emu_to_u64(T v) { unsigned long long u = 0; static_assert(sizeof(T) <= 8, "shfl payload"); memcpy(&u, &v, sizeof(T)); return u; }
template <class T> static inline T emu_from_u64(unsigned long long u) { T v; memcpy(&v, &u, sizeof(T)); return v; }

// every lane publishes a value, then `f(slots, lane)` computes this lane's result from all published values
template <class T, class F> static inline auto emu_collective(unsigned mask, T v, F&& f) {
    unsigned long long* s = ::emu::warp_slots();
    const unsigned lane = ::emu::lane_id();
    s[lane] = emu_to_u64(v);
    ::emu::warp_barrier(mask);
    auto r = f(s, lane);
    ::emu::warp_barrier(mask);   // nobody overwrites a slot before everyone has read
    return r;
}
template <class T> static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    (void)width;
    return emu_collective(mask, v, [&](unsigned long long* s, unsigned) { return emu_from_u64<T>(s[(unsigned)src & 31u]); });
}
template <class T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    (void)width;
    return emu_collective(mask, v, [&](unsigned long long* s, unsigned lane) { return lane >= delta ? emu_from_u64<T>(s[lane - delta]) : v; });
}
template <class T> static inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    (void)width;
    return emu_collective(mask, v, [&](unsigned long long* s, unsigned lane) { return lane + delta < 32u ? emu_from_u64<T>(s[lane + delta]) : v; });
}
template <class T> static inline T __shfl_xor_sync(unsigned mask, T v, int lanemask, int width = 32) {
    (void)width;
    return emu_collective(mask, v, [&](unsigned long long* s, unsigned lane) { return emu_from_u64<T>(s[(lane ^ (unsigned)lanemask) & 31u]); });
}
// lanes named in `mask` that have exited do not publish: the engine zeroes the slots of exited lanes, and a zero is the
// neutral value for ballot / any / add / max-of-unsigned; min and match take the participation mask into account
namespace emu { unsigned participants(unsigned mask); }
static inline unsigned __ballot_sync(unsigned mask, int pred) {
    return emu_collective(mask, (unsigned)(pred ? 1u : 0u), [&](unsigned long long* s, unsigned) {
        unsigned r = 0;
        const unsigned part = ::emu::participants(mask);
        for (unsigned i = 0; i < 32u; ++i) if (((part >> i) & 1u) && s[i]) r |= 1u << i;
        return r;
    });
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0u; }
static inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == ::emu::participants(mask); }
template <class T, class OP> static inline T emu_reduce(unsigned mask, T v, OP op) {
    return emu_collective(mask, v, [&](unsigned long long* s, unsigned) {
        const unsigned part = ::emu::participants(mask);
        bool first = true;
        T acc = v;
        for (unsigned i = 0; i < 32u; ++i) {
            if (!((part >> i) & 1u)) continue;
            const T x = emu_from_u64<T>(s[i]);
            acc = first ? x : op(acc, x);
            first = false;
        }
        return acc;
    });
}
static inline unsigned __reduce_add_sync(unsigned mask, unsigned v) { return emu_reduce(mask, v, [](unsigned a, unsigned b) { return a + b; }); }
static inline int __reduce_add_sync(unsigned mask, int v) { return emu_reduce(mask, v, [](int a, int b) { return a + b; }); }
static inline unsigned __reduce_or_sync(unsigned mask, unsigned v) { return emu_reduce(mask, v, [](unsigned a, unsigned b) { return a | b; }); }
static inline unsigned __reduce_max_sync(unsigned mask, unsigned v) { return emu_reduce(mask, v, [](unsigned a, unsigned b) { return a > b ? a : b; }); }
static inline int __reduce_max_sync(unsigned mask, int v) { return emu_reduce(mask, v, [](int a, int b) { return a > b ? a : b; }); }
static inline unsigned __reduce_min_sync(unsigned mask, unsigned v) { return emu_reduce(mask, v, [](unsigned a, unsigned b) { return a < b ? a : b; }); }
static inline int __reduce_min_sync(unsigned mask, int v) { return emu_reduce(mask, v, [](int a, int b) { return a < b ? a : b; }); }
template <class T> static inline unsigned __match_any_sync(unsigned mask, T v) {
    return emu_collective(mask, v, [&](unsigned long long* s, unsigned) {
        const unsigned part = ::emu::participants(mask);
        const unsigned long long mine = emu_to_u64(v);
        unsigned r = 0;
        for (unsigned i = 0; i < 32u; ++i) if (((part >> i) & 1u) && s[i] == mine) r |= 1u << i;
        return r;
    });
}

// ---------------------------------------------------------------------------------------------- runtime API
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorNotSupported = 801 };
typedef struct emuStream* cudaStream_t;
typedef struct emuEvent { std::chrono::steady_clock::time_point t; }* cudaEvent_t;
typedef struct emuGraph* cudaGraph_t;
typedef struct emuGraphExec* cudaGraphExec_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaStreamNonBlocking = 1, cudaStreamCaptureModeThreadLocal = 1, cudaEventRecordExternal = 1, cudaEventDisableTiming = 2 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9 };
enum { cudaSharedmemCarveoutMaxShared = 100 };

static inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : (e == cudaErrorNotSupported ? "not supported by the CPU test build" : "error"); }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
template <class T> static inline cudaError_t cudaMalloc(T** p, size_t n) {
    *p = static_cast<T*>(malloc(n ? n : 1));
    if (!*p) return cudaErrorMemoryAllocation;
    memset(*p, 0xCD, n);
    return cudaSuccess;
}
// ---- CUDA IPC stand-in: "device memory" another rank PROCESS can map = a POSIX shared-memory segment (emu_runtime.cpp)
typedef struct cudaIpcMemHandle_st { char reserved[64]; } cudaIpcMemHandle_t;
enum { cudaIpcMemLazyEnablePeerAccess = 1 };
namespace emu {
cudaError_t ipc_alloc(void** p, size_t bytes);
void ipc_free(void* p);
cudaError_t ipc_get_handle(cudaIpcMemHandle_t* h, void* p);
cudaError_t ipc_open(void** p, const cudaIpcMemHandle_t& h);
cudaError_t ipc_close(void* p);
}  // namespace emu
static inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) { return emu::ipc_get_handle(h, p); }
static inline cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned) { return emu::ipc_open(p, h); }
static inline cudaError_t cudaIpcCloseMemHandle(void* p) { return emu::ipc_close(p); }
static inline long long clock64() { return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static inline void __threadfence_system() { __sync_synchronize(); }
static inline void __threadfence() { __sync_synchronize(); }
template <class T> static inline cudaError_t cudaMallocHost(T** p, size_t n) { *p = static_cast<T*>(calloc(1, n ? n : 1)); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { if (n) memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { if (n) memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { if (n) memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new emuEvent(); return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = new emuEvent(); return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }   // every stream runs synchronously here
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
static inline cudaError_t cudaEventRecordWithFlags(cudaEvent_t e, cudaStream_t, unsigned) { return cudaEventRecord(e); }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count(); return cudaSuccess; }
// CUDA graphs do not exist here: the BLOBS_EMU build of World never captures (graphs_on = false)
static inline cudaError_t cudaStreamBeginCapture(cudaStream_t, int) { return cudaErrorNotSupported; }
static inline cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t* g) { *g = nullptr; return cudaErrorNotSupported; }
static inline cudaError_t cudaGraphInstantiate(cudaGraphExec_t* e, cudaGraph_t, unsigned long long) { *e = nullptr; return cudaErrorNotSupported; }
static inline cudaError_t cudaGraphLaunch(cudaGraphExec_t, cudaStream_t) { return cudaErrorNotSupported; }
static inline cudaError_t cudaGraphDestroy(cudaGraph_t) { return cudaSuccess; }
static inline cudaError_t cudaGraphExecDestroy(cudaGraphExec_t) { return cudaSuccess; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
