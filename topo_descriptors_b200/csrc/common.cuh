// Shared device/host utilities for libtopo_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/topo_b200.h"

namespace topo {

// ---- error plumbing (never throw across the C ABI) ---------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

// Execution-path switches (topo_set_option): every setting computes the same results through a different kernel
// shape; the tests flip them to cross-check the shapes bit for bit.  Never read from the environment.
enum Option { kOptOctagon = 0, kOptTiny = 1, kOptSxTma = 2, kOptGaussFft = 3, kOptGradFused = 4, kOptDiscFft = 5, kOptCount = 6 };
bool option_enabled(int opt);

// Optional per-kernel timing (topo_profile_enable / topo_profile_dump): CUDA events recorded on the
// launching stream around each kernel, aggregated by kernel name.  Off by default (zero overhead).
struct ProfScope {
    int slot;
    cudaStream_t stream;
    ProfScope(const char* name, cudaStream_t s);
    ~ProfScope();
};

#define TOPO_CHECK(cond, ...)            \
    do {                                 \
        if (!(cond)) {                   \
            topo::set_error(__VA_ARGS__); \
            return -1;                   \
        }                                \
    } while (0)

#define TOPO_CUDA(call)                                                                   \
    do {                                                                                  \
        cudaError_t e__ = (call);                                                         \
        if (e__ != cudaSuccess) {                                                         \
            topo::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, \
                            __LINE__);                                                    \
            return -2;                                                                    \
        }                                                                                 \
    } while (0)

// TOPO_LAUNCH("kernel_name", stream, kernel<<<grid, block, smem, stream>>>(args));
#define TOPO_LAUNCH(name, stream, ...)           \
    do {                                         \
        {                                        \
            topo::ProfScope prof__(name, stream); \
            __VA_ARGS__;                         \
        }                                        \
        TOPO_LAUNCH_CHECK();                     \
    } while (0)

#define TOPO_LAUNCH_CHECK()                                                               \
    do {                                                                                  \
        cudaError_t e__ = cudaGetLastError();                                             \
        if (e__ != cudaSuccess) {                                                         \
            topo::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__),  \
                            __FILE__, __LINE__);                                          \
            return -3;                                                                    \
        }                                                                                 \
        topo::count_launch();                                                             \
    } while (0)

inline int validate_view(const topo_view* v) {
    TOPO_CHECK(v != nullptr, "null view");
    TOPO_CHECK(v->nx > 0 && v->gny > 0, "empty image (%d x %d)", v->gny, v->nx);
    TOPO_CHECK(v->in_rows > 0 && v->out_rows >= 0, "bad band rows");
    TOPO_CHECK(v->in_gy0 >= 0 && v->in_gy0 + v->in_rows <= v->gny, "input band outside image");
    TOPO_CHECK(v->out_gy0 >= 0 && v->out_gy0 + v->out_rows <= v->gny, "output band outside image");
    return 0;
}

constexpr int kNumSMs = 148;  // B200

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- device helpers ----------------------------------------------------------------------------
// ndimage 'reflect' (d c b a | a b c d | d c b a), any distance outside [0, n).
__host__ __device__ __forceinline__ int reflect_index(int i, int n) {
    if (i >= 0 && i < n) return i;
    int p = 2 * n;
    int r = i % p;
    if (r < 0) r += p;
    return r < n ? r : p - 1 - r;
}

__device__ __forceinline__ float4 ldg4(const float* p) {
    return __ldg(reinterpret_cast<const float4*>(p));
}

// streaming 128-bit store (outputs are written once and not re-read by the same kernel)
__device__ __forceinline__ void st4_streaming(float* p, float4 v) {
    __stcs(reinterpret_cast<float4*>(p), v);
}

__device__ __forceinline__ uint32_t warp_inclusive_scan_u32(uint32_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

}  // namespace topo
