// Sx: maximum upwind slope along the rays of an azimuth sector (Winstral), reference topo.py:775-858
// (host geometry) and 928-953 (_sx_rolling, numba).  The host passes the de-duplicated ray samples of
// each azimuth as (dy, dx, 1/d); the device scans them for every interior pixel:
//     sx = rad2deg(atan(max_k ((z[p + o_k] - z[p]) - h) / d_k))
// atan is monotonic, so the max is taken over the tangents (float32 FMNMX, which skips NaN samples
// exactly like np.nanmax) and a single float64 atan is evaluated per pixel.
#include <math.h>

#include "common.cuh"

namespace topo {

struct SxParams {
    const float* dem;
    float* out;
    int64_t ld_in, ld_out, az_stride;
    int nx, gny, in_gy0, in_rows, out_gy0, out_rows;
    const int* offsets;     // (dy, dx) pairs
    const float* inv_dist;  // 1 / distance
    const int* az_begin;    // CSR over azimuths
    int window;
    float height;
};

constexpr int kSxChunk = 1024;  // samples staged in shared memory per round
constexpr int kSxRows = 4;      // rows per CTA (block = 64 x 4)

struct SxSample {
    int off;    // dy * ld + dx  (element offset from the centre pixel)
    float inv;  // 1 / distance
};

__global__ void __launch_bounds__(256) sx_kernel(const SxParams p) {
    __shared__ SxSample smp[kSxChunk];
    const int az = blockIdx.z;
    const int begin = p.az_begin[az], end = p.az_begin[az + 1];
    const int x = blockIdx.x * 64 + (threadIdx.x & 63);
    const int gy = p.out_gy0 + blockIdx.y * kSxRows + (threadIdx.x >> 6);
    const bool valid = x < p.nx && gy < p.out_gy0 + p.out_rows;
    const bool interior = valid && gy >= p.window && gy < p.gny - p.window && x >= p.window && x < p.nx - p.window;

    const float* c = p.dem + (int64_t)((interior ? gy : p.in_gy0) - p.in_gy0) * p.ld_in + (interior ? x : 0);
    const float z0 = interior ? __ldg(c) : 0.f;
    float best = __int_as_float(0x7fc00000);  // NaN: fmaxf(NaN, t) = t

    for (int s0 = begin; s0 < end; s0 += kSxChunk) {
        const int n = min(kSxChunk, end - s0);
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += 256) {
            const int dy = p.offsets[2 * (s0 + i)], dx = p.offsets[2 * (s0 + i) + 1];
            smp[i].off = dy * (int)p.ld_in + dx;
            smp[i].inv = p.inv_dist[s0 + i];
        }
        __syncthreads();
        if (interior) {
#pragma unroll 4
            for (int i = 0; i < n; ++i) {
                const SxSample sm = smp[i];
                const float zk = __ldg(c + sm.off);
                const float t = __fmul_rn(__fsub_rn(__fsub_rn(zk, z0), p.height), sm.inv);
                best = fmaxf(best, t);
            }
        }
    }
    if (valid) {
        float r = 0.f;  // the frame of `window` pixels stays 0 (np.zeros_like, topo.py:939-941)
        if (interior) r = (float)(atan((double)best) * 57.29577951308232);
        p.out[(int64_t)az * p.az_stride + (int64_t)(gy - p.out_gy0) * p.ld_out + x] = r;
    }
}

}  // namespace topo

using namespace topo;

extern "C" {

int topo_sx_f32(const float* dem, int64_t ld_in, float* out, int64_t ld_out, int64_t az_stride, const topo_view* v,
                const int* offsets, const float* inv_dist, const int* az_begin, int n_az, int window, float height,
                int dy_min, int dy_max, void* stream) {
    TOPO_CHECK(dem && out && az_begin, "null pointer");
    if (validate_view(v)) return -1;
    TOPO_CHECK(n_az >= 1 && n_az <= 65535, "n_az outside [1, 65535]");
    TOPO_CHECK(window >= 0, "negative window");
    TOPO_CHECK(ld_in < (1ll << 30), "row pitch too large for 32-bit sample offsets");
    TOPO_CHECK(dy_min >= -window && dy_max <= window, "sample offsets exceed the window");
    if (v->out_rows == 0) return 0;
    {
        // interior output rows and the rows their samples touch must be inside the band
        int lo = v->out_gy0 > window ? v->out_gy0 : window;
        int hi = v->out_gy0 + v->out_rows < v->gny - window ? v->out_gy0 + v->out_rows : v->gny - window;
        if (lo < hi) {
            TOPO_CHECK(v->in_gy0 <= lo + (dy_min < 0 ? dy_min : 0) && v->in_gy0 + v->in_rows >= hi + (dy_max > 0 ? dy_max : 0),
                       "sx: input band [%d,%d) does not cover rows [%d,%d)", v->in_gy0, v->in_gy0 + v->in_rows,
                       lo + dy_min, hi + dy_max);
        }
    }
    SxParams p{dem, out, ld_in, ld_out, az_stride, v->nx, v->gny, v->in_gy0, v->in_rows, v->out_gy0, v->out_rows,
               offsets, inv_dist, az_begin, window, height};
    dim3 grid(ceil_div(v->nx, 64), ceil_div(v->out_rows, kSxRows), n_az);
    TOPO_CHECK(grid.y <= 65535, "too many rows for one launch");
    cudaStream_t s = (cudaStream_t)stream;
    TOPO_LAUNCH("sx", s, sx_kernel<<<grid, 256, 0, s>>>(p));
    return 0;
}

}  // extern "C"
