// Sx: maximum upwind slope along the rays of an azimuth sector (Winstral), reference topo.py:775-858
// (host geometry) and 928-953 (_sx_rolling, numba).  The host passes the de-duplicated ray samples of
// each azimuth as (dy, dx, 1/d); the device scans them for every interior pixel:
//     sx = rad2deg(atan(max_k ((z[p + o_k] - z[p]) - h) / d_k))
// atan is monotonic, so the max is taken over the tangents (float32 FMNMX, which skips NaN samples
// exactly like np.nanmax) and a single atanf is evaluated per pixel and azimuth.  The tangent is
// fma(z[p + o_k] - z[p], 1/d_k, -(h/d_k)): the difference of neighbouring float32 heights is (nearly) exact and
// the fma rounds once, so the tangent carries ~1e-7 relative error (6e-6 degrees).
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "tma.cuh"

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

struct __align__(16) SxSample {
    int off;    // dy * ld + dx  (element offset from the centre pixel)
    float inv;  // 1 / distance
    float nh;   // -(height / distance)
    int pad;
};

// slope angle of the largest tangent, in degrees.  float32 atanf (<= 2 ulp) + one rounding: <= 3e-5 degrees
// from the float64 result, well inside the 1e-3 degree tolerance -- and 1/10 of the cost of a float64 atan,
// which at 72 azimuths x 16.7 Mpx would otherwise take as long as the scan itself.
__device__ __forceinline__ float sx_degrees(float tangent) { return atanf(tangent) * 57.295779513082321f; }

__device__ __forceinline__ float sx_tangent(float zk, float z0, const SxSample& sm) {
    return __fmaf_rn(__fsub_rn(zk, z0), sm.inv, sm.nh);
}

// 8-byte samples for the gather kernel: a 16-byte broadcast LDS per sample measured 25% slower there
struct SxSample8 {
    int off;
    float inv;
};

__global__ void __launch_bounds__(256) sx_kernel(const SxParams p) {
    __shared__ SxSample8 smp[kSxChunk];
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
                const SxSample8 s8 = smp[i];
                const SxSample sm{s8.off, s8.inv, -__fmul_rn(p.height, s8.inv), 0};
                const float zk = __ldg(c + sm.off);
                const float t = sx_tangent(zk, z0, sm);
                best = fmaxf(best, t);
            }
        }
    }
    if (valid) {
        float r = 0.f;  // the frame of `window` pixels stays 0 (np.zeros_like, topo.py:939-941)
        if (interior) r = sx_degrees(best);
        p.out[(int64_t)az * p.az_stride + (int64_t)(gy - p.out_gy0) * p.ld_out + x] = r;
    }
}

// ---- TMA-staged variant -------------------------------------------------------------------------------
// The DEM tile plus the bounding box of the azimuth's ray samples is brought into shared memory by ONE
// cp.async.bulk.tensor (TMA) box per CTA (out-of-image elements arrive as zeros; interior pixels never read
// them), the samples become shared-memory offsets, and the scan runs on conflict-free LDS (lanes = columns)
// instead of L1-cached gathers.  Used when the box fits (<= 256 x 256 elements and the shared-memory budget).
constexpr int kSxTW = 128;  // output tile: 128 columns x 16 rows, 256 threads, 8 rows per thread
constexpr int kSxTH = 16;

struct SxTmaParams {
    float* out;
    int64_t ld_out, az_stride;
    int nx, gny, in_gy0, out_gy0, out_rows;
    const int* offsets;
    const float* inv_dist;
    const int* az_begin;
    int n_az, az_per_cta;
    int dy_org, dx_org;  // tile(0, 0) = pixel (tile row 0 + dy_org, tile column 0 + dx_org); dx_org % 4 == 0
    int window, box_h;
    float height;
};

// PITCH = box width = shared-memory row pitch, a compile-time constant so that the 8 rows a thread owns are
// reached with immediate LDS offsets: per sample 1 broadcast LDS.64 + 1 IADD + 8 x (LDS + 2 FADD + FMUL + FMNMX).
// One CTA scans az_per_cta consecutive azimuths on the same staged tile (the box spans the samples of all of them).
template <int PITCH>
__global__ void __launch_bounds__(256) sx_tma_kernel(const __grid_constant__ CUtensorMap tmap, const SxTmaParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* tile = reinterpret_cast<float*>(smem_raw);  // [box_h][PITCH]
    __shared__ SxSample smp[kSxChunk];
    __shared__ __align__(8) uint64_t bar;

    const int x0 = blockIdx.x * kSxTW;
    const int y0 = p.out_gy0 + blockIdx.y * kSxTH;  // global row of the tile's first output row

    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, (uint32_t)(PITCH * p.box_h * sizeof(float)));
        // the box origin's x must be a multiple of 4 elements (16 bytes): measured on sm_100a, an unaligned x
        // coordinate raises "illegal instruction" with INTERLEAVE_NONE / SWIZZLE_NONE (profiles/micro/tma_probe.cu);
        // the tensor map's row 0 is global row in_gy0
        tma_load_2d(tile, &tmap, x0 + p.dx_org, y0 + p.dy_org - p.in_gy0, &bar);
    }

    const int tx = threadIdx.x & (kSxTW - 1);
    const int tyg = threadIdx.x >> 7;  // 0..1: rows tyg*8 .. tyg*8+7
    const int x = x0 + tx;
    const int c0 = -p.dy_org * PITCH - p.dx_org;              // tile offset of a pixel relative to its (ty, tx) base
    const float* b = tile + tyg * (kSxTH / 2) * PITCH + tx;  // first of this thread's 8 rows
    float z0[kSxTH / 2];
    bool landed = false;

    const int az_end = min(p.n_az, (int)(blockIdx.z + 1) * p.az_per_cta);
    for (int az = blockIdx.z * p.az_per_cta; az < az_end; ++az) {
        const int begin = p.az_begin[az], end = p.az_begin[az + 1];
        float best[kSxTH / 2];
#pragma unroll
        for (int r = 0; r < kSxTH / 2; ++r) best[r] = __int_as_float(0x7fc00000);  // NaN: fmaxf(NaN, t) = t
        for (int s0 = begin; s0 < end; s0 += kSxChunk) {
            const int n = min(kSxChunk, end - s0);
            __syncthreads();
            for (int i = threadIdx.x; i < n; i += 256) {
                const int dy = p.offsets[2 * (s0 + i)], dx = p.offsets[2 * (s0 + i) + 1];
                smp[i].off = (dy - p.dy_org) * PITCH + (dx - p.dx_org);
                smp[i].inv = p.inv_dist[s0 + i];
                smp[i].nh = -__fmul_rn(p.height, smp[i].inv);
            }
            __syncthreads();
            if (!landed) {
                mbar_wait(&bar, 0);
                landed = true;
#pragma unroll
                for (int r = 0; r < kSxTH / 2; ++r) z0[r] = b[c0 + r * PITCH];
            }
#pragma unroll 2
            for (int i = 0; i < n; ++i) {
                const SxSample sm = smp[i];
                const float* q = b + sm.off;
#pragma unroll
                for (int r = 0; r < kSxTH / 2; ++r) best[r] = fmaxf(best[r], sx_tangent(q[r * PITCH], z0[r], sm));
            }
        }
        if (x < p.nx) {
#pragma unroll
            for (int r = 0; r < kSxTH / 2; ++r) {
                const int gy = y0 + tyg * (kSxTH / 2) + r;
                if (gy >= p.out_gy0 + p.out_rows) break;
                const bool interior = gy >= p.window && gy < p.gny - p.window && x >= p.window && x < p.nx - p.window;
                // the frame of `window` pixels stays 0 (np.zeros_like, topo.py:939-941)
                __stcs(&p.out[(int64_t)az * p.az_stride + (int64_t)(gy - p.out_gy0) * p.ld_out + x],
                       interior ? sx_degrees(best[r]) : 0.f);
            }
        }
    }
    if (!landed) mbar_wait(&bar, 0);  // no samples at all: still drain the copy before the CTA exits
}

}  // namespace topo

using namespace topo;

extern "C" {

int topo_sx_f32(const float* dem, int64_t ld_in, float* out, int64_t ld_out, int64_t az_stride, const topo_view* v,
                const int* offsets, const float* inv_dist, const int* az_begin, int n_az, int window,
                float height, int dy_min, int dy_max, int dx_min, int dx_max, void* stream) {
    TOPO_CHECK(dem && out && az_begin, "null pointer");
    if (validate_view(v)) return -1;
    TOPO_CHECK(n_az >= 1 && n_az <= 65535, "n_az outside [1, 65535]");
    TOPO_CHECK(window >= 0, "negative window");
    TOPO_CHECK(ld_in < (1ll << 30), "row pitch too large for 32-bit sample offsets");
    TOPO_CHECK(dy_min >= -window && dy_max <= window && dx_min >= -window && dx_max <= window, "sample offsets exceed the window");
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
    cudaStream_t s = (cudaStream_t)stream;
    // ---- TMA-staged path: the per-azimuth sample bounding box (<= the global extents given here) must fit a box
    if (option_enabled(kOptSxTma)) {
        const int dyl = dy_min < 0 ? dy_min : 0, dyh = dy_max > 0 ? dy_max : 0;
        const int dxl = (dx_min < 0 ? dx_min : 0) & ~3, dxh = dx_max > 0 ? dx_max : 0;  // origin x aligned down to 4
        // conservative box: every azimuth's own bounding box (which includes the centre) is within the global span
        const int box_h = kSxTH + (dyh - dyl);
        int box_w = ((kSxTW + (dxh - dxl)) + 31) & ~31;  // 160 / 192 / 224 / 256: one instantiation each
        if (box_w < 160) box_w = 160;
        const size_t smem = (size_t)box_w * box_h * sizeof(float);
        CUtensorMap tmap;
        if (box_w <= 256 && box_h <= 256 && smem <= 150 * 1024 &&
            make_tmap_2d_f32(&tmap, dem, (uint64_t)v->nx, (uint64_t)v->in_rows, (uint64_t)ld_in, (uint32_t)box_w, (uint32_t)box_h)) {
            auto kern = box_w <= 160 ? sx_tma_kernel<160> : box_w <= 192 ? sx_tma_kernel<192>
                        : box_w <= 224 ? sx_tma_kernel<224> : sx_tma_kernel<256>;
            static bool attr_set[64][4] = {{false}};
            int dev = 0;
            TOPO_CUDA(cudaGetDevice(&dev));
            const int kidx = (box_w - 129) / 32;
            if (dev < 64 && !attr_set[dev][kidx]) {
                TOPO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
                attr_set[dev][kidx] = true;
            }
            // azimuths per CTA: as many as still leave >= 16 CTAs per SM in the grid
            const int64_t tiles = (int64_t)ceil_div(v->nx, kSxTW) * ceil_div(v->out_rows, kSxTH);
            int groups = (int)((148 * 16 + tiles - 1) / tiles);
            if (groups > n_az) groups = n_az;
            if (groups < 1) groups = 1;
            const int az_per_cta = ceil_div(n_az, groups);
            SxTmaParams q{out, ld_out, az_stride, v->nx, v->gny, v->in_gy0, v->out_gy0, v->out_rows, offsets, inv_dist,
                          az_begin, n_az, az_per_cta, dyl, dxl, window, box_h, height};
            dim3 grid(ceil_div(v->nx, kSxTW), ceil_div(v->out_rows, kSxTH), ceil_div(n_az, az_per_cta));
            TOPO_CHECK(grid.y <= 65535, "too many rows for one launch");
            TOPO_LAUNCH("sx_tma", s, kern<<<grid, 256, smem, s>>>(tmap, q));
            return 0;
        }
    }
    SxParams p{dem, out, ld_in, ld_out, az_stride, v->nx, v->gny, v->in_gy0, v->in_rows, v->out_gy0, v->out_rows,
               offsets, inv_dist, az_begin, window, height};
    dim3 grid(ceil_div(v->nx, 64), ceil_div(v->out_rows, kSxRows), n_az);
    TOPO_CHECK(grid.y <= 65535, "too many rows for one launch");
    TOPO_LAUNCH("sx", s, sx_kernel<<<grid, 256, 0, s>>>(p));
    return 0;
}

}  // extern "C"
