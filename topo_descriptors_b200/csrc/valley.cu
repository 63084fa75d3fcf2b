// Valley / ridge index, reference topo.py:389-453 (+ 466-531 for the kernel bank).
//
// One launch evaluates the whole rotated-kernel bank: for every angle the n_ch channel kernels
// (already channel-mixed and flipped by the host, see topo_b200.h) are correlated with the z-scored
// DEM, the maximum over channels is taken and folded into a running strict-'>' (max, argmax) that
// lives in registers for all 180 angles; norm is clipped at 0 on the way out.  Zero padding outside
// the global image, scipy 'same' centring through the per-angle anchors.
//
// Register blocking: a thread owns 4 vertically adjacent pixels in each of 4 columns 32 apart (lanes = columns,
// so the shared-memory DEM reads are conflict-free) and walks a kernel column top to bottom: each step loads one
// weight vector (warp-uniform 128-bit load served by L1) and one DEM sample per column and issues
// 4 x 4 x n_ch FMAs (48 for the usual 3 flats) with a rotating window of 4 weight vectors -- ~90% of the issue
// slots are FFMA (128 registers, 2 CTAs per SM; 2 columns per thread measured 13% slower on B200).
// The DEM tile + halo is staged once per CTA and reused by all angles.
#include <math.h>

#include "common.cuh"

#ifndef TOPO_VALLEY_VC
#define TOPO_VALLEY_VC 4
#endif

namespace topo {

constexpr int kVQ = 4;            // pixels per thread and column (vertical)
constexpr int kVC = TOPO_VALLEY_VC;  // columns per thread, 32 apart
constexpr int kVTileX = 32 * kVC; // block (32, 8): 128 columns x 8 thread rows x 4 pixels
constexpr int kVTileY = 8 * kVQ;  // 32 output rows per CTA

struct ValleyParams {
    const float* dem;  // z-scored DEM
    float* norm;
    float* dir;
    int64_t ld_in, ld_out;
    int nx, gny, in_gy0, in_rows, out_gy0, out_rows;
    const float* bank;        // per angle: [w][hp][4] floats (column-major kernel, 4 channel slots)
    const int* bank_hw;       // per angle: h, w, hp, first column's index in bank_cols
    const int2* bank_cols;    // per kernel column: (lo, n) row range to walk, multiples of 4
    const int64_t* bank_off;  // per angle: element offset into bank
    int n_angles;
    int HT, HB, HL, HR;  // halos: max anchor / max (h-1-anchor) over the angles
    int tile_pitch;
    int resume;  // flat lists longer than 4 run in channel groups: start from the raw (max, argmax) already in norm / dir
    int raw;     // ... and leave them raw (no clip) for the next group
};

template <int NCH, bool SMEM>
__global__ void __launch_bounds__(256) valley_kernel(const ValleyParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* tile = reinterpret_cast<float*>(smem_raw);
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int x0 = blockIdx.x * kVTileX;
    const int y0 = p.out_gy0 + blockIdx.y * kVTileY;  // global row of the tile's first output row
    const int in_end = p.in_gy0 + p.in_rows;

    if (SMEM) {
        const int tile_rows = kVTileY + p.HT + p.HB + 3;
        const int tile_cols = kVTileX + p.HL + p.HR;
        for (int r = ty; r < tile_rows; r += 8) {
            const int gy = y0 - p.HT + r;
            const bool row_ok = gy >= 0 && gy < p.gny && gy >= p.in_gy0 && gy < in_end;
            for (int c = tx; c < tile_cols; c += 32) {
                const int gx = x0 - p.HL + c;
                float v = 0.f;
                if (row_ok && gx >= 0 && gx < p.nx) v = __ldg(p.dem + (int64_t)(gy - p.in_gy0) * p.ld_in + gx);
                tile[r * p.tile_pitch + c] = v;
            }
        }
        __syncthreads();
    }

    float best[kVC][kVQ], bdir[kVC][kVQ];
#pragma unroll
    for (int c = 0; c < kVC; ++c)
#pragma unroll
        for (int q = 0; q < kVQ; ++q) best[c][q] = -INFINITY, bdir[c][q] = 0.f;
    if (p.resume) {
#pragma unroll
        for (int c = 0; c < kVC; ++c) {
            const int x = x0 + tx + 32 * c;
#pragma unroll
            for (int q = 0; q < kVQ; ++q) {
                const int gy = y0 + ty * kVQ + q;
                if (x < p.nx && gy < p.out_gy0 + p.out_rows) {
                    const int64_t o = (int64_t)(gy - p.out_gy0) * p.ld_out + x;
                    best[c][q] = p.norm[o], bdir[c][q] = p.dir[o];
                }
            }
        }
    }

    for (int a = 0; a < p.n_angles; ++a) {
        const int h = p.bank_hw[4 * a], w = p.bank_hw[4 * a + 1], hp = p.bank_hw[4 * a + 2];
        const int2* cols = p.bank_cols + p.bank_hw[4 * a + 3];
        const float4* wb = reinterpret_cast<const float4*>(p.bank + p.bank_off[a]);
        const int oy = p.HT - h / 2, ox = p.HL - w / 2;  // anchors: h/2, w/2 (flipped kernel)
        float acc[kVC][kVQ][NCH];
#pragma unroll
        for (int c = 0; c < kVC; ++c)
#pragma unroll
            for (int q = 0; q < kVQ; ++q)
#pragma unroll
                for (int ch = 0; ch < NCH; ++ch) acc[c][q][ch] = 0.f;

        for (int j = 0; j < w; ++j) {
            float4 wr[kVQ];
#pragma unroll
            for (int q = 0; q < kVQ; ++q) wr[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            const int2 cr = __ldg(cols + j);  // rows [cr.x, cr.x + cr.y) hold every non-zero weight of this column
            const float4* wcol = wb + (int64_t)j * hp;
            const int row0 = ty * kVQ + oy, col = tx + j + ox;  // tile coordinates of the sample under weight row 0
            const float* tp = tile + (row0 + cr.x) * p.tile_pitch + col;
            for (int s0 = cr.x; s0 < cr.x + cr.y; s0 += kVQ) {
#pragma unroll
                for (int s = 0; s < kVQ; ++s) {
#pragma unroll
                    for (int q = kVQ - 1; q > 0; --q) wr[q] = wr[q - 1];
                    wr[0] = __ldg(wcol + s0 + s);  // rows >= h are zero padding
                    float d[kVC];
                    if (SMEM) {
#pragma unroll
                        for (int c = 0; c < kVC; ++c) d[c] = tp[32 * c];  // sample row is independent of q (sliding window)
                        tp += p.tile_pitch;
                    } else {
                        const int gy = y0 - p.HT + row0 + s0 + s;
                        const bool row_ok = gy >= 0 && gy < p.gny && gy >= p.in_gy0 && gy < in_end;
#pragma unroll
                        for (int c = 0; c < kVC; ++c) {
                            const int gx = x0 - p.HL + col + 32 * c;
                            d[c] = 0.f;
                            if (row_ok && gx >= 0 && gx < p.nx) d[c] = __ldg(p.dem + (int64_t)(gy - p.in_gy0) * p.ld_in + gx);
                        }
                    }
#pragma unroll
                    for (int c = 0; c < kVC; ++c)
#pragma unroll
                        for (int q = 0; q < kVQ; ++q) {
                            acc[c][q][0] = fmaf(wr[q].x, d[c], acc[c][q][0]);
                            if constexpr (NCH > 1) acc[c][q][1] = fmaf(wr[q].y, d[c], acc[c][q][1]);
                            if constexpr (NCH > 2) acc[c][q][2] = fmaf(wr[q].z, d[c], acc[c][q][2]);
                            if constexpr (NCH > 3) acc[c][q][3] = fmaf(wr[q].w, d[c], acc[c][q][3]);
                        }
                }
            }
        }
#pragma unroll
        for (int c = 0; c < kVC; ++c)
#pragma unroll
            for (int q = 0; q < kVQ; ++q) {
                float m = acc[c][q][0];
#pragma unroll
                for (int ch = 1; ch < NCH; ++ch) m = fmaxf(m, acc[c][q][ch]);
                // strict '>' over the angles in order = the first angle that reaches the maximum; written so that it
                // also holds when a later channel group revisits earlier angles
                if (m > best[c][q] || (m == best[c][q] && (float)a < bdir[c][q])) best[c][q] = m, bdir[c][q] = (float)a;
            }
    }

#pragma unroll
    for (int c = 0; c < kVC; ++c) {
        const int x = x0 + tx + 32 * c;
        if (x >= p.nx) continue;
#pragma unroll
        for (int q = 0; q < kVQ; ++q) {
            const int gy = y0 + ty * kVQ + q;
            if (gy < p.out_gy0 + p.out_rows) {
                const int64_t o = (int64_t)(gy - p.out_gy0) * p.ld_out + x;
                p.norm[o] = p.raw ? best[c][q] : fmaxf(best[c][q], 0.f);
                p.dir[o] = bdir[c][q];
            }
        }
    }
}

}  // namespace topo

using namespace topo;

extern "C" {

int topo_valley_ridge_f32(const float* dem_norm, int64_t ld_in, float* norm, float* dir, int64_t ld_out,
                          const topo_view* v, const float* bank, const int* bank_hw, const int64_t* bank_off,
                          const int* bank_cols, int n_angles, int n_ch, int hmax, int wmax, int group_flags, void* stream) {
    TOPO_CHECK(dem_norm && norm && dir && bank && bank_hw && bank_off && bank_cols, "null pointer");
    if (validate_view(v)) return -1;
    TOPO_CHECK(n_angles >= 1, "no angles");
    TOPO_CHECK(n_ch >= 1 && n_ch <= 4, "%d channels in one bank (1..4; longer flat lists run in groups, see topo_b200.h)", n_ch);
    TOPO_CHECK(group_flags >= 0 && group_flags <= 3, "bad group_flags");
    TOPO_CHECK(hmax >= 1 && wmax >= 1, "bad kernel extents");
    if (v->out_rows == 0) return 0;
    ValleyParams p;
    p.dem = dem_norm, p.norm = norm, p.dir = dir, p.ld_in = ld_in, p.ld_out = ld_out;
    p.nx = v->nx, p.gny = v->gny, p.in_gy0 = v->in_gy0, p.in_rows = v->in_rows;
    p.out_gy0 = v->out_gy0, p.out_rows = v->out_rows;
    p.bank = bank, p.bank_hw = bank_hw, p.bank_off = bank_off, p.n_angles = n_angles;
    p.bank_cols = reinterpret_cast<const int2*>(bank_cols);
    p.resume = group_flags & 1, p.raw = (group_flags >> 1) & 1;
    // anchors are h/2, w/2: rows above <= hmax/2; rows below h - 1 - h/2 <= h/2 <= hmax/2 for every h
    p.HT = hmax / 2, p.HB = hmax / 2;
    p.HL = wmax / 2, p.HR = wmax / 2;
    {
        const int lo = v->out_gy0 - p.HT > 0 ? v->out_gy0 - p.HT : 0;
        const int hi = v->out_gy0 + v->out_rows + p.HB < v->gny ? v->out_gy0 + v->out_rows + p.HB : v->gny;
        TOPO_CHECK(v->in_gy0 <= lo && v->in_gy0 + v->in_rows >= hi, "valley_ridge: band [%d,%d) does not cover rows [%d,%d)",
                   v->in_gy0, v->in_gy0 + v->in_rows, lo, hi);
    }
    const int tile_rows = kVTileY + p.HT + p.HB + 3, tile_cols = kVTileX + p.HL + p.HR;
    p.tile_pitch = tile_cols;
    const size_t smem = (size_t)tile_rows * p.tile_pitch * sizeof(float);
    const bool use_smem = smem <= 200 * 1024;
    dim3 grid(ceil_div(v->nx, kVTileX), ceil_div(v->out_rows, kVTileY));
    cudaStream_t s = (cudaStream_t)stream;
#define TOPO_VALLEY_LAUNCH(N)                                                                                       \
    do {                                                                                                            \
        topo::ProfScope prof__("valley_bank", s);                                                                  \
        if (use_smem) {                                                                                             \
            TOPO_CUDA(cudaFuncSetAttribute(valley_kernel<N, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                           227 * 1024));                                                            \
            valley_kernel<N, true><<<grid, dim3(32, 8), smem, s>>>(p);                                              \
        } else {                                                                                                    \
            valley_kernel<N, false><<<grid, dim3(32, 8), 0, s>>>(p);                                                \
        }                                                                                                           \
    } while (0)
    switch (n_ch) {
        case 1: TOPO_VALLEY_LAUNCH(1); break;
        case 2: TOPO_VALLEY_LAUNCH(2); break;
        case 3: TOPO_VALLEY_LAUNCH(3); break;
        default: TOPO_VALLEY_LAUNCH(4); break;
    }
#undef TOPO_VALLEY_LAUNCH
    TOPO_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
