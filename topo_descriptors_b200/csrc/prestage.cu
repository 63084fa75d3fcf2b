// Pre-stage of the reference script (scripts/compute_topo_descriptors.py:17-19) on the device:
//   helpers.py:17-31  get_dem_netcdf: `dem.where(dem > CFG.min_elevation)`      -> the `mask_below` threshold
//   helpers.py:137-154 fill_na: ind_nans = np.where(np.isnan(dem));
//                      dem.interpolate_na(dim="x", method="nearest", fill_value="extrapolate")
// xarray hands every row to scipy.interpolate.interp1d(kind="nearest", fill_value="extrapolate") over the valid
// samples: a missing cell takes the value of the valid cell of its row that is nearest in x, exact half-way points
// go to the neighbour with the LOWER x ("rounds half down"), cells beyond the first / last valid one take that
// one, rows without any valid cell stay NaN.
//
// One CTA per row.  The validity of the row is condensed into one ballot word per 32 cells in shared memory; two
// warp-level scans over those words give, per word, the last valid cell before it and the first valid cell after
// it; a second sweep over the row (L1/L2 hits) resolves every missing cell with two bit tricks + at most two
// shared-memory reads, independent of the length of the gap (sea / lake masks are thousands of cells long).
// 8 B/px of HBM traffic (read + write); the index extraction is a second 4 B/px pass.
#include <limits.h>
#include <math.h>

#include "common.cuh"

namespace topo {

constexpr int kRowThreads = 256;
constexpr int kRowWarps = kRowThreads / 32;

__device__ __forceinline__ bool cell_valid(float v, int use_mask, float mask_below) {
    return !isnan(v) && (!use_mask || v > mask_below);
}

// masks[c] = ballot of valid cells of chunk c (32 cells); returns this thread's share of the missing-cell count
__device__ __forceinline__ int stage_row_masks(const float* row, int nx, int nchunks, int use_mask, float mask_below,
                                               uint32_t* masks) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int missing = 0;
    for (int c = warp; c < nchunks; c += kRowWarps) {
        const int i = c * 32 + lane;
        const bool ok = i < nx && cell_valid(__ldg(row + i), use_mask, mask_below);
        const uint32_t m = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) {
            masks[c] = m;
            const int cells = min(32, nx - c * 32);
            missing += cells - __popc(m);
        }
    }
    return missing;
}

__global__ void __launch_bounds__(kRowThreads) fill_na_rows_kernel(const float* __restrict__ in, int64_t ld_in,
                                                                   float* __restrict__ out, int64_t ld_out, int nx,
                                                                   const double* __restrict__ x, int use_mask,
                                                                   float mask_below, int* __restrict__ row_missing) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nchunks = (nx + 31) / 32;
    uint32_t* masks = reinterpret_cast<uint32_t*>(smem_raw);
    int* last_before = reinterpret_cast<int*>(masks + nchunks);  // last valid cell in chunks < c, or -1
    int* first_after = last_before + nchunks;                     // first valid cell in chunks > c, or INT_MAX
    __shared__ int missing_total;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* row = in + (int64_t)blockIdx.x * ld_in;
    float* orow = out + (int64_t)blockIdx.x * ld_out;

    if (threadIdx.x == 0) missing_total = 0;
    __syncthreads();
    const int missing = stage_row_masks(row, nx, nchunks, use_mask, mask_below, masks);
    if (missing) atomicAdd(&missing_total, missing);
    __syncthreads();

    if (warp == 0) {  // exclusive running maximum, left to right
        int carry = -1;
        for (int base = 0; base < nchunks; base += 32) {
            const int c = base + lane;
            const uint32_t m = c < nchunks ? masks[c] : 0u;
            int v = m ? c * 32 + 31 - __clz(m) : -1;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int o = __shfl_up_sync(0xffffffffu, v, d);
                if (lane >= d) v = max(v, o);
            }
            int excl = __shfl_up_sync(0xffffffffu, v, 1);
            if (lane == 0) excl = -1;
            if (c < nchunks) last_before[c] = max(carry, excl);
            carry = max(carry, __shfl_sync(0xffffffffu, v, 31));
        }
    } else if (warp == 1) {  // exclusive running minimum, right to left
        int carry = INT_MAX;
        for (int base = ((nchunks - 1) / 32) * 32; base >= 0; base -= 32) {
            const int c = base + lane;
            const uint32_t m = c < nchunks ? masks[c] : 0u;
            int v = m ? c * 32 + __ffs(m) - 1 : INT_MAX;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int o = __shfl_down_sync(0xffffffffu, v, d);
                if (lane + d < 32) v = min(v, o);
            }
            int excl = __shfl_down_sync(0xffffffffu, v, 1);
            if (lane == 31) excl = INT_MAX;
            if (c < nchunks) first_after[c] = min(carry, excl);
            carry = min(carry, __shfl_sync(0xffffffffu, v, 0));
        }
    }
    __syncthreads();
    if (threadIdx.x == 0 && row_missing) row_missing[blockIdx.x] = missing_total;

    for (int c = warp; c < nchunks; c += kRowWarps) {
        const int i = c * 32 + lane;
        if (i >= nx) continue;
        const uint32_t m = masks[c];
        float v = __ldg(row + i);
        if (!((m >> lane) & 1u)) {
            const uint32_t lm = m & ((2u << lane) - 1u);  // valid cells at or before this lane (lane 31: all)
            const uint32_t rm = m >> lane;                // valid cells at or after this lane
            const int li = lm ? c * 32 + 31 - __clz(lm) : last_before[c];
            const int ri = rm ? i + __ffs(rm) - 1 : first_after[c];
            int pick;
            if (li < 0 && ri == INT_MAX) {
                pick = -1;  // no valid cell in this row: stays missing
            } else if (li < 0) {
                pick = ri;
            } else if (ri == INT_MAX) {
                pick = li;
            } else if (x) {
                const double xq = x[i], xl = x[li], xr = x[ri];
                const double dl = fabs(xq - xl), dr = fabs(xr - xq);
                pick = dl < dr ? li : dr < dl ? ri : (xl < xr ? li : ri);  // ties: the lower-x neighbour
            } else {
                pick = (i - li) <= (ri - i) ? li : ri;  // uniform ascending x
            }
            v = pick >= 0 ? __ldg(row + pick) : __int_as_float(0x7fc00000);
        }
        orow[i] = v;
    }
}

// exclusive prefix sum of the per-row counts (one CTA; rows is at most a few 10^4)
__global__ void __launch_bounds__(1024) row_offsets_kernel(const int* __restrict__ counts, int rows,
                                                           int64_t* __restrict__ offsets) {
    __shared__ int64_t warp_sums[32];
    __shared__ int64_t carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < rows; base += 1024) {
        const int i = base + threadIdx.x;
        const int64_t own = i < rows ? counts[i] : 0;
        int64_t v = own;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int64_t o = __shfl_up_sync(0xffffffffu, v, d);
            if (lane >= d) v += o;
        }
        if (lane == 31) warp_sums[warp] = v;
        __syncthreads();
        if (warp == 0) {
            int64_t w = warp_sums[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int64_t o = __shfl_up_sync(0xffffffffu, w, d);
                if (lane >= d) w += o;
            }
            warp_sums[lane] = w;  // inclusive over warps
        }
        __syncthreads();
        const int64_t before = carry_s + (warp ? warp_sums[warp - 1] : 0);
        if (i < rows) offsets[i] = before + v - own;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = before + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) offsets[rows] = carry_s;
}

// (row, col) of every missing cell in row-major order == np.where(np.isnan(dem))
__global__ void __launch_bounds__(kRowThreads) nan_indices_kernel(const float* __restrict__ in, int64_t ld_in, int nx,
                                                                  int use_mask, float mask_below,
                                                                  const int64_t* __restrict__ offsets,
                                                                  int* __restrict__ out_rows, int* __restrict__ out_cols) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nchunks = (nx + 31) / 32;
    uint32_t* masks = reinterpret_cast<uint32_t*>(smem_raw);
    int* before = reinterpret_cast<int*>(masks + nchunks);  // missing cells of this row in chunks < c
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t base_pos = offsets[blockIdx.x];
    if (offsets[blockIdx.x + 1] == base_pos) return;  // nothing missing in this row (uniform for the CTA)
    const float* row = in + (int64_t)blockIdx.x * ld_in;
    stage_row_masks(row, nx, nchunks, use_mask, mask_below, masks);
    __syncthreads();
    if (warp == 0) {
        int carry = 0;
        for (int b = 0; b < nchunks; b += 32) {
            const int c = b + lane;
            const int cells = c < nchunks ? min(32, nx - c * 32) : 0;
            const int own = c < nchunks ? cells - __popc(masks[c]) : 0;
            int v = own;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int o = __shfl_up_sync(0xffffffffu, v, d);
                if (lane >= d) v += o;
            }
            if (c < nchunks) before[c] = carry + v - own;
            carry += __shfl_sync(0xffffffffu, v, 31);
        }
    }
    __syncthreads();
    for (int c = warp; c < nchunks; c += kRowWarps) {
        const int i = c * 32 + lane;
        const uint32_t cells = nx - c * 32 >= 32 ? 0xffffffffu : ((1u << (nx - c * 32)) - 1u);
        const uint32_t miss = ~masks[c] & cells;
        if ((miss >> lane) & 1u) {
            const int64_t pos = base_pos + before[c] + __popc(miss & ((1u << lane) - 1u));
            out_rows[pos] = blockIdx.x;
            out_cols[pos] = i;
        }
    }
}

}  // namespace topo

using namespace topo;

extern "C" {

int topo_fill_na_f32(const float* dem, int64_t ld_in, float* out, int64_t ld_out, int rows, int nx, const double* x,
                     int use_mask, float mask_below, int* row_missing, void* stream) {
    TOPO_CHECK(dem && out, "null pointer");
    TOPO_CHECK(rows >= 0 && nx >= 1 && ld_in >= nx && ld_out >= nx, "bad geometry");
    TOPO_CHECK(nx <= 32 * 4096, "rows longer than 131072 cells are not supported");
    if (rows == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t smem = (size_t)((nx + 31) / 32) * 3 * sizeof(int);
    if (smem > 40 * 1024) TOPO_CUDA(cudaFuncSetAttribute(fill_na_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    TOPO_LAUNCH("fill_na_rows", s,
                fill_na_rows_kernel<<<rows, kRowThreads, smem, s>>>(dem, ld_in, out, ld_out, nx, x, use_mask, mask_below,
                                                                  row_missing));
    return 0;
}

int topo_nan_indices_f32(const float* dem, int64_t ld_in, int rows, int nx, int use_mask, float mask_below,
                         const int* row_missing, int64_t* row_offsets, int* out_rows, int* out_cols, void* stream) {
    TOPO_CHECK(dem && row_missing && row_offsets, "null pointer");
    TOPO_CHECK(rows >= 0 && nx >= 1 && ld_in >= nx, "bad geometry");
    TOPO_CHECK(nx <= 32 * 4096, "rows longer than 131072 cells are not supported");
    cudaStream_t s = (cudaStream_t)stream;
    if (!out_rows || !out_cols) {  // first call: offsets only (row_offsets[rows] = total, the caller sizes the outputs)
        TOPO_LAUNCH("row_offsets", s, row_offsets_kernel<<<1, 1024, 0, s>>>(row_missing, rows, row_offsets));
        return 0;
    }
    if (rows == 0) return 0;
    const size_t smem = (size_t)((nx + 31) / 32) * 2 * sizeof(int);
    TOPO_LAUNCH("nan_indices", s,
                nan_indices_kernel<<<rows, kRowThreads, smem, s>>>(dem, ld_in, nx, use_mask, mask_below, row_offsets,
                                                                 out_rows, out_cols));
    return 0;
}

}  // extern "C"
