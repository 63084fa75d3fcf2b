// Device-side rotation of the valley / ridge kernel bank.
// Reference: _rotate_kernels, topo.py:521-531 -- ndimage.rotate(kernel, angle, axes=(1, 2), reshape=True, order=2,
// mode="constant", cval=-9999), mask == -9999, z-score over the valid support, fill 0; followed by the channel mixing
// of the 3-D convolution (topo.py:431, 443).  At km-scale kernels the host call costs tens of seconds (180 angles x F
// kernels of up to 1133 x 1133 taps at size 801) -- here it is one launch per step for the whole bank.
//
// What stays on the host, on scipy itself, because it is cheap and then bit-identical by construction: cosdg / sindg,
// the rotated bounding boxes and offsets (scipy/ndimage/_interpolation.py rotate) and the quadratic-spline prefilter of
// the F source kernels (angle independent).  Restated here (profiles/proto/rotate_restated.py pins it bit for bit
// against scipy for every angle): NI_GeometricTransform's affine branch -- source coordinate with separate roundings
// (no fused multiply-add), "outside" test of mode="constant", the three quadratic B-spline weights per axis
// (w1 = 3/4 - x^2, w0 = (1/2 - x)^2 / 2, w2 = 1 - w0 - w1), nine taps (c * wy) * wx summed row-major with mirrored
// coefficient indices at the edges, result rounded to float32.
#include <math.h>

#include "common.cuh"

namespace topo {

struct RotAngle {
    double m00, m01, m10, m11, off0, off1;  // rotation matrix and offset of scipy.ndimage.rotate
    long long out_off;                      // element offset of this angle's [F][oh][ow] block
    int oh, ow;
};

__device__ __forceinline__ int mirror_index(int i, int n) {
    if (n == 1) return 0;
    const int p = 2 * n - 2;
    i = i < 0 ? -i : i;
    i %= p;
    return i >= n ? p - i : i;
}

// one thread per output pixel of one angle; all F kernels share the coordinate and the weights
__global__ void __launch_bounds__(256) rotate_bank_kernel(const double* __restrict__ coef, int F, int H, int W,
                                                          const RotAngle* __restrict__ angles, float cval, float* __restrict__ out) {
    const RotAngle a = angles[blockIdx.y];
    const long long npix = (long long)a.oh * a.ow;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < npix; i += (long long)gridDim.x * 256) {
        const int oy = (int)(i / a.ow), ox = (int)(i - (long long)oy * a.ow);
        // icoor[h] = (0 + o_y * m[h][0]) + o_x * m[h][1]; icoor[h] += shift[h]
        const double cy = __dadd_rn(__dadd_rn(__dmul_rn((double)oy, a.m00), __dmul_rn((double)ox, a.m01)), a.off0);
        const double cx = __dadd_rn(__dadd_rn(__dmul_rn((double)oy, a.m10), __dmul_rn((double)ox, a.m11)), a.off1);
        float* o = out + a.out_off + i;
        if (cy < 0.0 || cy > (double)(H - 1) || cx < 0.0 || cx > (double)(W - 1)) {
            for (int f = 0; f < F; ++f) o[(long long)f * npix] = cval;
            continue;
        }
        const double fy = floor(__dadd_rn(cy, 0.5)), fx = floor(__dadd_rn(cx, 0.5));
        const double xy = __dadd_rn(cy, -fy), xx = __dadd_rn(cx, -fx);
        double wy[3], wx[3];
        wy[1] = __dadd_rn(0.75, -__dmul_rn(xy, xy));
        double t = __dadd_rn(0.5, -xy);
        wy[0] = __dmul_rn(__dmul_rn(0.5, t), t);
        wy[2] = __dadd_rn(__dadd_rn(1.0, -wy[0]), -wy[1]);
        wx[1] = __dadd_rn(0.75, -__dmul_rn(xx, xx));
        t = __dadd_rn(0.5, -xx);
        wx[0] = __dmul_rn(__dmul_rn(0.5, t), t);
        wx[2] = __dadd_rn(__dadd_rn(1.0, -wx[0]), -wx[1]);
        int yi[3], xj[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) yi[k] = mirror_index((int)fy - 1 + k, H), xj[k] = mirror_index((int)fx - 1 + k, W);
        for (int f = 0; f < F; ++f) {
            const double* c = coef + (long long)f * H * W;
            double acc = 0.0;
#pragma unroll
            for (int p = 0; p < 3; ++p)
#pragma unroll
                for (int q = 0; q < 3; ++q)
                    acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(__ldg(c + (long long)yi[p] * W + xj[q]), wy[p]), wx[q]));
            o[(long long)f * npix] = (float)acc;
        }
    }
}

// z-score over the valid support (numpy: (v - v.mean()) / v.std(), float64, two passes), invalid -> 0; in place.
// One CTA per (angle, kernel); fixed reduction order.
__device__ double block_sum(double v) {
    __shared__ double sm[8];
    __shared__ double total;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
    __syncthreads();
    if (lane == 0) sm[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += sm[w];
        total = t;
    }
    __syncthreads();
    return total;
}

__global__ void __launch_bounds__(256) rotate_zscore_kernel(const RotAngle* __restrict__ angles, int F, float cval,
                                                            float* __restrict__ data) {
    const RotAngle a = angles[blockIdx.y];
    const long long npix = (long long)a.oh * a.ow;
    float* k = data + a.out_off + (long long)blockIdx.x * npix;
    double s = 0.0, n = 0.0;
    for (long long i = threadIdx.x; i < npix; i += 256) {
        const float v = k[i];
        if (v != cval) s += (double)v, n += 1.0;
    }
    s = block_sum(s);
    n = block_sum(n);
    const double mean = s / n;
    double ss = 0.0;
    for (long long i = threadIdx.x; i < npix; i += 256) {
        const float v = k[i];
        if (v != cval) {
            const double d = (double)v - mean;
            ss += d * d;
        }
    }
    ss = block_sum(ss);
    const double sd = sqrt(ss / n);
    for (long long i = threadIdx.x; i < npix; i += 256) {
        const float v = k[i];
        k[i] = v != cval ? (float)(((double)v - mean) / sd) : 0.f;
    }
    (void)F;
}

// channel mixing of the reference's 3-D convolution: out[m] = sum of k_j over 0 <= m + s - j <= F-1, s = (F-1)/2
__global__ void __launch_bounds__(256) rotate_mix_kernel(const RotAngle* __restrict__ angles, int F, const float* __restrict__ in,
                                                         float* __restrict__ out) {
    const RotAngle a = angles[blockIdx.y];
    const long long npix = (long long)a.oh * a.ow;
    const int s = (F - 1) / 2;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < npix; i += (long long)gridDim.x * 256) {
        for (int m = 0; m < F; ++m) {
            const int lo = m + s - (F - 1) > 0 ? m + s - (F - 1) : 0, hi = m + s < F - 1 ? m + s : F - 1;
            double acc = 0.0;
            for (int j = lo; j <= hi; ++j) acc += (double)in[a.out_off + (long long)j * npix + i];
            out[a.out_off + (long long)m * npix + i] = (float)acc;
        }
    }
}

}  // namespace topo

using namespace topo;

extern "C" {

int topo_rotate_bank_f32(const double* coef, int n_kernels, int h, int w, const void* angles, int n_angles, float cval,
                         float* scratch, float* out, void* stream) {
    TOPO_CHECK(coef && angles && scratch && out, "null pointer");
    TOPO_CHECK(n_kernels >= 1 && h >= 1 && w >= 1 && n_angles >= 1 && n_angles <= 65535, "bad bank shape");
    cudaStream_t s = (cudaStream_t)stream;
    const RotAngle* ang = reinterpret_cast<const RotAngle*>(angles);
    const int blocks = kNumSMs * 2;
    TOPO_LAUNCH("rotate_bank", s, rotate_bank_kernel<<<dim3(blocks, n_angles), 256, 0, s>>>(coef, n_kernels, h, w, ang, cval, scratch));
    TOPO_LAUNCH("rotate_zscore", s, rotate_zscore_kernel<<<dim3(n_kernels, n_angles), 256, 0, s>>>(ang, n_kernels, cval, scratch));
    TOPO_LAUNCH("rotate_mix", s, rotate_mix_kernel<<<dim3(blocks, n_angles), 256, 0, s>>>(ang, n_kernels, scratch, out));
    return 0;
}

}  // extern "C"
