// Image-plane post-processing of the T_b maps on the device: the Gaussian
// beam the workflow convolves its maps with and the NaN patching of failed pixels.
//
// Replaces
//   scipy.ndimage.gaussian_filter(emission_map, sigma=...)  called at
//     script/resample_with_ray_tracing.py:618-624 and script/pub/compare_on_off_scaling_factor.py:51-69
//   raytracingGRFF/util.py:6-77  patch_nan_emission_map / _patch_nan_2d
// Planes are independent (one per frequency); layout double[plane][ny][nx], x fastest.
#pragma once

#include "common.cuh"

namespace rtgrff {

// scipy's 'reflect' boundary (d c b a | a b c d | d c b a), any distance outside the line.
__device__ __forceinline__ int reflect_index(int i, int n)
{
    if ((unsigned)i < (unsigned)n) return i;
    const int p = 2 * n;
    i %= p;
    if (i < 0) i += p;
    return i < n ? i : p - 1 - i;
}

// One pass of the separable filter along `axis` (0: over rows i, 1: over columns j).  The
// summation order is scipy's symmetric correlate1d: centre tap first, then the tap pairs from the
// farthest to the nearest, (left + right) * w.  w[0..radius] = weights of offsets 0..radius.
__global__ void __launch_bounds__(256) gaussian_pass_kernel(const double *__restrict__ in, double *__restrict__ out,
                                                            const double *__restrict__ w, int radius, int ny, int nx,
                                                            int n_planes, int axis)
{
    const int64_t total = (int64_t)n_planes * ny * nx;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total;
         q += (int64_t)gridDim.x * blockDim.x) {
        const int j = (int)(q % nx), i = (int)((q / nx) % ny);
        const double *pl = in + (q - (int64_t)i * nx - j);
        double acc = pl[(int64_t)i * nx + j] * w[0];
        if (axis == 0) {
            for (int d = radius; d >= 1; --d) {
                const int a = reflect_index(i - d, ny), b = reflect_index(i + d, ny);
                acc += (pl[(int64_t)a * nx + j] + pl[(int64_t)b * nx + j]) * w[d];
            }
        } else {
            const double *row = pl + (int64_t)i * nx;
            for (int d = radius; d >= 1; --d)
                acc += (row[reflect_index(j - d, nx)] + row[reflect_index(j + d, nx)]) * w[d];
        }
        out[q] = acc;
    }
}

// For every pixel the value of the nearest finite pixel strictly to its right in the row (axis 1) or
// strictly above it in the column (axis 0) — NaN if there is none.  One thread per row / column,
// walking backwards.  These two directions point at pixels the sequential patching loop has not
// reached yet in a pass, so they only depend on the state at the start of the pass.
__global__ void next_finite_kernel(const double *__restrict__ a, double *__restrict__ right, double *__restrict__ up,
                                   int ny, int nx, int n_planes)
{
    const int64_t lines = (int64_t)n_planes * (ny + nx);
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < lines;
         q += (int64_t)gridDim.x * blockDim.x) {
        const int p = (int)(q / (ny + nx)), l = (int)(q % (ny + nx));
        const double *pl = a + (int64_t)p * ny * nx;
        double last = nan("");
        if (l < ny) {
            double *o = right + (int64_t)p * ny * nx + (int64_t)l * nx;
            const double *r = pl + (int64_t)l * nx;
            for (int j = nx - 1; j >= 0; --j) { o[j] = last; if (isfinite(r[j])) last = r[j]; }
        } else {
            const int j = l - ny;
            double *o = up + (int64_t)p * ny * nx + j;
            for (int i = ny - 1; i >= 0; --i) {
                o[(int64_t)i * nx] = last;
                if (isfinite(pl[(int64_t)i * nx + j])) last = pl[(int64_t)i * nx + j];
            }
        }
    }
}

// One pass of _patch_nan_2d (util.py:45-76) on one plane per block.  The reference walks the
// non-finite pixels in row-major order and patches IN PLACE, so a pixel sees the pixels patched
// before it in the same pass: its left neighbour is the pixel at j-1 if that one is finite now
// (a pixel that could not be patched has no finite pixel to its left either), likewise the pixel
// below; right and up come from the start-of-pass state.  That is a recurrence over (i, j-1) and
// (i-1, j): anti-diagonals i + j = const are independent and are processed one after the other.
// new value = mean of the available neighbours in the order left, right, down, up (np.mean of a
// short list: sequential sum, then one division).  fixed[plane] counts the patched pixels.
__global__ void __launch_bounds__(1024) patch_nan_pass_kernel(double *__restrict__ a, const unsigned char *__restrict__ bad,
                                                              const double *__restrict__ right,
                                                              const double *__restrict__ up, int ny, int nx,
                                                              int *__restrict__ fixed)
{
    const int p = blockIdx.x;
    double *pl = a + (int64_t)p * ny * nx;
    const unsigned char *bd = bad + (int64_t)p * ny * nx;
    const double *rt = right + (int64_t)p * ny * nx, *upv = up + (int64_t)p * ny * nx;
    int my_fixed = 0;
    for (int d = 0; d < ny + nx - 1; ++d) {
        const int i_lo = max(0, d - (nx - 1)), i_hi = min(ny - 1, d);
        for (int i = i_lo + (int)threadIdx.x; i <= i_hi; i += blockDim.x) {
            const int j = d - i;
            const int64_t o = (int64_t)i * nx + j;
            if (!bd[o]) continue;
            double sum = 0.0;
            int n = 0;
            if (j > 0 && isfinite(pl[o - 1])) { sum += pl[o - 1]; ++n; }
            if (isfinite(rt[o])) { sum += rt[o]; ++n; }
            if (i > 0 && isfinite(pl[o - nx])) { sum += pl[o - nx]; ++n; }
            if (isfinite(upv[o])) { sum += upv[o]; ++n; }
            if (n) { pl[o] = sum / (double)n; ++my_fixed; }
        }
        __syncthreads();
    }
    if (my_fixed) atomicAdd(fixed + p, my_fixed);
}

__global__ void mark_nonfinite_kernel(const double *__restrict__ a, unsigned char *__restrict__ bad, int64_t n,
                                      int *__restrict__ any_bad)
{
    int found = 0;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
        const unsigned char b = isfinite(a[q]) ? 0 : 1;
        bad[q] = b;
        found |= b;
    }
    if (found) atomicOr(any_bad, 1);
}

}  // namespace rtgrff
