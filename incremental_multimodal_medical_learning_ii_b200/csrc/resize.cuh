// Frame pre-processing on the GPU: Resize(size) -> CenterCrop(crop) of 8-bit grayscale frames with the exact integer
// arithmetic of Pillow's 8bpc resampler, so that the bytes fed to the stem are the bytes the reference feeds.
//
// Replaces (reference): transforms.Resize + transforms.CenterCrop on a PIL image in get_bio_vil_pipeline,
// DataRetrieval.py:175-180, and create_chest_xray_transform_for_inference,
// health_multimodal/image/data/transforms.py:30-41.  The arithmetic is Pillow's (requirements.txt:64, Pillow==9.3.0)
// src/libImaging/Resample.c: ImagingResampleHorizontal_8bpc followed by ImagingResampleVertical_8bpc with 22-bit
// fixed-point coefficients; the coefficient tables are built on the host in double exactly like precompute_coeffs /
// normalize_coeffs_8bpc (see build_resample_table in biovil_b200.cu).  Only the cropped window is computed: every
// output pixel of either pass depends on its own taps only, so skipping the cropped-away pixels changes nothing.
//
// Both passes are HBM/L2-bound byte work: one thread per output byte, taps read through the read-only path,
// consecutive threads on consecutive output columns (coalesced stores; the horizontal taps of neighbouring threads
// overlap, so a warp touches one contiguous span of the source row).
#pragma once
#include <stdint.h>

namespace bv {

constexpr int kResamplePrecisionBits = 32 - 8 - 2;

struct ResamplePass {
    const uint8_t* src;  // [n][src_rows][src_pitch]
    uint8_t* dst;        // [n][dst_rows][dst_pitch]
    const int* bounds;   // [out_size][2] = (first tap, tap count) along the resampled axis
    const int* kk;       // [out_size][ksize] fixed-point coefficients
    int ksize;
    int n;
    int src_rows, src_pitch;
    int dst_rows, dst_cols, dst_pitch;
    int out0;            // first output index along the resampled axis that is computed (crop offset)
    int other0;          // offset along the other axis into src (rows for the horizontal pass, columns for the vertical)
    int tap_origin;      // subtracted from the tap index (the vertical pass reads a temp image that starts at a later row)
};

__device__ __forceinline__ uint8_t resample_clip8(int ss) {
    const int v = ss >> kResamplePrecisionBits;
    return static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// dst[img][y][x] = clip8( 2^21 + sum_k src[img][other0 + y][xmin(out0 + x) + k] * kk[out0 + x][k] )
__global__ void __launch_bounds__(256) resample_horizontal_kernel(const ResamplePass p) {
    const long long total = static_cast<long long>(p.n) * p.dst_rows * p.dst_cols;
    for (long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; t < total;
         t += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int x = static_cast<int>(t % p.dst_cols);
        const long long r = t / p.dst_cols;
        const int y = static_cast<int>(r % p.dst_rows);
        const int img = static_cast<int>(r / p.dst_rows);
        const int xx = p.out0 + x;
        const int xmin = __ldg(p.bounds + 2 * xx) - p.tap_origin;
        const int cnt = __ldg(p.bounds + 2 * xx + 1);
        const uint8_t* s = p.src + (static_cast<size_t>(img) * p.src_rows + p.other0 + y) * p.src_pitch + xmin;
        const int* k = p.kk + static_cast<size_t>(xx) * p.ksize;
        int ss = 1 << (kResamplePrecisionBits - 1);
        for (int i = 0; i < cnt; ++i) ss += static_cast<int>(__ldg(s + i)) * __ldg(k + i);
        p.dst[(static_cast<size_t>(img) * p.dst_rows + y) * p.dst_pitch + x] = resample_clip8(ss);
    }
}

// dst[img][y][x] = clip8( 2^21 + sum_k src[img][ymin(out0 + y) - tap_origin + k][other0 + x] * kk[out0 + y][k] )
__global__ void __launch_bounds__(256) resample_vertical_kernel(const ResamplePass p) {
    const long long total = static_cast<long long>(p.n) * p.dst_rows * p.dst_cols;
    for (long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; t < total;
         t += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int x = static_cast<int>(t % p.dst_cols);
        const long long r = t / p.dst_cols;
        const int y = static_cast<int>(r % p.dst_rows);
        const int img = static_cast<int>(r / p.dst_rows);
        const int yy = p.out0 + y;
        const int ymin = __ldg(p.bounds + 2 * yy) - p.tap_origin;
        const int cnt = __ldg(p.bounds + 2 * yy + 1);
        const uint8_t* s = p.src + (static_cast<size_t>(img) * p.src_rows + ymin) * p.src_pitch + p.other0 + x;
        const int* k = p.kk + static_cast<size_t>(yy) * p.ksize;
        int ss = 1 << (kResamplePrecisionBits - 1);
        for (int i = 0; i < cnt; ++i) ss += static_cast<int>(__ldg(s + static_cast<size_t>(i) * p.src_pitch)) * __ldg(k + i);
        p.dst[(static_cast<size_t>(img) * p.dst_rows + y) * p.dst_pitch + x] = resample_clip8(ss);
    }
}

// plain window copy for the degenerate case where neither axis changes size
__global__ void __launch_bounds__(256) crop_copy_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int n,
                                                       int h, int w, int top, int left, int crop) {
    const long long total = static_cast<long long>(n) * crop * crop;
    for (long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; t < total;
         t += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int x = static_cast<int>(t % crop);
        const long long r = t / crop;
        const int y = static_cast<int>(r % crop);
        const int img = static_cast<int>(r / crop);
        dst[t] = __ldg(src + (static_cast<size_t>(img) * h + top + y) * w + left + x);
    }
}

}  // namespace bv
