// HBM-bound helper kernels of the BioViL image path: stem patch gather, NHWC max-pool, global average pool,
// projector tail (128->128 conv + patch mean + L2 norms) fused with the image x prompt cosine scorer.
// All of them are plain CUDA-core kernels: the work is bytes, not FLOPs.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace bv {

constexpr int kEmbDim = 128;  // JOINT_FEATURE_SIZE, health_multimodal/image/model/model.py:25

// ----------------------------------------------------------------------------------------------
// Stem patch gather: 7x7 stride-2 pad-3 windows of the input frame -> bf16 rows of a [M, K] matrix
// (M = B*Ho*Wo, K = 64 for one input channel [49 taps + 15 zeros], 192 for three [147 + 45 zeros]).
// k = c*49 + r*7 + s, the flatten order of the OIHW stem weight (resnet.py:34 conv1).
// The u8 path keeps pixel integers exact in bf16; 1/255 (ToTensor, transforms.py:37) lives in the weights.
// ----------------------------------------------------------------------------------------------
template <typename TIn>
__device__ __forceinline__ float stem_px(const TIn* p) {
    return static_cast<float>(*p);
}

template <typename TIn, int CIN>
__global__ void __launch_bounds__(256) stem_patch_kernel(const TIn* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                        int B, int H, int W, int Ho, int Wo) {
    constexpr int K = (CIN == 1) ? 64 : 192;
    constexpr int CHUNKS = K / 8;
    const long long total = static_cast<long long>(B) * Ho * Wo * CHUNKS;
    for (long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; t < total;
         t += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int chunk = static_cast<int>(t % CHUNKS);
        const long long pix = t / CHUNKS;
        const int q = static_cast<int>(pix % Wo);
        const int pr = static_cast<int>((pix / Wo) % Ho);
        const int b = static_cast<int>(pix / (static_cast<long long>(Wo) * Ho));
        uint32_t packed[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float v[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int k = chunk * 8 + e * 2 + u;
                float val = 0.0f;
                if (k < CIN * 49) {
                    const int c = k / 49;
                    const int tap = k - c * 49;
                    const int r = tap / 7;
                    const int s = tap - r * 7;
                    const int iy = pr * 2 - 3 + r;
                    const int ix = q * 2 - 3 + s;
                    if (iy >= 0 && iy < H && ix >= 0 && ix < W)
                        val = stem_px(x + ((static_cast<size_t>(b) * CIN + c) * H + iy) * W + ix);
                }
                v[u] = val;
            }
            const __nv_bfloat162 h = __floats2bfloat162_rn(v[0], v[1]);
            packed[e] = *reinterpret_cast<const uint32_t*>(&h);
        }
        reinterpret_cast<uint4*>(out)[t] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
    }
}

// ----------------------------------------------------------------------------------------------
// MaxPool2d(kernel 3, stride 2, padding 1) on NHWC bf16 (resnet.py:37).  One thread = one output pixel x 8
// channels (16-byte vectors, coalesced over channels).  Padding never wins: inputs are post-ReLU (>= 0) and
// the centre tap is always in range.
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) maxpool3x3s2_nhwc_kernel(const __nv_bfloat16* __restrict__ in,
                                                               __nv_bfloat16* __restrict__ out, int B, int H, int W,
                                                               int C, int Ho, int Wo) {
    const int chunks = C / 8;
    const long long total = static_cast<long long>(B) * Ho * Wo * chunks;
    for (long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; t < total;
         t += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int ch = static_cast<int>(t % chunks);
        const long long pix = t / chunks;
        const int q = static_cast<int>(pix % Wo);
        const int pr = static_cast<int>((pix / Wo) % Ho);
        const int b = static_cast<int>(pix / (static_cast<long long>(Wo) * Ho));
        __nv_bfloat162 m[4];
        bool first = true;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int iy = pr * 2 - 1 + r;
            if (iy < 0 || iy >= H) continue;
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                const int ix = q * 2 - 1 + s;
                if (ix < 0 || ix >= W) continue;
                const uint4 v = __ldg(reinterpret_cast<const uint4*>(
                    in + ((static_cast<size_t>(b) * H + iy) * W + ix) * C + ch * 8));
                const __nv_bfloat162* hv = reinterpret_cast<const __nv_bfloat162*>(&v);
                if (first) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) m[e] = hv[e];
                    first = false;
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e) m[e] = __hmax2(m[e], hv[e]);
                }
            }
        }
        uint4 o;
        o.x = *reinterpret_cast<uint32_t*>(&m[0]);
        o.y = *reinterpret_cast<uint32_t*>(&m[1]);
        o.z = *reinterpret_cast<uint32_t*>(&m[2]);
        o.w = *reinterpret_cast<uint32_t*>(&m[3]);
        reinterpret_cast<uint4*>(out)[t] = o;
    }
}

// ----------------------------------------------------------------------------------------------
// adaptive_avg_pool2d(x4, (1,1)) + flatten (model.py:201): [B, P, C] bf16 -> [B, C] fp32.
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) avgpool_nhwc_kernel(const __nv_bfloat16* __restrict__ in,
                                                          float* __restrict__ out, int B, int P, int C) {
    const int chunks = C / 8;
    const long long total = static_cast<long long>(B) * chunks;
    const float inv = 1.0f / static_cast<float>(P);
    for (long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; t < total;
         t += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int ch = static_cast<int>(t % chunks);
        const int b = static_cast<int>(t / chunks);
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const __nv_bfloat16* base = in + static_cast<size_t>(b) * P * C + ch * 8;
        for (int pidx = 0; pidx < P; ++pidx) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(base + static_cast<size_t>(pidx) * C));
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                acc[2 * e] += __uint_as_float(w[e] << 16);
                acc[2 * e + 1] += __uint_as_float(w[e] & 0xFFFF0000u);
            }
        }
        float4* o = reinterpret_cast<float4*>(out + static_cast<size_t>(b) * C + ch * 8);
        o[0] = make_float4(acc[0] * inv, acc[1] * inv, acc[2] * inv, acc[3] * inv);
        o[1] = make_float4(acc[4] * inv, acc[5] * inv, acc[6] * inv, acc[7] * inv);
    }
}

// ----------------------------------------------------------------------------------------------
// Prompt preparation: y[j] / ||y[j]||_2 for every prompt vector (plain division, no eps: the semantics of
// torchmetrics pairwise_cosine_similarity used by Trainer.myCosineSimilarity, Trainer.py:1682-1704).
// One warp per prompt.
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) prompt_normalize_kernel(const float* __restrict__ y, float* __restrict__ yn,
                                                              int NP) {
    const int j = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (j >= NP) return;
    const float4 v = reinterpret_cast<const float4*>(y + static_cast<size_t>(j) * kEmbDim)[lane];
    float ss = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float nrm = sqrtf(ss);
    reinterpret_cast<float4*>(yn + static_cast<size_t>(j) * kEmbDim)[lane] =
        make_float4(v.x / nrm, v.y / nrm, v.z / nrm, v.w / nrm);
}

struct ScoreOut {
    float* sim;     // [B, L, 2]  (pos, neg) cosine
    float* prob;    // [B, L]     sigmoid(pos - neg) = softmax over {pos, neg} at pos
    uint8_t* pred;  // [B, L]     1 iff pos > neg (argmax([neg, pos]), ties -> 0; Trainer.py:836)
    float* score;   // [B, L]     (pos + 1) / 2 (Trainer.py:825)
};

// One warp scores one image embedding `x` (128 fp32, 4 per lane, already divided by ||x||) against
// yn[L][2][P][128] unit prompt vectors: cosine per prompt, max over the P prompts of a polarity
// (P == 1 when the prompts were mean-reduced beforehand, Trainer.py:1665-1666; P > 1 is the MAX_EMB
// variant, Trainer.py:1691-1694).
__device__ __forceinline__ void score_image_warp(const float4 xn, const float* __restrict__ yn, int L, int P, int b,
                                                 const ScoreOut& o, int lane) {
    for (int l = 0; l < L; ++l) {
        float best[2];
#pragma unroll
        for (int pol = 0; pol < 2; ++pol) {
            float m = -INFINITY;
            bool any_nan = false;
            for (int pp = 0; pp < P; ++pp) {
                const float4 y =
                    reinterpret_cast<const float4*>(yn + (static_cast<size_t>(l * 2 + pol) * P + pp) * kEmbDim)[lane];
                float d = xn.x * y.x + xn.y * y.y + xn.z * y.z + xn.w * y.w;
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) d += __shfl_xor_sync(0xffffffffu, d, off);
                any_nan |= (d != d);
                m = fmaxf(m, d);
            }
            best[pol] = any_nan ? __int_as_float(0x7fc00000) : m;
        }
        if (lane == 0) {
            const float pos = best[0], neg = best[1];
            const size_t i = static_cast<size_t>(b) * L + l;
            if (o.sim) {
                o.sim[2 * i] = pos;
                o.sim[2 * i + 1] = neg;
            }
            if (o.prob) o.prob[i] = 1.0f / (1.0f + expf(-(pos - neg)));
            // torch.argmax(cat([neg, pos])) (Trainer.py:836): NaN counts as the maximum, the FIRST maximum wins - a zero
            // embedding or prompt (0/0 cosine) must give the reference's label, not just "pos > neg"
            if (o.pred) o.pred[i] = (pos != pos) ? ((neg != neg) ? 0 : 1) : ((pos > neg) ? 1 : 0);
            if (o.score) o.score[i] = (pos + 1.0f) * 0.5f;
        }
    }
}

// Stand-alone scorer for cached embeddings (the Trainer.val/test path): emb [B,128] un-normalised.
__global__ void __launch_bounds__(256) score_kernel(const float* __restrict__ emb, const float* __restrict__ yn, int B,
                                                   int L, int P, ScoreOut o) {
    const int warps = (gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    for (int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < B; b += warps) {
        float4 x = reinterpret_cast<const float4*>(emb + static_cast<size_t>(b) * kEmbDim)[lane];
        float ss = x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
        const float nrm = sqrtf(ss);
        x = make_float4(x.x / nrm, x.y / nrm, x.z / nrm, x.w / nrm);
        score_image_warp(x, yn, L, P, b, o, lane);
    }
}

// Trainer.myCosineSimilarity(x, y) as one launch (Trainer.py:1682-1704): out[b, p] = cos(x[b], y[p]) for x [B,128],
// y [P,128], both un-normalised (rows divided by their L2 norm, plain division - torchmetrics semantics); with
// `reduce_max` the max over the P prompts (the MAX_EMB branch, :1691-1694) -> out [B,1].  One warp per image row; the
// prompt norms are recomputed per warp (P is 1..8: cheaper than a second launch or a device allocation).
__global__ void __launch_bounds__(256) pairwise_cosine_kernel(const float* __restrict__ x, const float* __restrict__ y, int B,
                                                             int P, int reduce_max, float* __restrict__ out) {
    const int warps = (gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    for (int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < B; b += warps) {
        float4 xv = reinterpret_cast<const float4*>(x + static_cast<size_t>(b) * kEmbDim)[lane];
        float ss = xv.x * xv.x + xv.y * xv.y + xv.z * xv.z + xv.w * xv.w;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
        const float nx = sqrtf(ss);
        xv = make_float4(xv.x / nx, xv.y / nx, xv.z / nx, xv.w / nx);
        float best = -INFINITY;
        bool any_nan = false;
        for (int pp = 0; pp < P; ++pp) {
            float4 yv = reinterpret_cast<const float4*>(y + static_cast<size_t>(pp) * kEmbDim)[lane];
            float sy = yv.x * yv.x + yv.y * yv.y + yv.z * yv.z + yv.w * yv.w;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) sy += __shfl_xor_sync(0xffffffffu, sy, off);
            const float ny = sqrtf(sy);
            yv = make_float4(yv.x / ny, yv.y / ny, yv.z / ny, yv.w / ny);
            float d = xv.x * yv.x + xv.y * yv.y + xv.z * yv.z + xv.w * yv.w;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) d += __shfl_xor_sync(0xffffffffu, d, off);
            if (!reduce_max) {
                if (lane == 0) out[static_cast<size_t>(b) * P + pp] = d;
            } else {
                any_nan |= (d != d);
                best = fmaxf(best, d);
            }
        }
        if (reduce_max && lane == 0) out[b] = any_nan ? __int_as_float(0x7fc00000) : best;
    }
}

// ----------------------------------------------------------------------------------------------
// Float frames -> 8-bit frames when they ARE 8-bit data: ToTensor + ExpandChannels (transforms.py:12-38) hand the
// model k/255 with k a byte and three identical channels.  One pass reads the fp32 frames [B,C,H,W] (C = 1 or 3),
// writes k as uint8 [B,1,H,W] and raises *bad (zeroed by the caller) when any pixel is not within 1e-3 grey levels of an integer in 0..255 or
// the channels differ - the caller then takes the general float stem instead.  Replaces five torch reductions with
// three host synchronisations by one launch and one flag read.  4 pixels per thread (16-byte loads, 4-byte stores).
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) quantize_frames_kernel(const float* __restrict__ x, uint8_t* __restrict__ out, int C,
                                                             long long hw4, long long total4, int* __restrict__ bad) {
    bool good = true;
    for (long long t = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; t < total4;
         t += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long b = t / hw4, r = t - b * hw4;
        const float4* src = reinterpret_cast<const float4*>(x) + b * C * hw4 + r;
        const float4 v = __ldg(src);
        const float f[4] = {v.x, v.y, v.z, v.w};
        for (int c = 1; c < C; ++c) {
            const float4 w = __ldg(src + c * hw4);
            good = good && (w.x == v.x) && (w.y == v.y) && (w.z == v.z) && (w.w == v.w);
        }
        uint32_t packed = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float k255 = f[e] * 255.0f;
            const float k = rintf(k255);
            good = good && (fabsf(k255 - k) <= 1e-3f) && (k >= 0.0f) && (k <= 255.0f);   // NaN fails the first test
            packed |= (static_cast<uint32_t>(fminf(fmaxf(k, 0.0f), 255.0f)) & 0xFFu) << (8 * e);
        }
        reinterpret_cast<uint32_t*>(out)[t] = packed;
    }
    if (!__all_sync(0xffffffffu, good) && (threadIdx.x & 31) == 0) atomicOr(bad, 1);
}

// ----------------------------------------------------------------------------------------------
// Projector tail + embeddings + scoring, one CTA (256 threads) per image.
//   hid [B*P, 128] fp32  = ReLU(BN(conv1x1_2048->128(x4)))   (written by the tcgen05 GEMM, modules.py:43-46)
//   proj[p, d] = b2[d] + sum_k hid[p, k] * W2[d, k]            (modules.py:47, fp32 FMA)
//   global[d]  = mean_p proj[p, d]                             (model.py:145)  -> global_out [B,128] un-normalised
//   patch_out [B, P, 128] = proj or proj / max(||proj||, 1e-12) (model.py:172-174, F.normalize eps)
//   heat_out  [B, P, L]   = normalised patch . unit prompt(l, pos, first prompt)  (vlp/inference_engine.py:107)
//   scores of the global embedding against the prepared prompts (Trainer.py:805-837).
// ----------------------------------------------------------------------------------------------
struct HeadParams {
    const float* hid;
    const float* w2t;  // [128 k][128 d] fp32: W2 transposed so that thread d reads consecutive addresses
    const float* b2;
    int B, P;
    float* global_out;
    float* patch_out;
    int normalize_patch;
    const float* yn;  // prepared prompts or nullptr
    int L, NPP;       // labels, prompts per polarity
    ScoreOut score;
    float* heat_out;      // [B, P, L] or nullptr
    const float* heat_t;  // [L, 128] unit text vectors for heat-maps
};

// Global-embedding-only variant: the projector's last conv and the patch mean are both linear, so
//   mean_p (W2 h_p + b2) = W2 (mean_p h_p) + b2
// and the 225 matrix-vector products collapse into one (model.py:144-145 computed per patch only because the
// reference materialises the patch embeddings).  Used when neither patch embeddings nor heat-maps are requested.
__global__ void __launch_bounds__(128) head_global_kernel(const HeadParams hp) {
    __shared__ float hbar[kEmbDim];
    __shared__ float gvec[kEmbDim];
    const int b = blockIdx.x;
    const int d = threadIdx.x;
    const int lane = d & 31;
    asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");   // programmatic dependent launch, see ptx.cuh
    asm volatile("griddepcontrol.wait;\n" ::: "memory");
    const float* hp_b = hp.hid + static_cast<size_t>(b) * hp.P * kEmbDim + d;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    int pidx = 0;
    for (; pidx + 4 <= hp.P; pidx += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[u] += __ldg(hp_b + static_cast<size_t>(pidx + u) * kEmbDim);
    }
    for (; pidx < hp.P; ++pidx) acc[0] += __ldg(hp_b + static_cast<size_t>(pidx) * kEmbDim);
    hbar[d] = ((acc[0] + acc[1]) + (acc[2] + acc[3])) / static_cast<float>(hp.P);
    __syncthreads();
    float g = __ldg(hp.b2 + d);
#pragma unroll 8
    for (int k = 0; k < kEmbDim; ++k) g = fmaf(hbar[k], __ldg(hp.w2t + k * kEmbDim + d), g);
    if (hp.global_out) hp.global_out[static_cast<size_t>(b) * kEmbDim + d] = g;
    gvec[d] = g;
    __syncthreads();
    if (hp.yn != nullptr && d < 32) {
        float4 x = reinterpret_cast<const float4*>(gvec)[lane];
        float ss = x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
        const float nrm = sqrtf(ss);
        x = make_float4(x.x / nrm, x.y / nrm, x.z / nrm, x.w / nrm);
        score_image_warp(x, hp.yn, hp.L, hp.NPP, b, hp.score, lane);
    }
}

// Patch path.  The two halves of the CTA (128 threads = 128 output channels each) take the even / odd patches; kHeadPB
// patches per half are in flight per iteration (their 128-float hidden rows are staged in shared memory by one coalesced
// load, every projector weight is read once per iteration and reused for all of them), so an image needs
// ceil(225 / 16) = 15 load -> compute -> normalise rounds instead of 113.  Arithmetic (FMA order over k, the order in
// which patches are added into the mean, the reduction trees) is exactly that of the one-patch-at-a-time formulation.
constexpr int kHeadPB = 8;
constexpr int kHeadSmemBytes = (kEmbDim * kEmbDim + 2 * (2 * kHeadPB) * kEmbDim + 2 * kHeadPB * 4 + 2 * kEmbDim) * 4;

__global__ void __launch_bounds__(256) head_kernel(const HeadParams hp) {
    extern __shared__ float hsm[];
    float* w2t = hsm;                                   // 128 * 128
    float* hrow = w2t + kEmbDim * kEmbDim;              // [2 * PB][128] hidden rows of the patches in flight
    float* prow = hrow + 2 * kHeadPB * kEmbDim;         // [2 * PB][128] projected (normalised) rows, for the heat-maps
    float* red = prow + 2 * kHeadPB * kEmbDim;          // [2 * PB][4] partial sums of squares
    float* gsum = red + 2 * kHeadPB * 4;                // 2 * 128
    const int b = blockIdx.x;
    const int tid = threadIdx.x;
    const int half = tid >> 7;
    const int d = tid & 127;
    const int lane = tid & 31;
    const int wq = (tid >> 5) & 3;  // warp within the half
    for (int i = tid; i < kEmbDim * kEmbDim; i += 256) w2t[i] = __ldg(hp.w2t + i);
    const float bias = __ldg(hp.b2 + d);
    float gacc = 0.0f;
    asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");   // programmatic dependent launch, see ptx.cuh
    asm volatile("griddepcontrol.wait;\n" ::: "memory");                 // weights above are constants; hid is not
    const bool need_norm = (hp.patch_out && hp.normalize_patch) || hp.heat_out;
    const float* hid_b = hp.hid + static_cast<size_t>(b) * hp.P * kEmbDim;
    for (int p0 = 0; p0 < hp.P; p0 += 2 * kHeadPB) {
        __syncthreads();   // the previous round's hrow / prow / red have been consumed (and w2t is complete)
#pragma unroll
        for (int i = 0; i < kHeadPB; ++i) {   // slot s = 2 j + half holds patch p0 + s
            const int idx = tid + i * 256;
            const int pidx = p0 + (idx >> 7);
            hrow[idx] = pidx < hp.P ? __ldg(hid_b + static_cast<size_t>(pidx) * kEmbDim + (idx & 127)) : 0.0f;
        }
        __syncthreads();
        float acc[kHeadPB];
#pragma unroll
        for (int jj = 0; jj < kHeadPB; ++jj) acc[jj] = bias;
        const float* hr = hrow + half * kEmbDim;
#pragma unroll 4
        for (int k = 0; k < kEmbDim; ++k) {
            const float w = w2t[k * kEmbDim + d];
#pragma unroll
            for (int jj = 0; jj < kHeadPB; ++jj) acc[jj] = fmaf(hr[2 * jj * kEmbDim + k], w, acc[jj]);
        }
#pragma unroll
        for (int jj = 0; jj < kHeadPB; ++jj)
            if (p0 + 2 * jj + half < hp.P) gacc += acc[jj];
        float inv[kHeadPB];
#pragma unroll
        for (int jj = 0; jj < kHeadPB; ++jj) inv[jj] = 1.0f;
        if (need_norm) {
#pragma unroll
            for (int jj = 0; jj < kHeadPB; ++jj) {
                float ss = acc[jj] * acc[jj];
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
                if (lane == 0) red[(2 * jj + half) * 4 + wq] = ss;
            }
            __syncthreads();
#pragma unroll
            for (int jj = 0; jj < kHeadPB; ++jj) {
                const float* r4 = red + (2 * jj + half) * 4;
                const float tot = r4[0] + r4[1] + r4[2] + r4[3];
                inv[jj] = 1.0f / fmaxf(sqrtf(tot), 1e-12f);
            }
        }
        if (hp.patch_out) {
#pragma unroll
            for (int jj = 0; jj < kHeadPB; ++jj) {
                const int pidx = p0 + 2 * jj + half;
                if (pidx < hp.P)
                    hp.patch_out[(static_cast<size_t>(b) * hp.P + pidx) * kEmbDim + d] = hp.normalize_patch ? acc[jj] * inv[jj] : acc[jj];
            }
        }
        if (hp.heat_out) {
#pragma unroll
            for (int jj = 0; jj < kHeadPB; ++jj) prow[(2 * jj + half) * kEmbDim + d] = acc[jj] * inv[jj];
            __syncthreads();
            // (patch slot, label) pairs dealt over the 8 warps
            const int warp = tid >> 5;
            for (int pr = warp; pr < 2 * kHeadPB * hp.L; pr += 8) {
                const int s = pr / hp.L, l = pr - s * hp.L;
                const int pidx = p0 + s;
                const float4 pv = reinterpret_cast<const float4*>(prow + s * kEmbDim)[lane];
                const float4 tv = reinterpret_cast<const float4*>(hp.heat_t + static_cast<size_t>(l) * kEmbDim)[lane];
                float dsum = pv.x * tv.x + pv.y * tv.y + pv.z * tv.z + pv.w * tv.w;
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) dsum += __shfl_xor_sync(0xffffffffu, dsum, off);
                if (lane == 0 && pidx < hp.P) hp.heat_out[(static_cast<size_t>(b) * hp.P + pidx) * hp.L + l] = dsum;
            }
        }
    }
    gsum[tid] = gacc;
    __syncthreads();
    if (tid < kEmbDim) {
        const float g = (gsum[tid] + gsum[tid + kEmbDim]) / static_cast<float>(hp.P);
        gsum[tid] = g;
        if (hp.global_out) hp.global_out[static_cast<size_t>(b) * kEmbDim + tid] = g;
    }
    __syncthreads();
    if (hp.yn != nullptr && tid < 32) {
        float4 x = reinterpret_cast<const float4*>(gsum)[lane];
        float ss = x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
        const float nrm = sqrtf(ss);
        x = make_float4(x.x / nrm, x.y / nrm, x.z / nrm, x.w / nrm);
        score_image_warp(x, hp.yn, hp.L, hp.NPP, b, hp.score, lane);
    }
}

// ----------------------------------------------------------------------------------------------
// Heat-map smoothing: scipy.ndimage.gaussian_filter(sigma=(s,s), order=0, mode='reflect', truncate=4.0) applied to
// every [gh, gw] patch-similarity map (vlp/inference_engine.py:107-109), separable, rows axis first like scipy.
// heat / out are [B, gh, gw, L]; one CTA per (image, label) map; weights are computed on the host in double
// exactly as scipy's _gaussian_kernel1d does and passed by value.
// ----------------------------------------------------------------------------------------------
constexpr int kSmoothMaxRadius = 16;
constexpr int kSmoothMaxCells = 32 * 32;
struct SmoothParams {
    const float* heat;
    float* out;
    int B, gh, gw, L;
    int radius;
    float w[2 * kSmoothMaxRadius + 1];
};

// 'reflect' = half-sample symmetric: (d c b a | a b c d | d c b a)
__device__ __forceinline__ int reflect_index(int i, int n) {
    while (i < 0 || i >= n) i = (i < 0) ? (-i - 1) : (2 * n - 1 - i);
    return i;
}

__global__ void __launch_bounds__(256) heat_smooth_kernel(const SmoothParams p) {
    __shared__ float a[kSmoothMaxCells];
    __shared__ float t[kSmoothMaxCells];
    const int b = blockIdx.x / p.L, l = blockIdx.x - b * p.L;
    const int cells = p.gh * p.gw;
    const float* src = p.heat + static_cast<size_t>(b) * cells * p.L + l;
    for (int i = threadIdx.x; i < cells; i += blockDim.x) a[i] = src[static_cast<size_t>(i) * p.L];
    __syncthreads();
    for (int i = threadIdx.x; i < cells; i += blockDim.x) {  // axis 0 (rows)
        const int y = i / p.gw, x = i - y * p.gw;
        float acc = 0.0f;
        for (int k = -p.radius; k <= p.radius; ++k) acc = fmaf(p.w[k + p.radius], a[reflect_index(y + k, p.gh) * p.gw + x], acc);
        t[i] = acc;
    }
    __syncthreads();
    float* dst = p.out + static_cast<size_t>(b) * cells * p.L + l;
    for (int i = threadIdx.x; i < cells; i += blockDim.x) {  // axis 1 (columns)
        const int y = i / p.gw, x = i - y * p.gw;
        float acc = 0.0f;
        for (int k = -p.radius; k <= p.radius; ++k) acc = fmaf(p.w[k + p.radius], t[y * p.gw + reflect_index(x + k, p.gw)], acc);
        dst[static_cast<size_t>(i) * p.L] = acc;
    }
}

// ----------------------------------------------------------------------------------------------
// Patch-grid similarity maps -> image size (vlp/inference_engine.py:113-155 with interpolation="nearest", the
// reference's default): F.interpolate(mode="nearest") of the [gh, gw] grid to a `side_h x side_w` window (the centre
// crop expressed in the original image's pixels) placed at (`top`, `left`) of the [height, width] image, NaN everywhere
// else (F.pad(value=NaN); negative margins crop, as F.pad does).  PyTorch's nearest index: dst == src size -> identity,
// dst == 2 * src -> dst >> 1, else min(int(floorf(dst * (float(src) / dst_size))), src - 1) - reproduced exactly.
// heat [B, gh, gw, L] (channel-last, as bv_forward writes it) -> out [B, L, height, width]; one thread per output pixel.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ int nearest_src_index(int dst, int src_size, int dst_size, float scale) {
    if (dst_size == src_size) return dst;
    if (dst_size == 2 * src_size) return dst >> 1;
    return min(static_cast<int>(floorf(static_cast<float>(dst) * scale)), src_size - 1);
}

__global__ void __launch_bounds__(256) heat_to_image_kernel(const float* __restrict__ heat, float* __restrict__ out, int gh,
                                                           int gw, int L, int height, int width, int side_h, int side_w,
                                                           int top, int left) {
    const int map = blockIdx.y;                 // b * L + l
    const int b = map / L, l = map - b * L;
    const float scale_h = static_cast<float>(gh) / static_cast<float>(side_h);
    const float scale_w = static_cast<float>(gw) / static_cast<float>(side_w);
    const float* src = heat + static_cast<size_t>(b) * gh * gw * L + l;
    float* dst = out + static_cast<size_t>(map) * height * width;
    const int total = height * width;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int y = i / width, x = i - y * width;
        const int yy = y - top, xx = x - left;
        float v = __int_as_float(0x7fc00000);   // NaN
        if (yy >= 0 && yy < side_h && xx >= 0 && xx < side_w) {
            const int sy = nearest_src_index(yy, gh, side_h, scale_h);
            const int sx = nearest_src_index(xx, gw, side_w, scale_w);
            v = __ldg(src + (static_cast<size_t>(sy) * gw + sx) * L);
        }
        dst[i] = v;
    }
}

}  // namespace bv
