// Implicit-GEMM convolution on tcgen05 tensor cores (sm_100a), NHWC bf16 activations, K-major packed weights.
//
//   D[m, n] = sum_seg sum_tap sum_c  A_seg[pixel(m) shifted by tap, c] * W[n, k(seg,tap,c)]      (fp32 in TMEM)
//   out[m, n] = act( D[m, n] + bias[n] (+ residual[m, n]) )                                       (bf16 or fp32)
//
// m enumerates OUTPUT pixels (n_img, p, q) row-major, so `out` is the NHWC output tensor viewed as [M, Cout].
// One CTA tile is 128 consecutive output pixels x BN output channels.  A tiles are fetched by TMA straight
// from the NHWC input: im2col-mode TMA walks 128 output pixels across image rows / images, applies the conv
// stride and zero-fills the padding halo, so neither an im2col matrix nor a halo copy ever exists in HBM.
// Up to two A "segments" accumulate into the same TMEM tile: segment 1 is how the Bottleneck's strided 1x1
// downsample branch is folded into conv3 (weights concatenated along K, biases summed), which removes the
// identity tensor's write + read.
//
// Replaces (reference): torch.nn.Conv2d + BatchNorm2d + ReLU (+ residual add) as executed by torchvision
// Bottleneck.forward under health_multimodal/image/model/resnet.py:34-42 and the projector's first conv,
// health_multimodal/image/model/modules.py:43-46.  BatchNorm (eval) is folded into W / bias on the host.
//
// Warp roles (192 threads, 1 CTA / SM, persistent over tiles):
//   warp 0     : TMA producer (one elected lane)           smem ring  full[]/empty[]
//   warp 1     : TMEM allocator + tcgen05.mma issuer        TMEM ring  tmem_full[]/tmem_empty[] (2 accumulators)
//   warps 2..5 : epilogue, TMEM -> registers -> bias/residual/ReLU -> global
#pragma once
#include "ptx.cuh"

namespace bv {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // 64 bf16 = one 128-byte swizzle span
constexpr int kGemmThreads = 192;
constexpr int kABytes = kBlockM * kBlockK * 2;

enum : int { kSegTiled = 0, kSegIm2col = 1 };

struct ConvSeg {
    int kblocks;  // taps * (Cin / 64)
    int cblocks;  // Cin / 64
    int S;        // filter width (tap = r * S + s)
    int mode;     // kSegTiled (1x1 stride 1: A is the [M, Cin] matrix itself) or kSegIm2col
    int stride;   // conv stride
    int lower;    // -padding (window origin of output pixel 0)
};

struct ConvGemmParams {
    CUtensorMap tmA[2];
    CUtensorMap tmB[2];  // per-segment weights [N, taps*Cin], K-major
    ConvSeg seg[2];
    int nseg;
    int Ho, Wo;  // output spatial size, to split m into (image, p, q)
    int M, N;
    int num_m_blocks, num_n_blocks;
    const float* bias[2];           // per-segment [N] fp32 (summed)
    const __nv_bfloat16* residual;  // [M, N] or nullptr
    void* out;                      // [M, N] bf16 (or fp32 when out_fp32)
    int relu;
    int out_fp32;
};

template <int BN>
struct ConvGemmCfg {
    static constexpr int kBBytes = BN * kBlockK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kStages = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
    static constexpr int kTmemCols = 2 * BN;
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BN>
__global__ void __launch_bounds__(kGemmThreads, 1) conv_gemm_kernel(const __grid_constant__ ConvGemmParams p) {
    using Cfg = ConvGemmCfg<BN>;
    constexpr int kStages = Cfg::kStages;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + kStages * kABytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + kStages;
    uint64_t* tmem_full = bars + 2 * kStages;
    uint64_t* tmem_empty = bars + 2 * kStages + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_tiles = p.num_m_blocks * p.num_n_blocks;
    const int total_kb = p.seg[0].kblocks + (p.nseg > 1 ? p.seg[1].kblocks : 0);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tmA[0]);
        if (p.nseg > 1) tma_prefetch_desc(&p.tmA[1]);
        tma_prefetch_desc(&p.tmB[0]);
        if (p.nseg > 1) tma_prefetch_desc(&p.tmB[1]);
        for (int i = 0; i < kStages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 4);  // one arrival per epilogue warp
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_ptr, Cfg::kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const int hw = p.Ho * p.Wo;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m_blk = tile / p.num_n_blocks;
                const int n_blk = tile - m_blk * p.num_n_blocks;
                const int m0 = m_blk * kBlockM;
                const int img = m0 / hw;
                const int rem = m0 - img * hw;
                const int op = rem / p.Wo;
                const int oq = rem - op * p.Wo;
                for (int s = 0; s < p.nseg; ++s) {
                    const ConvSeg sg = p.seg[s];
                    int tap = 0, cb = 0, kofs = 0;
                    for (int kb = 0; kb < sg.kblocks; ++kb) {
                        mbar_wait(&empty_bar[stage], phase ^ 1u);
                        mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
                        void* dst_a = smem_a + stage * kABytes;
                        if (sg.mode == kSegTiled) {
                            tma_load_2d(&p.tmA[s], &full_bar[stage], dst_a, cb * kBlockK, m0, kEvictNormal);
                        } else {
                            const int r = tap / sg.S;
                            const int ss = tap - r * sg.S;
                            tma_load_im2col_4d(&p.tmA[s], &full_bar[stage], dst_a, cb * kBlockK,
                                               sg.lower + oq * sg.stride, sg.lower + op * sg.stride, img,
                                               static_cast<uint16_t>(ss), static_cast<uint16_t>(r), kEvictNormal);
                        }
                        tma_load_2d(&p.tmB[s], &full_bar[stage], smem_b + stage * Cfg::kBBytes, kofs, n_blk * BN,
                                    kEvictLast);
                        kofs += kBlockK;
                        if (++cb == sg.cblocks) {
                            cb = 0;
                            ++tap;
                        }
                        if (++stage == kStages) {
                            stage = 0;
                            phase ^= 1u;
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16_f32(kBlockM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1u;
                mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
                for (int kb = 0; kb < total_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint64_t adesc = umma_desc_k_sw128(smem_u32(smem_a + stage * kABytes));
                    const uint64_t bdesc = umma_desc_k_sw128(smem_u32(smem_b + stage * Cfg::kBBytes));
#pragma unroll
                    for (int k = 0; k < kBlockK / 16; ++k) {
                        // +32 bytes per UMMA_K=16 step inside the 128-byte swizzle span (>>4 encoded => +2)
                        umma_bf16_ss(d_tmem, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k),
                                     idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[stage]);  // frees this smem stage once the MMAs have read it
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                umma_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
            }
        }
    } else {
        // ===================== epilogue (warps 2..5) =====================
        const int quarter = warp & 3;  // TMEM lane quarter this warp may access
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int m_blk = tile / p.num_n_blocks;
            const int n_blk = tile - m_blk * p.num_n_blocks;
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1u;
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const int row = m_blk * kBlockM + quarter * 32 + lane;
            const bool row_ok = row < p.M;
            const size_t row_off = static_cast<size_t>(row) * p.N + static_cast<size_t>(n_blk) * BN;
            const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                                   static_cast<uint32_t>(acc * BN);
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                uint32_t v[32];
                tmem_ld_32x32(t_row + static_cast<uint32_t>(c * 32), v);
                // residual loads overlap the TMEM read
                uint4 res[4];
                const bool has_res = (p.residual != nullptr) && row_ok;
                if (has_res) {
                    const uint4* rp = reinterpret_cast<const uint4*>(p.residual + row_off + c * 32);
#pragma unroll
                    for (int j = 0; j < 4; ++j) res[j] = __ldg(rp + j);
                }
                tmem_ld_wait();
                const float4* bp = reinterpret_cast<const float4*>(p.bias[0] + n_blk * BN + c * 32);
                const float4* bp2 =
                    (p.nseg > 1) ? reinterpret_cast<const float4*>(p.bias[1] + n_blk * BN + c * 32) : nullptr;
                float f[32];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float4 b = __ldg(bp + j);
                    if (bp2) {
                        const float4 b2 = __ldg(bp2 + j);
                        b.x += b2.x; b.y += b2.y; b.z += b2.z; b.w += b2.w;
                    }
                    f[4 * j + 0] = __uint_as_float(v[4 * j + 0]) + b.x;
                    f[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + b.y;
                    f[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + b.z;
                    f[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + b.w;
                }
                if (has_res) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint32_t w[4] = {res[j].x, res[j].y, res[j].z, res[j].w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            // bf16 -> fp32 is a 16-bit shift
                            f[8 * j + 2 * e + 0] += __uint_as_float(w[e] << 16);
                            f[8 * j + 2 * e + 1] += __uint_as_float(w[e] & 0xFFFF0000u);
                        }
                    }
                }
                if (p.relu) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.0f);
                }
                if (row_ok) {
                    if (p.out_fp32) {
                        float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + row_off + c * 32);
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            op[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
                    } else {
                        uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + row_off + c * 32);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            uint32_t w[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const __nv_bfloat162 h = __floats2bfloat162_rn(f[8 * j + 2 * e], f[8 * j + 2 * e + 1]);
                                w[e] = *reinterpret_cast<const uint32_t*>(&h);
                            }
                            op[j] = make_uint4(w[0], w[1], w[2], w[3]);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::kTmemCols);
    }
}

}  // namespace bv
