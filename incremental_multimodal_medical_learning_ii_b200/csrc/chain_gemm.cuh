// Two chained GEMMs in one kernel: the tail of one Bottleneck and the head of the next.
//
//   y[m, :]  = relu( conv3(t2)[m, :] + bias3 (+ downsample(x)[m, :] + bias_ds | + residual[m, :]) )   -> out1 [M, N1] bf16
//   t1'[m, :] = relu( W1' . y[m, :] + bias1' )                                                        -> out2 [M, N2] bf16
//
// The next block's conv1 is a 1x1 convolution over exactly the pixels the current block's conv3 has just produced,
// so the 128-pixel x N1 tile that the epilogue has rounded to bf16 and laid out (128B-swizzled, K-major) for its TMA
// store is ALSO a valid tcgen05 A operand.  The second GEMM reads it from shared memory; the block output is written
// to HBM once (it is still needed as the next residual) but is never read back for conv1.  In layer1/layer2, where
// every kernel sits on the HBM roofline, that removes a quarter of all bytes moved.
//
// Replaces (reference): Bottleneck.forward's `conv3 -> bn3 -> (+identity | downsample) -> relu` followed by the next
// Bottleneck's `conv1 -> bn1 -> relu` (torchvision resnet.py Bottleneck.forward, driven by
// health_multimodal/image/model/resnet.py:38-42).
//
// N1 is processed in chunks of 128 columns.  Per chunk q (global over the CTA's tiles):
//   G1(q): D1[q&1] (128 TMEM columns)  = A1 tile x W3[chunk]^T           operands through the smem ring
//   E1(q): 8 epilogue warps: D1 -> +bias (+residual, TMA-prefetched into the staging tile) -> ReLU -> bf16, in place
//          in two 128x64 staging sub-tiles; DMA warp TMA-stores them to out1
//   G2(q): D2 (N2 TMEM columns, accumulated over the chunks of a tile) += staging sub-tiles x W1'[:, chunk]^T
//   E2   : after the tile's last chunk: D2 -> +bias -> ReLU -> bf16 -> second staging area -> TMA store to out2
// The MMA warp issues G1(q) BEFORE G2(q-1), so the tensor core always has the next chunk's first GEMM queued while
// the epilogue converts the previous one.
//
// Warp roles (640 threads, 1 CTA / SM, persistent): 0 TMA producer, 1 MMA issuer (+TMEM alloc), 2..17 epilogue math
// (the epilogue is a chain of latencies - TMEM load, smem read-modify-write, barrier - so it is spread over 16 warps
// of 16 columns each; with 8 warps the no-residual block was epilogue-bound), 18 DMA for out1 (residual prefetch +
// stores), 19 DMA for out2.
#pragma once
#include "conv_gemm.cuh"

namespace bv {

constexpr int kChainBN1 = 128;                  // columns of N1 per chunk
constexpr int kChainEpiWarps = 16;              // four per TMEM lane quarter, 16 of a sub-tile's 64 columns each
constexpr int kChainThreads = (2 + kChainEpiWarps + 2) * 32;
constexpr int kChainStageBytes = 32 * 1024;     // [A 16 KB | B 16 KB]; a G2 stage holds only weights (N2 x 128 B)
constexpr int kChainDma1Warp = 2 + kChainEpiWarps;
constexpr int kChainDma2Warp = 3 + kChainEpiWarps;

struct ChainParams {
    CUtensorMap tmA[2];   // GEMM1 A operands (segment 0: conv3 input, segment 1: downsample input)
    CUtensorMap tmB1[2];  // GEMM1 weights [N1, K], box 64 x 128
    CUtensorMap tmB2;     // GEMM2 weights [N2, N1], box 64 x N2
    CUtensorMap tmRes;    // residual [M, N1], box 64 x 128
    CUtensorMap tmOut1;   // block output [M, N1], box 64 x 128
    CUtensorMap tmOut2;   // next conv1 output [M, N2], box 64 x 128
    ConvSeg seg[2];
    int nseg;
    int Ho, Wo;
    int M, N1;
    int num_m_blocks;
    // biases by value (constant bank; see ConvGemmParams::bias_c): the first GEMM's summed segment biases, the second GEMM's
    float4 bias1_c[1024 / 4];
    float4 bias2_c[256 / 4];
    int has_res;
    int l2_prefetch;      // producer / DMA pull the next tile's A and residual into L2 ahead of the smem pipeline
};

template <int N2, int STAGES, int NB1>
struct ChainCfg {
    static_assert(N2 == 64 || N2 == 128 || N2 == 256, "second GEMM width");
    static constexpr int kNB2 = N2 / kChunkCols;
    static constexpr int kD2Bufs = (N2 <= 128) ? 2 : 1;
    static constexpr int kNumBars = 2 * STAGES + 8 + 3 * NB1 + 2 * kNB2;
    static constexpr int kSmemBytes =
        STAGES * kChainStageBytes + NB1 * kStagingBytes + kNB2 * kStagingBytes + kNumBars * 8 + 16;
    static_assert(kSmemBytes <= 232448, "exceeds the 227 KB of shared memory a CTA may use");
};

// bias (+ second bias of the fused downsample) + optional in-place residual + ReLU + bf16 pack of
// 16 accumulator columns (16-byte groups 2 cg and 2 cg + 1 of the 64-column sub-tile) of one row
__device__ __forceinline__ void chain_convert_row16(const uint32_t (&v)[16], const float4 (&bq)[4], bool has_res,
                                                    uint8_t* row_ptr, int cg, int r_in_tile) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        float f[8];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const float4 bb = bq[2 * j + q];
            f[4 * q + 0] = __uint_as_float(v[8 * j + 4 * q + 0]) + bb.x;
            f[4 * q + 1] = __uint_as_float(v[8 * j + 4 * q + 1]) + bb.y;
            f[4 * q + 2] = __uint_as_float(v[8 * j + 4 * q + 2]) + bb.z;
            f[4 * q + 3] = __uint_as_float(v[8 * j + 4 * q + 3]) + bb.w;
        }
        const int jj = cg * 2 + j;
        uint4* sp = reinterpret_cast<uint4*>(row_ptr + ((jj ^ (r_in_tile & 7)) << 4));
        if (has_res) {
            const uint4 rv = *sp;
            const uint32_t w[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                f[2 * e + 0] += __uint_as_float(w[e] << 16);
                f[2 * e + 1] += __uint_as_float(w[e] & 0xFFFF0000u);
            }
        }
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const __nv_bfloat162 h2 = __floats2bfloat162_rn(fmaxf(f[2 * e], 0.0f), fmaxf(f[2 * e + 1], 0.0f));
            w[e] = *reinterpret_cast<const uint32_t*>(&h2);
        }
        *sp = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

__device__ __forceinline__ void chain_tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}

template <int N2, int STAGES, int NB1>
__global__ void __launch_bounds__(kChainThreads, 1) chain_gemm_kernel(const __grid_constant__ ChainParams p) {
    using Cfg = ChainCfg<N2, STAGES, NB1>;
    constexpr int kNB2 = Cfg::kNB2;
    constexpr int kD2Bufs = Cfg::kD2Bufs;

    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0u) __trap();
    uint8_t* ring = smem;
    uint8_t* stg1 = ring + STAGES * kChainStageBytes;
    uint8_t* stg2 = stg1 + NB1 * kStagingBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(stg2 + kNB2 * kStagingBytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + STAGES;
    uint64_t* d1_full = bars + 2 * STAGES;
    uint64_t* d1_empty = d1_full + 2;
    uint64_t* d2_full = d1_full + 4;
    uint64_t* d2_empty = d1_full + 6;
    uint64_t* buf_ready = d1_full + 8;            // staging tile free (+ residual landed)      DMA -> math
    uint64_t* buf_written = buf_ready + NB1;      // staging tile holds finished bf16 outputs   math -> DMA, MMA
    uint64_t* buf_consumed = buf_written + NB1;   // second GEMM has finished reading the tile  MMA -> DMA
    uint64_t* buf2_ready = buf_consumed + NB1;
    uint64_t* buf2_written = buf2_ready + kNB2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + Cfg::kNumBars);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int C = p.N1 / kChainBN1;  // chunks per tile
    const int my_tiles = (p.num_m_blocks - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                         static_cast<int>(gridDim.x);
    const int Q = my_tiles * C;      // chunks this CTA processes
    const int total_kb = p.seg[0].kblocks + (p.nseg > 1 ? p.seg[1].kblocks : 0);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tmA[0]);
        tma_prefetch_desc(&p.tmB1[0]);
        if (p.nseg > 1) {
            tma_prefetch_desc(&p.tmA[1]);
            tma_prefetch_desc(&p.tmB1[1]);
        }
        tma_prefetch_desc(&p.tmB2);
        tma_prefetch_desc(&p.tmOut1);
        tma_prefetch_desc(&p.tmOut2);
        if (p.has_res) tma_prefetch_desc(&p.tmRes);
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&d1_full[i], 1);
            mbar_init(&d1_empty[i], kChainEpiWarps);
            mbar_init(&d2_full[i], 1);
            mbar_init(&d2_empty[i], kChainEpiWarps);
        }
        for (int i = 0; i < NB1; ++i) {
            mbar_init(&buf_ready[i], 1);
            mbar_init(&buf_written[i], kChainEpiWarps);
            mbar_init(&buf_consumed[i], 1);
        }
        for (int i = 0; i < kNB2; ++i) {
            mbar_init(&buf2_ready[i], 1);
            mbar_init(&buf2_written[i], kChainEpiWarps);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_ptr, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_launch_dependents();   // the next kernel's prologue may overlap this kernel's tail ...
    pdl_wait();                // ... and this kernel touches activations only once its predecessors have completed

    if (Q == 0) {
        // nothing to do (the host never launches more CTAs than tiles)
    } else if (warp == 0) {
        // ===================== TMA producer: ring stages in the order the MMA warp consumes them =====================
        int stage = 0;
        uint32_t phase = 0;
        const int hw = p.Ho * p.Wo;
        auto advance = [&]() {
            if (++stage == STAGES) {
                stage = 0;
                phase ^= 1u;
            }
        };
        auto load_g1 = [&](int q) {
            const int it = q / C, c = q - it * C;
            const int tile = static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x);
            const int m0 = tile * kBlockM;
            const int img = m0 / hw;
            const int rem = m0 - img * hw;
            const int op = rem / p.Wo;
            const int oq = rem - op * p.Wo;
            if (p.l2_prefetch == 1 && c == 0 && it + 1 < my_tiles && p.seg[0].mode == kSegTiled && elect_one()) {
                // next tile's conv3 input: start its HBM fetch now, one whole tile ahead of the smem ring
                const int m1 = m0 + static_cast<int>(gridDim.x) * kBlockM;
                for (int cb = 0; cb < p.seg[0].cblocks; ++cb) tma_prefetch_l2_2d(&p.tmA[0], cb * kBlockK, m1);
            }
            // the A tile is fetched once per chunk: keep it in L2 until the tile's last chunk has taken it
            const uint64_t a_policy = (c == C - 1) ? kEvictFirst : kEvictLast;
            for (int s = 0; s < p.nseg; ++s) {
                const ConvSeg sg = p.seg[s];
                int cb = 0, kofs = 0, tr = 0, ts = 0;
                for (int kb = 0; kb < sg.kblocks; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1u);
                    if (elect_one()) {
                        uint8_t* dst = ring + stage * kChainStageBytes;
                        mbar_arrive_expect_tx(&full_bar[stage], kChainStageBytes);
                        if (sg.mode == kSegTiled) {
                            tma_load_2d(&p.tmA[s], &full_bar[stage], dst, cb * kBlockK, m0, a_policy);
                        } else {
                            tma_load_im2col_4d(&p.tmA[s], &full_bar[stage], dst, cb * kBlockK,
                                               sg.lower + oq * sg.stride, sg.lower + op * sg.stride, img,
                                               static_cast<uint16_t>(ts), static_cast<uint16_t>(tr), a_policy);
                        }
                        tma_load_2d(&p.tmB1[s], &full_bar[stage], dst + kABytes, kofs, c * kChainBN1, kEvictLast);
                    }
                    __syncwarp();
                    kofs += kBlockK;
                    if (++cb == sg.cblocks) {
                        cb = 0;
                        if (++ts == sg.S) {
                            ts = 0;
                            ++tr;
                        }
                    }
                    advance();
                }
            }
        };
        auto load_g2 = [&](int q) {
            // both 64-wide k-blocks of the chunk's second-GEMM weights travel in ONE ring stage when they fit
            // (N2 <= 128: 2 x N2 x 128 B <= 32 KB); N2 = 256 takes one stage per k-block
            const int c = q % C;
            constexpr int kPerStage = (N2 <= 128) ? 2 : 1;
            for (int j = 0; j < 2; j += kPerStage) {
                mbar_wait(&empty_bar[stage], phase ^ 1u);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&full_bar[stage], kPerStage * N2 * 128);
                    for (int u = 0; u < kPerStage; ++u)
                        tma_load_2d(&p.tmB2, &full_bar[stage], ring + stage * kChainStageBytes + u * N2 * 128,
                                    (2 * c + j + u) * kBlockK, 0, kEvictLast);
                }
                __syncwarp();
                advance();
            }
        };
        for (int q = 0; q < Q; ++q) {
            load_g1(q);
            if (q >= 1) load_g2(q - 1);
        }
        load_g2(Q - 1);
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc1 = umma_idesc_bf16_f32(kBlockM, kChainBN1);
        constexpr uint32_t idesc2 = umma_idesc_bf16_f32(kBlockM, N2);
        const uint32_t ring_base = smem_u32(ring);
        const uint32_t stg_base = smem_u32(stg1);
        int stage = 0;
        uint32_t phase = 0;
        auto advance = [&]() {
            if (++stage == STAGES) {
                stage = 0;
                phase ^= 1u;
            }
        };
        auto g1 = [&](int q) {
            const int d = q & 1;
            mbar_wait(&d1_empty[d], ((q >> 1) & 1u) ^ 1u);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(d * kChainBN1);
            for (int kb = 0; kb < total_kb; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t adesc = umma_desc_k_sw128(ring_base + static_cast<uint32_t>(stage * kChainStageBytes));
                    const uint64_t bdesc =
                        umma_desc_k_sw128(ring_base + static_cast<uint32_t>(stage * kChainStageBytes + kABytes));
#pragma unroll
                    for (int k = 0; k < kBlockK / 16; ++k)
                        umma_bf16_ss(d_tmem, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k),
                                     idesc1, (kb != 0 || k != 0) ? 1u : 0u);
                    umma_commit(&empty_bar[stage]);
                    if (kb == total_kb - 1) umma_commit(&d1_full[d]);
                }
                __syncwarp();
                advance();
            }
        };
        auto g2 = [&](int q) {
            const int it = q / C, c = q - it * C;
            const int acc2 = (kD2Bufs == 2) ? (it & 1) : 0;
            const uint32_t ph2 = (kD2Bufs == 2) ? ((it >> 1) & 1u) : (it & 1u);
            if (c == 0) {
                mbar_wait(&d2_empty[acc2], ph2 ^ 1u);
                tc_fence_after();
            }
            const uint32_t d_tmem = tmem_base + 256u + static_cast<uint32_t>(acc2 * 128);
            constexpr int kPerStage = (N2 <= 128) ? 2 : 1;
            for (int j = 0; j < 2; ++j) {
                const int g = 2 * q + j;
                const int b = g % NB1;
                const bool new_stage = (j % kPerStage) == 0;
                const bool last_of_stage = (j % kPerStage) == kPerStage - 1;
                mbar_wait(&buf_written[b], (g / NB1) & 1u);
                if (new_stage) mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t adesc = umma_desc_k_sw128(stg_base + static_cast<uint32_t>(b * kStagingBytes));
                    const uint64_t bdesc = umma_desc_k_sw128(
                        ring_base + static_cast<uint32_t>(stage * kChainStageBytes + (j % kPerStage) * N2 * 128));
#pragma unroll
                    for (int k = 0; k < kBlockK / 16; ++k)
                        umma_bf16_ss(d_tmem, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k),
                                     idesc2, (c != 0 || j != 0 || k != 0) ? 1u : 0u);
                    if (last_of_stage) umma_commit(&empty_bar[stage]);
                    umma_commit(&buf_consumed[b]);
                    if (c == C - 1 && j == 1) umma_commit(&d2_full[acc2]);
                }
                __syncwarp();
                if (last_of_stage) advance();
            }
        };
        for (int q = 0; q < Q; ++q) {
            g1(q);
            if (q >= 1) g2(q - 1);
        }
        g2(Q - 1);
    } else if (warp == kChainDma1Warp) {
        // ===================== DMA 1: residual prefetch + out1 stores =====================
        if (lane == 0) {
            const int total = 2 * Q;  // staging sub-tiles
            auto coords = [&](int g, int& row0, int& col0) {
                const int q = g >> 1;
                const int it = q / C, c = q - it * C;
                row0 = (static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x)) * kBlockM;
                col0 = c * kChainBN1 + (g & 1) * kChunkCols;
            };
            auto prepare = [&](int g) {
                const int b = g % NB1;
                if (p.has_res) {
                    int row0, col0;
                    coords(g, row0, col0);
                    mbar_arrive_expect_tx(&buf_ready[b], kStagingBytes);
                    tma_load_2d(&p.tmRes, &buf_ready[b], stg1 + b * kStagingBytes, col0, row0, kEvictFirst);
                } else {
                    mbar_arrive(&buf_ready[b]);
                }
            };
            // L2 prefetch of the residual one tile beyond the smem staging horizon
            auto prefetch_res = [&](int g) {
                if (p.has_res && p.l2_prefetch && g < total) {
                    int row0, col0;
                    coords(g, row0, col0);
                    tma_prefetch_l2_2d(&p.tmRes, col0, row0);
                }
            };
            const int ahead = p.l2_prefetch > 1 ? p.l2_prefetch : 2 * C;  // default: sub-tiles per tile
            for (int g = 0; g < NB1 + ahead; ++g) {
                if (g < NB1) {
                    if (g < total) prepare(g);
                } else {
                    prefetch_res(g);
                }
            }
            constexpr int kLag = 2;
            for (int g = 0; g < total; ++g) {
                const int b = g % NB1;
                mbar_wait(&buf_written[b], (g / NB1) & 1u);
                int row0, col0;
                coords(g, row0, col0);
                tma_store_2d(&p.tmOut1, stg1 + b * kStagingBytes, col0, row0);
                tma_store_commit();
                if (g >= kLag && g - kLag + NB1 < total) {
                    const int gg = g - kLag;
                    tma_store_wait_read<kLag>();                              // store gg has finished reading smem
                    mbar_wait(&buf_consumed[gg % NB1], (gg / NB1) & 1u);      // and so has the second GEMM
                    prepare(gg + NB1);
                    prefetch_res(gg + NB1 + ahead);
                }
            }
            tma_store_wait_all<0>();
        }
    } else if (warp == kChainDma2Warp) {
        // ===================== DMA 2: out2 stores =====================
        if (lane == 0) {
            for (int j = 0; j < kNB2; ++j) mbar_arrive(&buf2_ready[j]);
            for (int it = 0; it < my_tiles; ++it) {
                const int row0 = (static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x)) * kBlockM;
                for (int j = 0; j < kNB2; ++j) {
                    mbar_wait(&buf2_written[j], it & 1u);
                    tma_store_2d(&p.tmOut2, stg2 + j * kStagingBytes, j * kChunkCols, row0);
                    tma_store_commit();
                }
                if (it + 1 < my_tiles) {
                    tma_store_wait_read<0>();
                    for (int j = 0; j < kNB2; ++j) mbar_arrive(&buf2_ready[j]);
                }
            }
            tma_store_wait_all<0>();
        }
    } else {
        // ===================== epilogue math (warps 2..17) =====================
        const int quarter = warp & 3;
        const int cg = (warp - 2) >> 2;      // 16-column group of every 64-column sub-tile
        const int r_in_tile = quarter * 32 + lane;
        const bool has_res = p.has_res != 0;
        const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
        // E2 of a tile is deferred until after E1 of the NEXT tile's first chunk: the second GEMM's last partial
        // product (issued only once E1 of the last chunk has been written) then completes behind useful work.
        auto epilogue2 = [&](int it) {
            const int acc2 = (kD2Bufs == 2) ? (it & 1) : 0;
            const uint32_t ph2 = (kD2Bufs == 2) ? ((it >> 1) & 1u) : (it & 1u);
            mbar_wait(&d2_full[acc2], ph2);
            tc_fence_after();
#pragma unroll 1
            for (int sub = 0; sub < kNB2; ++sub) {
                uint32_t v[16];
                chain_tmem_ld_32x16(lane_base + 256u + static_cast<uint32_t>(acc2 * 128 + sub * kChunkCols + cg * 16), v);
                mbar_wait(&buf2_ready[sub], it & 1u);
                tmem_ld_wait();
                const int b4 = (sub * kChunkCols + cg * 16) >> 2;
                const float4 bq[4] = {p.bias2_c[b4], p.bias2_c[b4 + 1], p.bias2_c[b4 + 2], p.bias2_c[b4 + 3]};
                chain_convert_row16(v, bq, false, stg2 + sub * kStagingBytes + r_in_tile * 128, cg, r_in_tile);
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&buf2_written[sub]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&d2_empty[acc2]);
        };
        for (int q = 0; q < Q; ++q) {
            const int it = q / C, c = q - it * C;
            const int d = q & 1;
            mbar_wait(&d1_full[d], (q >> 1) & 1u);
            tc_fence_after();
#pragma unroll 1
            for (int sub = 0; sub < 2; ++sub) {
                const int g = 2 * q + sub;
                const int b = g % NB1;
                uint32_t v[16];
                chain_tmem_ld_32x16(lane_base + static_cast<uint32_t>(d * kChainBN1 + sub * kChunkCols + cg * 16), v);
                mbar_wait(&buf_ready[b], (g / NB1) & 1u);
                tmem_ld_wait();
                const int b4 = (c * kChainBN1 + sub * kChunkCols + cg * 16) >> 2;
                const float4 bq[4] = {p.bias1_c[b4], p.bias1_c[b4 + 1], p.bias1_c[b4 + 2], p.bias1_c[b4 + 3]};
                chain_convert_row16(v, bq, has_res, stg1 + b * kStagingBytes + r_in_tile * 128, cg, r_in_tile);
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&buf_written[b]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&d1_empty[d]);
            if (c == 0 && it > 0) epilogue2(it - 1);
        }
        epilogue2(my_tiles - 1);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace bv
