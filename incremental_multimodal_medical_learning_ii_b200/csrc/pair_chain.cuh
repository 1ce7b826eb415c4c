// Chained Bottleneck tail + next Bottleneck head on CTA PAIRS (tcgen05 cta_group::2), for the deep layers:
//
//   y[m, :]   = relu( conv3(t2)[m, :] + bias3 + identity[m, :] )      -> out1 [M, N1] bf16   (N1 = 512 / 1024)
//   t1'[m, :] = relu( W1' . y[m, :] + bias1' )                        -> out2 [M, N2] bf16   (N2 = 128 / 256)
//
// Same data flow as chain_gemm.cuh (the y tile the epilogue lays out for its TMA store is also the A operand of the second
// GEMM, so y is written to HBM once and never read back for conv1: -0.94 GB per layer3 block), rebuilt around what made
// the single-CTA form lose in layer3 (it moved 1.5 MB of operands from L2 per 128-row tile and was bound by that stream):
//
//   * the conv3 input tile (128 rows x K1 <= 256: 64 KB) is loaded ONCE per tile and stays resident while the N1 / 128
//     column chunks stream their weights against it (the single-CTA kernel re-fetched it for every chunk);
//   * the pair splits every weight stage: each CTA holds 64 of a chunk's 128 conv3 rows and N2 / 2 of the conv1 rows, so
//     the per-SM weight stream halves;  per tile and SM 0.58 MB of operands instead of 1.5 MB;
//   * t1' leaves through direct 32-byte stores (10 % of the tile's bytes), so the whole staging ring carries y sub-tiles
//     with uniform barrier phases and is deep enough (5-6 x 16 KB) to keep the identity stream in flight.
//
// Replaces (reference): torchvision Bottleneck.forward `conv3 -> bn3 -> += identity -> relu` followed by the next
// Bottleneck's `conv1 -> bn1 -> relu` (health_multimodal/image/model/resnet.py:40-41, layer2 / layer3).
//
// Per chunk q (global over the CTA's tiles; c = q % C, C = N1 / 128), pair-wide:
//   G1(q): D1[q & 1] (128 TMEM columns) = A tile x W3[chunk]^T          leader issues, cta_group::2, M = 256
//   E1(q): 16 epilogue warps per CTA: D1 -> + bias + identity (TMA-prefetched into the staging sub-tile) -> ReLU -> bf16
//          in place; the DMA warp TMA-stores the sub-tile to out1
//   G2(q): D2 (N2 columns, accumulated over the tile's chunks) += y sub-tiles x W1'[:, chunk]^T
//   E2   : after the tile's last chunk (deferred behind E1 of the next tile's first chunk): D2 -> + bias -> ReLU -> out2
// Protocol (pair_gemm.cuh): loads of both CTAs credit the LEADER's full barriers, the leader's MMA thread issues, commits
// are multicast to both CTAs, "drained" / "written" arrive on the leader's barriers from the epilogue warps of both.
#pragma once
#include "chain_gemm.cuh"
#include "pair_gemm.cuh"

namespace bv {

constexpr int kPcEpiWarps = 16;
constexpr int kPcThreads = (2 + kPcEpiWarps + 1) * 32;   // TMA producer, MMA issuer, 16 epilogue warps, epilogue DMA
constexpr int kPcDmaWarp = 2 + kPcEpiWarps;
constexpr int kPcStageBytes = 16 * 1024;

struct PairChainParams {
    CUtensorMap tmA;     // conv3 input [M, K1], box 64 x 128
    CUtensorMap tmB1;    // conv3 weights [N1, K1], box 64 x 64 (this CTA's half of a 128-column chunk)
    CUtensorMap tmB2;    // next conv1 weights [N2, N1], box 64 x N2/2
    CUtensorMap tmRes;   // identity [M, N1], box 64 x 128
    CUtensorMap tmOut1;  // block output [M, N1], box 64 x 128
    // biases by value (constant bank; see ConvGemmParams::bias_c)
    float4 bias1_c[1024 / 4];  // [N1]
    float4 bias2_c[256 / 4];   // [N2]
    __nv_bfloat16* out2; // [M, N2]
    int M, N1;
    int num_m_blocks;    // ceil(M / 128)
    int num_pair_tiles;  // ceil(num_m_blocks / 2)
    int res_prefetch;    // > 0: pull the identity sub-tile this many sub-tiles beyond the staging ring into L2 (experiment)
};

template <int N2, int KB1, int STAGES, int NSTG>
struct PairChainCfg {
    static_assert(N2 == 128 || N2 == 256, "second GEMM width");
    static_assert(KB1 == 2 || KB1 == 4, "conv3 K = 128 or 256");
    static constexpr int kG1Stages = KB1 / 2;                   // ring stages per chunk of G1: two 8 KB k-blocks each
    static constexpr int kB2Rows = N2 / 2;                      // conv1 weight rows per CTA
    static constexpr int kB2PerStage = kPcStageBytes / (kB2Rows * 128);   // k-blocks of W1' per stage: 1 (N2 = 256) or 2
    static constexpr int kG2Stages = 2 / kB2PerStage;
    static constexpr int kOffRing = KB1 * kABytes;
    static constexpr int kOffStg = kOffRing + STAGES * kPcStageBytes;
    static constexpr int kOffBars = kOffStg + NSTG * kStagingBytes;
    static constexpr int kNumBars = 2 * STAGES + 2 + 4 + 2 + 4 * NSTG;
    static constexpr int kSmemBytes = kOffBars + kNumBars * 8 + 16;
    static_assert(kSmemBytes <= 232448, "exceeds the 227 KB of shared memory a CTA may use");
};

template <int N2, int KB1, int STAGES, int NSTG>
__global__ void __launch_bounds__(kPcThreads, 1) pair_chain_kernel(const __grid_constant__ PairChainParams p) {
    using Cfg = PairChainCfg<N2, KB1, STAGES, NSTG>;
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0u) __trap();
    uint8_t* a_tile = smem;
    uint8_t* ring = smem + Cfg::kOffRing;
    uint8_t* stg = smem + Cfg::kOffStg;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kOffBars);
    uint64_t* full_bar = bars;                       // [STAGES] leader: the stage has landed in BOTH CTAs
    uint64_t* empty_bar = full_bar + STAGES;         // [STAGES] per CTA (multicast commit)
    uint64_t* a_full = empty_bar + STAGES;           // leader: the A tiles of both CTAs have landed
    uint64_t* a_free = a_full + 1;                   // per CTA (multicast commit after the tile's last G1)
    uint64_t* d1_full = a_free + 1;                  // [2] per CTA
    uint64_t* d1_empty = d1_full + 2;                // [2] leader, 32 epilogue warps of the pair
    uint64_t* d2_full = d1_empty + 2;                // per CTA
    uint64_t* d2_empty = d2_full + 1;                // leader, 32
    uint64_t* stg_ready = d2_empty + 1;              // [NSTG] per CTA: sub-tile buffer free (+ identity rows landed)   DMA -> math
    uint64_t* y_written = stg_ready + NSTG;          // [NSTG] leader, 32: y sub-tile written in BOTH CTAs              math -> MMA
    uint64_t* stg_local = y_written + NSTG;          // [NSTG] per CTA, 16: y sub-tile written here                     math -> DMA
    uint64_t* stg_consumed = stg_local + NSTG;       // [NSTG] per CTA (multicast commit): G2 has read the sub-tile      MMA -> DMA
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + Cfg::kNumBars);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1;
    const int num_pairs = gridDim.x >> 1;
    const int C = p.N1 / kChainBN1;                                              // chunks per tile
    const int T = (p.num_pair_tiles - pair + num_pairs - 1) / num_pairs;         // tiles of this pair
    const int Q = T * C;                                                         // chunks of this pair
    auto m_block = [&](int it) { return 2 * (pair + it * num_pairs) + static_cast<int>(rank); };

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tmA);
        tma_prefetch_desc(&p.tmB1);
        tma_prefetch_desc(&p.tmB2);
        tma_prefetch_desc(&p.tmRes);
        tma_prefetch_desc(&p.tmOut1);
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        mbar_init(a_full, 1);
        mbar_init(a_free, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&d1_full[i], 1);
            mbar_init(&d1_empty[i], 2 * kPcEpiWarps);
        }
        mbar_init(d2_full, 1);
        mbar_init(d2_empty, 2 * kPcEpiWarps);
        for (int i = 0; i < NSTG; ++i) {
            mbar_init(&stg_ready[i], 1);
            mbar_init(&y_written[i], 2 * kPcEpiWarps);
            mbar_init(&stg_local[i], kPcEpiWarps);
            mbar_init(&stg_consumed[i], 1);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc_pair(tmem_ptr, 512);
        tmem_relinquish_pair();
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_launch_dependents();
    pdl_wait();
    constexpr uint32_t kD2 = 256;   // TMEM column plan: D1[0] 0..127, D1[1] 128..255, D2 256..256+N2

    if (T <= 0) {
        // nothing to do (the host never launches more pairs than pair tiles)
    } else if (warp == 0) {
        // ===================== TMA producer (both CTAs): A tile per tile, weight stages in MMA order =====================
        int stage = 0;
        uint32_t phase = 0;
        auto advance = [&]() {
            if (++stage == STAGES) {
                stage = 0;
                phase ^= 1u;
            }
        };
        auto load_g1 = [&](int q) {
            const int it = q / C, c = q - it * C;
            if (c == 0) {
                // the tile's conv3 input: resident for all C chunks; free once the previous tile's last G1 has read it
                if (it > 0) mbar_wait(a_free, (it - 1) & 1u);
                if (elect_one()) {
                    if (rank == 0) mbar_arrive_expect_tx(a_full, 2u * KB1 * kABytes);
                    const int m0 = m_block(it) * kBlockM;
#pragma unroll
                    for (int kb = 0; kb < KB1; ++kb)
                        tma_load_2d_pair(&p.tmA, a_full, a_tile + kb * kABytes, kb * kBlockK, m0, kEvictFirst);
                }
                __syncwarp();
            }
            for (int s2 = 0; s2 < Cfg::kG1Stages; ++s2) {
                mbar_wait(&empty_bar[stage], phase ^ 1u);
                if (elect_one()) {
                    if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2u * kPcStageBytes);
                    uint8_t* dst = ring + stage * kPcStageBytes;
#pragma unroll
                    for (int h = 0; h < 2; ++h)   // this CTA's 64 of the chunk's 128 weight rows, two k-blocks per stage
                        tma_load_2d_pair(&p.tmB1, &full_bar[stage], dst + h * 8192, (2 * s2 + h) * kBlockK,
                                         c * kChainBN1 + static_cast<int>(rank) * 64, kEvictLast);
                }
                __syncwarp();
                advance();
            }
        };
        auto load_g2 = [&](int q) {
            const int c = q % C;
            for (int s2 = 0; s2 < Cfg::kG2Stages; ++s2) {
                mbar_wait(&empty_bar[stage], phase ^ 1u);
                if (elect_one()) {
                    if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2u * kPcStageBytes);
                    uint8_t* dst = ring + stage * kPcStageBytes;
#pragma unroll
                    for (int u = 0; u < Cfg::kB2PerStage; ++u)
                        tma_load_2d_pair(&p.tmB2, &full_bar[stage], dst + u * Cfg::kB2Rows * 128,
                                         (2 * c + s2 * Cfg::kB2PerStage + u) * kBlockK, static_cast<int>(rank) * Cfg::kB2Rows,
                                         kEvictLast);
                }
                __syncwarp();
                advance();
            }
        };
        // order = the MMA thread's: G1(q) before G2(q-1) inside a tile, G2(q-1) first at a tile boundary
        for (int q = 0; q < Q; ++q) {
            const bool boundary = (q % C) == 0;
            if (q >= 1 && boundary) load_g2(q - 1);
            load_g1(q);
            if (q >= 1 && !boundary) load_g2(q - 1);
        }
        load_g2(Q - 1);
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (rank == 0) {
            constexpr uint32_t idesc1 = umma_idesc_bf16_f32(2 * kBlockM, kChainBN1);
            constexpr uint32_t idesc2 = umma_idesc_bf16_f32(2 * kBlockM, N2);
            const uint32_t a_base = smem_u32(a_tile);
            const uint32_t ring_base = smem_u32(ring);
            const uint32_t stg_base = smem_u32(stg);
            int stage = 0;
            uint32_t phase = 0;
            auto advance = [&]() {
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1u;
                }
            };
            auto g1 = [&](int q) {
                const int it = q / C, c = q - it * C;
                const int d = q & 1;
                mbar_wait_cluster(&d1_empty[d], ((q >> 1) & 1u) ^ 1u);
                if (c == 0) mbar_wait_cluster(a_full, it & 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(d * kChainBN1);
                for (int s2 = 0; s2 < Cfg::kG1Stages; ++s2) {
                    mbar_wait_cluster(&full_bar[stage], phase);
                    tc_fence_after();
                    if (elect_one()) {
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int kb = 2 * s2 + h;
                            const uint64_t adesc = umma_desc_k_sw128(a_base + static_cast<uint32_t>(kb * kABytes));
                            const uint64_t bdesc = umma_desc_k_sw128(ring_base + static_cast<uint32_t>(stage * kPcStageBytes + h * 8192));
#pragma unroll
                            for (int k = 0; k < kBlockK / 16; ++k)
                                umma_bf16_ss_pair(d_tmem, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k),
                                                  idesc1, (kb != 0 || k != 0) ? 1u : 0u);
                        }
                        umma_commit_pair(&empty_bar[stage]);
                        if (s2 == Cfg::kG1Stages - 1) {
                            if (c == C - 1) umma_commit_pair(a_free);
                            umma_commit_pair(&d1_full[d]);
                        }
                    }
                    __syncwarp();
                    advance();
                }
            };
            auto g2 = [&](int q) {
                const int it = q / C, c = q - it * C;
                if (c == 0) {
                    mbar_wait_cluster(d2_empty, (it & 1u) ^ 1u);
                    tc_fence_after();
                }
                const uint32_t d_tmem = tmem_base + kD2;
#pragma unroll
                for (int j = 0; j < 2; ++j) {   // the chunk's two 64-wide y sub-tiles = two k-blocks of the second GEMM
                    const int g = 2 * q + j;
                    const int b = g % NSTG;
                    const int u = j % Cfg::kB2PerStage;
                    mbar_wait_cluster(&y_written[b], (g / NSTG) & 1u);
                    if (u == 0) mbar_wait_cluster(&full_bar[stage], phase);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t adesc = umma_desc_k_sw128(stg_base + static_cast<uint32_t>(b * kStagingBytes));
                        const uint64_t bdesc = umma_desc_k_sw128(
                            ring_base + static_cast<uint32_t>(stage * kPcStageBytes + u * Cfg::kB2Rows * 128));
#pragma unroll
                        for (int k = 0; k < kBlockK / 16; ++k)
                            umma_bf16_ss_pair(d_tmem, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k),
                                              idesc2, (c != 0 || j != 0 || k != 0) ? 1u : 0u);
                        if (u == Cfg::kB2PerStage - 1) umma_commit_pair(&empty_bar[stage]);
                        umma_commit_pair(&stg_consumed[b]);
                        if (c == C - 1 && j == 1) umma_commit_pair(d2_full);
                    }
                    __syncwarp();
                    if (u == Cfg::kB2PerStage - 1) advance();
                }
            };
            for (int q = 0; q < Q; ++q) {
                const bool boundary = (q % C) == 0;
                if (q >= 1 && boundary) g2(q - 1);   // do not queue the tile's last G2 behind the wait for the next A tile
                g1(q);
                if (q >= 1 && !boundary) g2(q - 1);
            }
            g2(Q - 1);
        }
    } else if (warp == kPcDmaWarp) {
        // ===================== epilogue DMA (per CTA): identity prefetch + out1 stores =====================
        if (lane == 0) {
            const int total = 2 * Q;
            auto coords = [&](int g, int& row0, int& col0) {
                const int q = g >> 1;
                const int it = q / C, c = q - it * C;
                row0 = m_block(it) * kBlockM;
                col0 = c * kChainBN1 + (g & 1) * kChunkCols;
            };
            auto prepare = [&](int g) {
                const int b = g % NSTG;
                int row0, col0;
                coords(g, row0, col0);
                mbar_arrive_expect_tx(&stg_ready[b], kStagingBytes);
                tma_load_2d(&p.tmRes, &stg_ready[b], stg + b * kStagingBytes, col0, row0, kEvictFirst);
                if (p.res_prefetch > 0 && g + p.res_prefetch < total) {
                    coords(g + p.res_prefetch, row0, col0);
                    tma_prefetch_l2_2d(&p.tmRes, col0, row0);
                }
            };
            for (int g = 0; g < NSTG && g < total; ++g) prepare(g);
            constexpr int kLag = 2;
            for (int g = 0; g < total; ++g) {
                const int b = g % NSTG;
                mbar_wait(&stg_local[b], (g / NSTG) & 1u);
                int row0, col0;
                coords(g, row0, col0);
                tma_store_2d(&p.tmOut1, stg + b * kStagingBytes, col0, row0);
                tma_store_commit();
                if (g >= kLag && g - kLag + NSTG < total) {
                    const int gg = g - kLag;
                    tma_store_wait_read<kLag>();                               // store gg has finished reading smem
                    mbar_wait(&stg_consumed[gg % NSTG], (gg / NSTG) & 1u);     // and so has the second GEMM
                    prepare(gg + NSTG);
                }
            }
            tma_store_wait_all<0>();
        }
    } else {
        // ===================== epilogue math (warps 2..17, both CTAs) =====================
        const int quarter = warp & 3;
        const int cg = (warp - 2) >> 2;      // 16-column group of every 64-column sub-tile
        const int r_in_tile = quarter * 32 + lane;
        const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
        auto epilogue2 = [&](int it) {
            mbar_wait(d2_full, it & 1u);
            tc_fence_after();
            const int row = m_block(it) * kBlockM + r_in_tile;
            __nv_bfloat16* orow = p.out2 + static_cast<size_t>(row) * N2 + cg * 16;
#pragma unroll 1
            for (int sub = 0; sub < N2 / kChunkCols; ++sub) {
                uint32_t v[16];
                chain_tmem_ld_32x16(lane_base + kD2 + static_cast<uint32_t>(sub * kChunkCols + cg * 16), v);
                tmem_ld_wait();
                const int b4 = (sub * kChunkCols + cg * 16) >> 2;
                uint32_t w[8];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 bb = p.bias2_c[b4 + j];
                    const __nv_bfloat162 h0 = __floats2bfloat162_rn(fmaxf(__uint_as_float(v[4 * j + 0]) + bb.x, 0.0f),
                                                                    fmaxf(__uint_as_float(v[4 * j + 1]) + bb.y, 0.0f));
                    const __nv_bfloat162 h1 = __floats2bfloat162_rn(fmaxf(__uint_as_float(v[4 * j + 2]) + bb.z, 0.0f),
                                                                    fmaxf(__uint_as_float(v[4 * j + 3]) + bb.w, 0.0f));
                    w[2 * j] = *reinterpret_cast<const uint32_t*>(&h0);
                    w[2 * j + 1] = *reinterpret_cast<const uint32_t*>(&h1);
                }
                if (row < p.M) {   // 32 contiguous bytes per thread = one full sector
                    uint4* op = reinterpret_cast<uint4*>(orow + sub * kChunkCols);
                    op[0] = make_uint4(w[0], w[1], w[2], w[3]);
                    op[1] = make_uint4(w[4], w[5], w[6], w[7]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(d2_empty);
        };
        for (int q = 0; q < Q; ++q) {
            const int it = q / C, c = q - it * C;
            const int d = q & 1;
            mbar_wait(&d1_full[d], (q >> 1) & 1u);
            tc_fence_after();
#pragma unroll 1
            for (int sub = 0; sub < 2; ++sub) {
                const int g = 2 * q + sub;
                const int b = g % NSTG;
                uint32_t v[16];
                chain_tmem_ld_32x16(lane_base + static_cast<uint32_t>(d * kChainBN1 + sub * kChunkCols + cg * 16), v);
                mbar_wait(&stg_ready[b], (g / NSTG) & 1u);
                tmem_ld_wait();
                const int b4 = (c * kChainBN1 + sub * kChunkCols + cg * 16) >> 2;
                const float4 bq[4] = {p.bias1_c[b4], p.bias1_c[b4 + 1], p.bias1_c[b4 + 2], p.bias1_c[b4 + 3]};
                chain_convert_row16(v, bq, true, stg + b * kStagingBytes + r_in_tile * 128, cg, r_in_tile);
                fence_proxy_async_smem();   // generic-proxy writes -> visible to the pair MMA and the TMA store
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive_leader(&y_written[b]);
                    mbar_arrive(&stg_local[b]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&d1_empty[d]);
            if (c == 0 && it > 0) epilogue2(it - 1);   // behind useful work: the tile's last G2 completes meanwhile
        }
        epilogue2(T - 1);
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, 512);
    }
}

}  // namespace bv
