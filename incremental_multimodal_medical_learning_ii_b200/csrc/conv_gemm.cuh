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
// kSegWide (3x3, stride 1, pad 1): the nine taps of a tile are NOT nine loads.  The im2col bounding box is widened by
// the padding on both sides, so the GEMM rows enumerate (image, row, column -1 .. W) and consecutive rows are
// consecutive pixels of the zero-padded image row.  One 130-pixel load per filter ROW then serves its three
// horizontal taps: tap s is the same smem tile read through a descriptor whose start is shifted by s pixel rows
// (128 B each; the swizzle follows absolute smem address bits, so base_offset stays 0).  A-tile fill traffic drops 3x; the two extra columns per image
// row are computed and dropped in the epilogue.
//
// Replaces (reference): torch.nn.Conv2d + BatchNorm2d + ReLU (+ residual add) as executed by torchvision
// Bottleneck.forward under health_multimodal/image/model/resnet.py:34-42 and the projector's first conv,
// health_multimodal/image/model/modules.py:43-46.  BatchNorm (eval) is folded into W / bias on the host.
//
// Warp roles (352 threads, 1 CTA / SM, persistent over tiles):
//   warp 0     : TMA producer (one elected lane)           smem ring  full[]/empty[]
//   warp 1     : TMEM allocator + tcgen05.mma issuer        TMEM ring  tmem_full[]/tmem_empty[] (2 accumulators)
//   warps 2..17: epilogue math: TMEM -> registers -> bias/residual/ReLU -> bf16, IN PLACE in a 128B-swizzled
//                staging tile of 128 rows x 64 channels (the residual was TMA-loaded into that same tile);
//                four warps per TMEM lane quarter, each taking 16 of the 64 channels of a sub-tile (the epilogue is
//                a chain of latencies, so it is spread over many warps)
//   warp 18    : epilogue DMA (one lane): TMA-loads the residual sub-tile ahead of the math warps and TMA-stores
//                finished sub-tiles, so global traffic of the epilogue is full 128-byte rows, never per-thread rows
// (fp32 output, used only by the tiny projector conv, keeps a direct per-thread store path.)
#pragma once
#include "ptx.cuh"
#include "pair_gemm.cuh"

namespace bv {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // 64 bf16 = one 128-byte swizzle span
constexpr int kMaxEpiWarps = 16;                     // epilogue warps are a per-configuration choice: 8 or 16
constexpr int kGemmMaxThreads = (3 + kMaxEpiWarps) * 32;
constexpr int gemm_threads(int epi_warps) { return (3 + epi_warps) * 32; }
constexpr int kChunkCols = 64;                       // bf16 columns per staging sub-tile (128 bytes per row)
constexpr int kStagingBytes = kBlockM * kChunkCols * 2;  // 16 KB
constexpr int kABytes = kBlockM * kBlockK * 2;

enum : int { kSegTiled = 0, kSegIm2col = 1, kSegWide = 2 };
constexpr int kWideRows = kBlockM + 2;                 // pixels per wide A tile: 128 outputs + 2 for the horizontal taps
constexpr int kWideABytes = 17 * 1024;                 // 130 x 128 B rounded up to the 1024-byte swizzle period

struct ConvSeg {
    int kblocks;  // taps * (Cin / 64)
    int cblocks;  // Cin / 64
    int S;        // filter width (tap = r * S + s)
    int mode;     // kSegTiled (1x1 stride 1: A is the [M, Cin] matrix itself) or kSegIm2col
    int stride;   // conv stride
    int lower;    // -padding (window origin of output pixel 0)
};

constexpr int kMaxBiasN = 2048;   // widest layer (layer4 conv3)

struct ConvGemmParams {
    CUtensorMap tmA[2];
    CUtensorMap tmB[2];  // per-segment weights [N, taps*Cin], K-major
    CUtensorMap tmOut;   // bf16 output [M, N], box 64 x 128, 128B swizzle (unused for fp32 output)
    CUtensorMap tmRes;   // bf16 residual [M, N], same box (unused without residual)
    ConvSeg seg[2];
    int nseg;
    int Ho, Wo;  // output spatial size, to split m into (image, p, q)
    int M, N;    // GEMM rows (for kSegWide: rows of the WIDENED pixel space, Wo + 2 per image row) and columns
    int Wwide;   // Wo + 2 in kSegWide mode, else 0
    int num_m_blocks, num_n_blocks;
    const float* bias[2];           // per-segment [N] fp32 (summed)
    const __nv_bfloat16* residual;  // [M, N] or nullptr
    void* out;                      // [M, N] bf16 (or fp32 when out_fp32)
    int relu;
    int out_fp32;
    long long* dbg;  // optional [gridDim][4] cycle counters: MMA warp {total, wait tmem_empty, wait full}, producer {wait empty}
    int nacc;  // independent TMEM accumulators the K steps are dealt over (1 .. 256/BN); summed in the epilogue
    int res_prefetch;  // > 0: pull the residual sub-tile this many sub-tiles beyond the staging ring into L2 (experiment)
    // The summed per-segment biases BY VALUE (the kernel's constant bank; kernel parameters may be 32 KB since CUDA 12.1).
    // The staged epilogue reads them with indexed constant loads: broadcast loads through L1 / shared memory cost LSU
    // data-pipe wavefronts, and in the epilogues with an identity stream that pipe is the in-step bound.
    float4 bias_c[kMaxBiasN / 4];
};

// BN = output channels per tile; STAGES = depth of the A/B smem ring; NBUF = epilogue staging tiles.
// Memory-bound layers want bytes in flight (Little's law: ~2 us loaded latency x 44 GB/s per SM ~ 100 KB):
// without a residual that is a deep A ring, with a residual (4x the A bytes) it is many staging tiles.
// BRES = the whole weight panel of the tile column (<= kMaxResidentKB k-blocks) is loaded into smem ONCE per CTA and
// only A tiles stream through the ring: for the narrow layer1 convs the per-tile weight re-fetch is a third of
// the L2->SM traffic, and those layers are L2-bandwidth-bound.
constexpr int kMaxResidentKB = 9;

// MT = m-tiles (128 rows each) that share one B tile per ring stage.  MT = 2 halves the weight traffic per output row
// and doubles the MMA work behind every stage (layer2's 3x3 convolutions sit between the L2->SM path and the TMA
// latency with MT = 1: 576 KB of operand loads per 128x128 tile); it needs 2 x BN <= 256 TMEM columns per stage.
// EPI = epilogue warps (8: two per TMEM lane quarter, 32 columns of a sub-tile each; 16: four per quarter, 16 columns each).
// The epilogue is a chain of latencies, so memory-bound configurations gain from 16 warps (layer2/3 conv3 + identity:
// -5..-11 %); the compute-bound 256-wide long-K configuration loses 3-9 % to the extra resident threads and keeps 8.
// TR = "transposed" tile for 128-wide layers: the WEIGHTS are the A operand (M = 128 output channels) and the 256 pixels
// of the stage's two m-tiles the B operand (N = 256).  A 128x128x16 MMA is bound by the 64 B/cycle A-operand read
// (78 cycles), a 128x256x16 one runs at the math rate (128 cycles for twice the work): -18 % tensor time.  The
// accumulator is then channel-major (TMEM lane = channel, column = pixel); the epilogue transposes it on the way into
// the swizzled staging tiles with 2-byte stores (32 per thread and sub-tile: the 32 lanes of a store cover 64 contiguous
// bytes of one pixel row; two lanes share each bank word = one extra wavefront, far off the critical path).
// PAIR = the tile is computed by a CTA PAIR (cluster of 2, tcgen05 cta_group::2): one MMA of M = 256 spans the two SMs, each
// CTA owns its own 128-pixel m-tile (A loads, TMEM lanes, epilogue, stores are CTA-local) and holds only HALF of every
// weight stage (BN/2 rows).  Per SM the weight stream from L2 halves and a ring stage shrinks from 48 to 32 KB; the
// layer3/layer4 convolutions, whose tiles are bound by the L2 -> SM operand stream, are the target.  Protocol as in
// pair_gemm.cuh / l1_block.cuh: both producers credit the LEADER's full barrier, only the leader issues MMAs, commits are
// multicast to both CTAs, "accumulator drained" arrives on the leader's barrier from the epilogue warps of both.
template <int BN, int STAGES, int NBUF, bool BRES = false, bool WIDE = false, int MT = 1, int EPI = 16, bool TR = false,
          bool PAIR = false>
struct ConvGemmCfg {
    static_assert(EPI == 8 || EPI == 16, "epilogue warps");
    static_assert(!TR || (MT == 2 && BN == 128 && EPI == 16 && NBUF == 2), "transposed tiles: 128 channels x 2 m-tiles");
    static constexpr int kEpiWarps = EPI;
    static constexpr int kThreads = gemm_threads(EPI);
    static_assert(!WIDE || BRES, "the wide 3x3 mode keeps the weights resident");
    static_assert(MT == 1 || (MT == 2 && BN == 128 && !BRES && !WIDE), "two m-tiles per stage: BN = 128, streamed weights only");
    static_assert(!PAIR || (BN == 256 && MT == 1 && !BRES && !WIDE && !TR), "CTA pairs: 256-wide tiles, streamed weights");
    static constexpr int kBBytes = (PAIR ? BN / 2 : BN) * kBlockK * 2;
    static constexpr int kAStage = WIDE ? kWideABytes : MT * kABytes;
    static constexpr int kStageBytes = kAStage + (BRES ? 0 : kBBytes);
    static constexpr int kResidentBytes = BRES ? kMaxResidentKB * kBBytes : 0;
    static constexpr int kStages = STAGES;
    static constexpr int kBufs = NBUF;
    // Back-to-back tcgen05.mma into the SAME accumulator serialise on the accumulate latency (~120 cycles), which
    // is 4x the work of a 128x64x16 MMA.  The 512 TMEM columns are therefore always fully used: two accumulator
    // stages of 256 columns, each holding 256/BN partial accumulators that take the K steps round-robin.
    static constexpr int kMaxAcc = 256 / BN;
    static constexpr int kAccStageCols = 256;
    static constexpr int kTmemCols = 512;
    static constexpr int kNumBars = 2 * STAGES + 4 + 2 * NBUF + 1;
    static constexpr int kSmemBytes =
        kStages * kStageBytes + kResidentBytes + NBUF * kStagingBytes + 1024 /*align slack*/ + kNumBars * 8 + 16;
    static_assert(kSmemBytes <= 232448, "exceeds the 227 KB of shared memory a CTA may use");
};

__device__ __forceinline__ void tmem_ld_32x16b(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}

template <int BN, int STAGES, int NBUF, bool BRES = false, bool WIDE = false, int MT = 1, int EPI = 16, bool TR = false,
          bool PAIR = false>
__global__ void __launch_bounds__(gemm_threads(EPI), 1) conv_gemm_kernel(const __grid_constant__ ConvGemmParams p) {
    using Cfg = ConvGemmCfg<BN, STAGES, NBUF, BRES, WIDE, MT, EPI, TR, PAIR>;
    constexpr int kEpiWarps = EPI;
    constexpr int kDmaWarp = 2 + EPI;
    constexpr int kWarpCols = kChunkCols / (EPI / 4);   // columns of a 64-column sub-tile per epilogue warp
    constexpr int kAStage = Cfg::kAStage;
    constexpr int kStages = Cfg::kStages;
    constexpr int kBufs = NBUF;

    // 1024-byte alignment (the 128B-swizzle period) is requested from the toolchain rather than obtained by rounding
    // a generic pointer: an integer round-trip hides the shared address space from the compiler and every staging
    // access becomes a generic LD.E/ST.E instead of LDS/STS.
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0u) __trap();
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + kStages * kAStage;  // per-stage B tiles, or the resident weight panel when BRES
    uint8_t* staging = smem + kStages * Cfg::kStageBytes + Cfg::kResidentBytes;  // NBUF x 16 KB, 1024-byte aligned
    uint64_t* bars = reinterpret_cast<uint64_t*>(staging + kBufs * kStagingBytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + kStages;
    uint64_t* tmem_full = bars + 2 * kStages;
    uint64_t* tmem_empty = bars + 2 * kStages + 2;
    uint64_t* buf_ready = bars + 2 * kStages + 4;            // staging tile free (+ residual landed)  DMA -> math
    uint64_t* buf_written = bars + 2 * kStages + 4 + kBufs;  // staging tile holds finished outputs   math -> DMA
    uint64_t* bres_bar = bars + 2 * kStages + 4 + 2 * kBufs;  // resident weight panel landed (BRES only)
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + Cfg::kNumBars);
    constexpr int kChunks = BN / kChunkCols;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    // a "tile" below is MT consecutive m-blocks x one n-block; with PAIR it is the pair's two m-blocks (this CTA computes
    // m-block 2q + rank) x one n-block, and tiles are dealt to CTA pairs instead of CTAs
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    const int tile0 = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
    const int tile_stride = PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
    const int num_tiles = ((p.num_m_blocks + (PAIR ? 2 : MT) - 1) / (PAIR ? 2 : MT)) * p.num_n_blocks;
    auto tile_mn = [&](int tile, int& m_blk, int& n_blk) {
        const int mq = tile / p.num_n_blocks;
        n_blk = tile - mq * p.num_n_blocks;
        m_blk = PAIR ? 2 * mq + static_cast<int>(rank) : mq;
    };
    // ring stages consumed per tile (kSegWide: one per filter row and channel block) / resident weight k-blocks
    const int total_kb = WIDE ? 3 * p.seg[0].cblocks : p.seg[0].kblocks + (p.nseg > 1 ? p.seg[1].kblocks : 0);
    const int total_wkb = p.seg[0].kblocks + (p.nseg > 1 ? p.seg[1].kblocks : 0);

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
            mbar_init(&tmem_empty[i], (PAIR ? 2 : 1) * kEpiWarps);  // one arrival per epilogue warp (of both CTAs of a pair)
        }
        for (int i = 0; i < kBufs; ++i) {
            mbar_init(&buf_ready[i], 1);
            mbar_init(&buf_written[i], kEpiWarps);
        }
        mbar_init(bres_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        if constexpr (PAIR) {
            tmem_alloc_pair(tmem_ptr, Cfg::kTmemCols);
            tmem_relinquish_pair();
        } else {
            tmem_alloc(tmem_ptr, Cfg::kTmemCols);
            tmem_relinquish();
        }
    }
    tc_fence_before();
    if constexpr (PAIR) cluster_sync_all();   // the peer's barriers must be initialised before any remote arrival
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_launch_dependents();   // the next kernel's prologue may overlap this kernel's tail ...
    pdl_wait();                // ... and this kernel touches activations only once its predecessors have completed

    if (warp == 0) {
        // ===================== TMA producer =====================
        // The whole warp runs the loop (all values stay warp-uniform, so the TMA operands live in uniform
        // registers); one elected lane issues.  Running the loop under `if (lane == 0)` instead makes the
        // compiler wrap every UTMALDG / UTCHMMA in an ELECT + R2UR.BROADCAST + BRA.U.ANY waterfall.
        int stage = 0;
        uint32_t phase = 0;
        const int hw = p.Ho * p.Wo;
        long long t_empty = 0;
        if constexpr (BRES) {
            // every tile of this CTA uses the same weight panel (the host only picks BRES when N == BN)
            if (elect_one()) {
                mbar_arrive_expect_tx(bres_bar, static_cast<uint32_t>(total_wkb) * Cfg::kBBytes);
                int slot = 0;
                for (int s = 0; s < p.nseg; ++s)
                    for (int kb = 0; kb < p.seg[s].kblocks; ++kb, ++slot)
                        tma_load_2d(&p.tmB[s], bres_bar, smem_b + slot * Cfg::kBBytes, kb * kBlockK, 0, kEvictLast);
            }
            __syncwarp();
        }
        for (int tile = tile0; tile < num_tiles; tile += tile_stride) {
            int m_blk, n_blk;
            tile_mn(tile, m_blk, n_blk);
            const int m0 = m_blk * kBlockM;
            if constexpr (WIDE) {
                // rows enumerate (image, p, q') with q' = -1 .. Wo: window origin of row m0 is (q' - 1, p - 1)
                const int hww = p.Ho * p.Wwide;
                const int img = m0 / hww;
                const int rem = m0 - img * hww;
                const int op = rem / p.Wwide;
                const int oq = rem - op * p.Wwide;
                const int cblocks = p.seg[0].cblocks;
                for (int tr = 0; tr < 3; ++tr) {
                    for (int cb = 0; cb < cblocks; ++cb) {
                        const long long tw = tclock();
                        mbar_wait(&empty_bar[stage], phase ^ 1u);
                        t_empty += tclock() - tw;
                        if (elect_one()) {
                            mbar_arrive_expect_tx(&full_bar[stage], kWideRows * 128);
                            tma_load_im2col_4d(&p.tmA[0], &full_bar[stage], smem_a + stage * kAStage, cb * kBlockK,
                                               oq - 1, op - 1, img, 0, static_cast<uint16_t>(tr), kEvictNormal);
                        }
                        __syncwarp();
                        if (++stage == kStages) {
                            stage = 0;
                            phase ^= 1u;
                        }
                    }
                }
                continue;
            }
            int img[MT], op[MT], oq[MT];
#pragma unroll
            for (int u = 0; u < MT; ++u) {
                const int mu = (m_blk * MT + u) * kBlockM;
                img[u] = mu / hw;
                const int rem = mu - img[u] * hw;
                op[u] = rem / p.Wo;
                oq[u] = rem - op[u] * p.Wo;
            }
            for (int s = 0; s < p.nseg; ++s) {
                const ConvSeg sg = p.seg[s];
                int tap = 0, cb = 0, kofs = 0, tr = 0, ts = 0;
                for (int kb = 0; kb < sg.kblocks; ++kb) {
                    const long long tw = tclock();
                    mbar_wait(&empty_bar[stage], phase ^ 1u);
                    t_empty += tclock() - tw;
                    if (elect_one()) {
                        if constexpr (PAIR) {
                            // the leader's barrier collects the bytes of both CTAs' loads of this stage
                            void* dst_a = smem_a + stage * kAStage;
                            if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2u * Cfg::kStageBytes);
                            if (sg.mode == kSegTiled)
                                tma_load_2d_pair(&p.tmA[s], &full_bar[stage], dst_a, cb * kBlockK, m_blk * kBlockM, kEvictNormal);
                            else
                                tma_load_im2col_4d_pair(&p.tmA[s], &full_bar[stage], dst_a, cb * kBlockK,
                                                        sg.lower + oq[0] * sg.stride, sg.lower + op[0] * sg.stride, img[0],
                                                        static_cast<uint16_t>(ts), static_cast<uint16_t>(tr), kEvictNormal);
                            // this CTA's half of the weight stage: rows [rank * BN/2, (rank + 1) * BN/2) of the n-block
                            tma_load_2d_pair(&p.tmB[s], &full_bar[stage], smem_b + stage * Cfg::kBBytes, kofs,
                                             n_blk * BN + static_cast<int>(rank) * (BN / 2), kEvictLast);
                        } else {
                        mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
#pragma unroll
                        for (int u = 0; u < MT; ++u) {
                            void* dst_a = smem_a + stage * kAStage + u * kABytes;
                            if (sg.mode == kSegTiled) {
                                tma_load_2d(&p.tmA[s], &full_bar[stage], dst_a, cb * kBlockK, (m_blk * MT + u) * kBlockM,
                                            kEvictNormal);
                            } else {
                                tma_load_im2col_4d(&p.tmA[s], &full_bar[stage], dst_a, cb * kBlockK,
                                                   sg.lower + oq[u] * sg.stride, sg.lower + op[u] * sg.stride, img[u],
                                                   static_cast<uint16_t>(ts), static_cast<uint16_t>(tr), kEvictNormal);
                            }
                        }
                        if constexpr (!BRES)
                            tma_load_2d(&p.tmB[s], &full_bar[stage], smem_b + stage * Cfg::kBBytes, kofs, n_blk * BN,
                                        kEvictLast);
                        }
                    }
                    __syncwarp();
                    kofs += kBlockK;
                    if (++cb == sg.cblocks) {  // next filter tap (tr, ts)
                        cb = 0;
                        ++tap;
                        if (++ts == sg.S) {
                            ts = 0;
                            ++tr;
                        }
                    }
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
        if (kTimingBuild && p.dbg && lane == 0) p.dbg[blockIdx.x * 4 + 3] = t_empty;
    } else if (warp == 1) {
        // ===================== MMA issuer (warp-convergent loop, one elected lane issues) =====================
        long long t_acc = 0, t_full = 0;
        const long long t_begin = tclock();
        constexpr uint32_t idesc = umma_idesc_bf16_f32(kBlockM, BN);
        const uint32_t a_base = smem_u32(smem_a);
        const uint32_t b_base = smem_u32(smem_b);
        const int nacc = p.nacc;       // power of two <= 256 / BN
        const int amask = nacc - 1;
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        if constexpr (BRES) mbar_wait(bres_bar, 0);
        // (CTA pairs: only the leader issues; cluster-scope acquires because the barriers collect remote arrivals)
        for (int tile = tile0; tile < num_tiles && (!PAIR || rank == 0); tile += tile_stride, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1u;
            long long tw = tclock();
            if constexpr (PAIR) mbar_wait_cluster(&tmem_empty[acc], acc_phase ^ 1u);
            else mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
            t_acc += tclock() - tw;
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * Cfg::kAccStageCols);
            for (int kb = 0; kb < total_kb; ++kb) {
                tw = tclock();
                if constexpr (PAIR) mbar_wait_cluster(&full_bar[stage], phase);
                else mbar_wait(&full_bar[stage], phase);
                t_full += tclock() - tw;
                tc_fence_after();
                if constexpr (WIDE) {
                    if (elect_one()) {
                        const int cblocks = p.seg[0].cblocks;
                        const int tr = kb / cblocks, cb = kb - tr * cblocks;
#pragma unroll
                        for (int ts = 0; ts < 3; ++ts) {
                            // tap (tr, ts): same tile, start shifted by ts pixel rows (128 B each).  The 128B swizzle is
                            // a function of the absolute smem address bits, so a start that is not 1024-byte aligned
                            // needs no correction: base_offset stays 0 (measured: setting it to ts reads wrong chunks).
                            const uint64_t adesc =
                                umma_desc_k_sw128(a_base + static_cast<uint32_t>(stage * kAStage + ts * 128));
                            const uint64_t bdesc = umma_desc_k_sw128(
                                b_base + static_cast<uint32_t>(((tr * 3 + ts) * cblocks + cb) * Cfg::kBBytes));
#pragma unroll
                            for (int k = 0; k < kBlockK / 16; ++k)
                                umma_bf16_ss(d_tmem + static_cast<uint32_t>((k & amask) * BN),
                                             adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k),
                                             idesc, (kb != 0 || ts != 0 || k >= nacc) ? 1u : 0u);
                        }
                        umma_commit(&empty_bar[stage]);
                        if (kb == total_kb - 1) umma_commit(&tmem_full[acc]);
                    }
                } else if constexpr (TR) {
                    if (elect_one()) {
                        // D[channel, pixel] += W[128 x 16] * X[256 x 16]^T: the two m-tiles of the stage are one contiguous
                        // 256-row swizzled tile
                        const uint64_t wdesc = umma_desc_k_sw128(b_base + static_cast<uint32_t>(stage * Cfg::kBBytes));
                        const uint64_t xdesc = umma_desc_k_sw128(a_base + static_cast<uint32_t>(stage * kAStage));
#pragma unroll
                        for (int k = 0; k < kBlockK / 16; ++k)
                            umma_bf16_ss(d_tmem, wdesc + static_cast<uint64_t>(2 * k), xdesc + static_cast<uint64_t>(2 * k),
                                         umma_idesc_bf16_f32(128, 256), (kb != 0 || k != 0) ? 1u : 0u);
                        umma_commit(&empty_bar[stage]);
                        if (kb == total_kb - 1) umma_commit(&tmem_full[acc]);
                    }
                } else if constexpr (PAIR) {
                    if (elect_one()) {
                        // D[256 x BN] over the two SMs: own A tile x (own half | peer half) of the weight stage
                        const uint64_t adesc = umma_desc_k_sw128(a_base + static_cast<uint32_t>(stage * kAStage));
                        const uint64_t bdesc = umma_desc_k_sw128(b_base + static_cast<uint32_t>(stage * Cfg::kBBytes));
#pragma unroll
                        for (int k = 0; k < kBlockK / 16; ++k)
                            umma_bf16_ss_pair(d_tmem, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k),
                                              umma_idesc_bf16_f32(2 * kBlockM, BN), (kb != 0 || k != 0) ? 1u : 0u);
                        umma_commit_pair(&empty_bar[stage]);                          // frees the stage in BOTH CTAs
                        if (kb == total_kb - 1) umma_commit_pair(&tmem_full[acc]);    // both epilogues may start
                    }
                } else if (elect_one()) {
                    const uint64_t bdesc =
                        umma_desc_k_sw128(b_base + static_cast<uint32_t>((BRES ? kb : stage) * Cfg::kBBytes));
#pragma unroll
                    for (int u = 0; u < MT; ++u) {
                        const uint64_t adesc = umma_desc_k_sw128(a_base + static_cast<uint32_t>(stage * kAStage + u * kABytes));
#pragma unroll
                        for (int k = 0; k < kBlockK / 16; ++k) {
                            // K step ks = 4*kb + k goes to partial accumulator (ks & amask); its first visit overwrites.
                            // +32 bytes per UMMA_K=16 step inside the 128-byte swizzle span (>>4 encoded => +2).
                            // With MT = 2 (nacc == 1) m-tile u accumulates in columns [u * BN, (u + 1) * BN) of the stage.
                            umma_bf16_ss(d_tmem + static_cast<uint32_t>((MT == 1 ? (k & amask) : u) * BN),
                                         adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k), idesc,
                                         (kb != 0 || k >= nacc) ? 1u : 0u);
                        }
                    }
                    umma_commit(&empty_bar[stage]);  // frees this smem stage once the MMAs have read it
                    if (kb == total_kb - 1) umma_commit(&tmem_full[acc]);  // accumulators complete -> epilogue
                }
                __syncwarp();
                if (++stage == kStages) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
        }
        if (kTimingBuild && p.dbg && lane == 0) {
            p.dbg[blockIdx.x * 4 + 0] = tclock() - t_begin;
            p.dbg[blockIdx.x * 4 + 1] = t_acc;
            p.dbg[blockIdx.x * 4 + 2] = t_full;
        }
    } else if (warp == kDmaWarp) {
        // ===================== epilogue DMA (residual prefetch + output stores) =====================
        if (!p.out_fp32 && !WIDE) {
            const bool has_res = p.residual != nullptr;
            const int my_tiles = (num_tiles - tile0 + tile_stride - 1) / tile_stride;
            const int total = my_tiles * MT * kChunks;  // sub-tiles this CTA produces
            auto coords = [&](int g, int& row0, int& col0) {
                const int tile = tile0 + (g / (MT * kChunks)) * tile_stride;
                int m_blk, n_blk;
                tile_mn(tile, m_blk, n_blk);
                row0 = (m_blk * MT + (g / kChunks) % MT) * kBlockM;
                col0 = n_blk * BN + (g % kChunks) * kChunkCols;
            };
            // every lane runs the loop; lane 0 (always the same lane: bulk async-groups are per thread) issues
            auto prepare = [&](int g) {  // staging tile (g % NBUF) is free here: hand it to the math warps
                const int b = g % kBufs;
                if (has_res) {
                    int row0, col0;
                    coords(g, row0, col0);
                    mbar_arrive_expect_tx(&buf_ready[b], kStagingBytes);
                    tma_load_2d(&p.tmRes, &buf_ready[b], staging + b * kStagingBytes, col0, row0, kEvictFirst);
                    if (p.res_prefetch > 0 && g + p.res_prefetch < total) {   // short-distance L2 prefetch of the identity rows
                        coords(g + p.res_prefetch, row0, col0);
                        tma_prefetch_l2_2d(&p.tmRes, col0, row0);
                    }
                } else {
                    mbar_arrive(&buf_ready[b]);
                }
            };
            if (lane == 0) {
            for (int g = 0; g < kBufs && g < total; ++g) prepare(g);
            for (int g = 0; g < total; ++g) {
                const int b = g % kBufs;
                mbar_wait(&buf_written[b], (g / kBufs) & 1u);
                int row0, col0;
                coords(g, row0, col0);
                tma_store_2d(&p.tmOut, staging + b * kStagingBytes, col0, row0);
                tma_store_commit();
                // Re-arm a staging tile once its store has finished READING smem.  With more than two tiles the
                // re-arm lags one store behind, so this thread never waits on the store it has just issued.
                constexpr int kLag = (kBufs > 2) ? 1 : 0;
                if (g >= kLag && g - kLag + kBufs < total) {
                    tma_store_wait_read<kLag>();
                    prepare(g - kLag + kBufs);
                }
            }
            tma_store_wait_all<0>();  // all global writes complete before the CTA may exit
            }
        }
    } else if (!p.out_fp32 && !WIDE) {
        // ===================== epilogue math (warps 2..17), staged bf16 output =====================
        const int quarter = warp & 3;        // TMEM lane quarter this warp may access
        const int cg = (warp - 2) >> 2;      // which kWarpCols of the sub-tile's 64 channels this warp converts
        const int r_in_tile = quarter * 32 + lane;
        const bool has_res = p.residual != nullptr;
        int it = 0;
        int g = 0;
        for (int tile = tile0; tile < num_tiles; tile += tile_stride, ++it) {
            int m_blk, n_blk;
            tile_mn(tile, m_blk, n_blk);
            (void)m_blk;
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1u;
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                                   static_cast<uint32_t>(acc * Cfg::kAccStageCols);
            if constexpr (TR) {
                // lane = output channel (quarters 0,1 -> channel chunk 0, quarters 2,3 -> chunk 1), column = pixel; the
                // four warps of a quarter take 32 pixels each.  Both channel chunks of an m-tile are written at once.
                const int cch = quarter >> 1;
                const int ch_local = (quarter & 1) * 32 + lane;
                const int pxg = cg;
                float bias_v = __ldg(p.bias[0] + n_blk * BN + cch * kChunkCols + ch_local);
                const uint32_t grp = static_cast<uint32_t>(ch_local >> 3), sub = static_cast<uint32_t>(ch_local & 7) * 2u;
#pragma unroll 1
                for (int u = 0; u < MT; ++u, g += kChunks) {
                    const int b0 = g % kBufs, b1 = (g + 1) % kBufs;
                    mbar_wait(&buf_ready[b0], (g / kBufs) & 1u);
                    mbar_wait(&buf_ready[b1], ((g + 1) / kBufs) & 1u);
                    uint32_t v[32];
                    tmem_ld_32x32(t_row + static_cast<uint32_t>(u * BN + pxg * 32), v);
                    tmem_ld_wait();
                    uint8_t* tile = staging + (cch ? b1 : b0) * kStagingBytes + pxg * 32 * 128;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        float f = __uint_as_float(v[j]) + bias_v;
                        if (p.relu) f = fmaxf(f, 0.0f);
                        const __nv_bfloat16 h = __float2bfloat16_rn(f);
                        *reinterpret_cast<__nv_bfloat16*>(tile + j * 128 + ((grp ^ static_cast<uint32_t>(j & 7)) << 4) + sub) = h;
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive(&buf_written[b0]);
                        mbar_arrive(&buf_written[b1]);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[acc]);
                continue;
            }
#pragma unroll 1
            for (int cu = 0; cu < MT * kChunks; ++cu, ++g) {   // m-tile u of the stage, 64-column sub-tile c
                const int u = cu / kChunks, c = cu - u * kChunks;
                const int b = g % kBufs;
                uint8_t* row_ptr = staging + b * kStagingBytes + r_in_tile * 128;
                mbar_wait(&buf_ready[b], (g / kBufs) & 1u);
                {
                    uint32_t v[kWarpCols];
                    if constexpr (kWarpCols == 32) tmem_ld_32x32(t_row + static_cast<uint32_t>(u * BN + c * kChunkCols + cg * kWarpCols), v);
                    else tmem_ld_32x16b(t_row + static_cast<uint32_t>(u * BN + c * kChunkCols + cg * kWarpCols), v);
                    tmem_ld_wait();
                    for (int a = 1; a < p.nacc; ++a) {  // add the other partial accumulators
                        uint32_t w2[kWarpCols];
                        if constexpr (kWarpCols == 32) tmem_ld_32x32(t_row + static_cast<uint32_t>(a * BN + c * kChunkCols + cg * kWarpCols), w2);
                        else tmem_ld_32x16b(t_row + static_cast<uint32_t>(a * BN + c * kChunkCols + cg * kWarpCols), w2);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < kWarpCols; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(w2[j]));
                    }
                    const int col4 = (n_blk * BN + c * kChunkCols + cg * kWarpCols) >> 2;
#pragma unroll
                    for (int j = 0; j < kWarpCols / 8; ++j) {  // 16-byte group = 8 channels
                        float f[8];
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            const float4 bb = p.bias_c[col4 + 2 * j + q];   // bias[0] (+ bias[1]), summed on the host
                            f[4 * q + 0] = __uint_as_float(v[8 * j + 4 * q + 0]) + bb.x;
                            f[4 * q + 1] = __uint_as_float(v[8 * j + 4 * q + 1]) + bb.y;
                            f[4 * q + 2] = __uint_as_float(v[8 * j + 4 * q + 2]) + bb.z;
                            f[4 * q + 3] = __uint_as_float(v[8 * j + 4 * q + 3]) + bb.w;
                        }
                        // 128B swizzle: 16-byte group jj of row r lives at group position jj ^ (r & 7)
                        const int jj = cg * (kWarpCols / 8) + j;
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
                        if (p.relu) {
#pragma unroll
                            for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], 0.0f);
                        }
                        uint32_t w[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const __nv_bfloat162 h2 = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
                            w[e] = *reinterpret_cast<const uint32_t*>(&h2);
                        }
                        *sp = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                }
                fence_proxy_async_smem();  // generic-proxy writes -> visible to the TMA store (async proxy)
                __syncwarp();
                if (lane == 0) mbar_arrive(&buf_written[b]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (PAIR) mbar_arrive_leader(&tmem_empty[acc]);   // the leader's MMA thread reuses the accumulator
                else mbar_arrive(&tmem_empty[acc]);
            }
        }
    } else {
        // ===================== epilogue, direct fp32 stores: the warps of a lane quarter split the BN/32 column chunks =====
        const int quarter = warp & 3;  // TMEM lane quarter this warp may access
        int it = 0;
        for (int tile = tile0; tile < num_tiles; tile += tile_stride, ++it) {
            int m_blk, n_blk;
            tile_mn(tile, m_blk, n_blk);
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1u;
            mbar_wait(&tmem_full[acc], acc_phase);
            tc_fence_after();
            for (int u = 0; u < MT; ++u) {
            int row = (m_blk * MT + u) * kBlockM + quarter * 32 + lane;
            bool row_ok = row < p.M;
            if constexpr (WIDE) {
                // widened row -> real output pixel; the two columns q' >= Wo of each image row are padding outputs
                const int line = row / p.Wwide;
                const int qq = row - line * p.Wwide;
                row_ok = row_ok && qq < p.Wo;
                row = line * p.Wo + qq;
            }
            const size_t row_off = static_cast<size_t>(row) * p.N + static_cast<size_t>(n_blk) * BN;
            const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                                   static_cast<uint32_t>(acc * Cfg::kAccStageCols + u * (MT == 1 ? 0 : BN));
#pragma unroll 1
            for (int c = (warp - 2) >> 2; c < BN / 32; c += kEpiWarps / 4) {
                uint32_t v[32];
                tmem_ld_32x32(t_row + static_cast<uint32_t>(c * 32), v);
                for (int a = 1; a < p.nacc; ++a) {
                    uint32_t u[32];
                    tmem_ld_wait();
                    tmem_ld_32x32(t_row + static_cast<uint32_t>(a * BN + c * 32), u);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(u[j]));
                }
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
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (PAIR) mbar_arrive_leader(&tmem_empty[acc]);
                else mbar_arrive(&tmem_empty[acc]);
            }
        }
    }

    tc_fence_before();
    if constexpr (PAIR) cluster_sync_all();   // neither CTA may exit (or free TMEM) while the pair's last MMAs / arrivals are in flight
    else __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        if constexpr (PAIR) tmem_dealloc_pair(tmem_base, Cfg::kTmemCols);
        else tmem_dealloc(tmem_base, Cfg::kTmemCols);
    }
}

}  // namespace bv
