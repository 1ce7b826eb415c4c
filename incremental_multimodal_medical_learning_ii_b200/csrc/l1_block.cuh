// One kernel for the tail of a layer1 Bottleneck and the head of the next one, on CTA pairs:
//
//   t2  = relu(conv2_3x3(t1) + b2)                      G0: tap-fused, N = 192 (conv3x3_tap3.cuh)
//   y   = relu(conv3_1x1(t2) + b3 + identity)           G1: N = 256, A = t2 tile written by the G0 epilogue   -> out1
//   t1' = relu(next.conv1_1x1(y) + b1')                 G2: N = N2,  A = y tile written by the G1 epilogue    -> out2
//
// Replaces (reference): torchvision Bottleneck.forward conv2/bn2/relu, conv3/bn3/+identity/relu and the next
// Bottleneck's conv1/bn1/relu under health_multimodal/image/model/resnet.py:39.  Neither t2 (0.94 GB per 512-frame
// batch) nor the re-read of y for conv1 (3.77 GB) touches HBM: per block the kernel reads t1 and the residual and
// writes y and t1' - 9.4 GB instead of 11.3 GB (two kernels) or 15.3 GB (three kernels).
//
// Why CTA pairs: the three weight matrices (72 + 32 + 32 KB) must stay resident next to the A ring and the staging
// tiles; with tcgen05 cta_group::2 each CTA holds HALF of every B operand (68 KB), which is what makes it fit.
// Each CTA of the pair owns its own 120-output tile (its 128 TMEM lanes): A loads, residual loads, epilogues and
// stores are CTA-local, only the MMAs and their barriers are joint (protocol: pair_gemm.cuh).
//
// Row space: as in the tap-fused 3x3 kernel a TMEM lane quarter holds 32 consecutive pixels of one image line and
// yields 30 outputs (the tap combination needs rows j+1, j+2).  Here the quarters are aligned to the image lines: the
// image width must be a multiple of 30 (layer1 of a 480-pixel frame: 120 = one tile per line), quarter k of a line
// covers outputs 30k .. 30k+29, so no quarter straddles a line and there are no padding outputs.  Residual loads
// (32 pixels) and output stores (30 pixels) are plain tiled 3-D TMA boxes over (C, W, B*H), always in bounds at
// their origin.  Other widths take the two-kernel path.
//
// Per tile t and CTA:     MMA thread (leader)              16 epilogue warps
//                         G1(t)      <- t2_ready(t)        E1(t):   D1 + bias + identity (in place in stg1), ReLU;
//                         G0(t+1)                                   the y sub-tiles feed G2 and the TMA stores
//                         G2(t)      <- sub_written[j]     E0(t+1): D0 -> taps combined, bias, ReLU -> t2 tile (smem)
//                                                          E2(t) (own warps): D2 + bias, ReLU -> t1' rows stored directly (128 B per thread)
// Warp roles (768 threads): 0 TMA producer (weights once, A ring), 1 MMA issuer (leader CTA) / idle (peer),
// 2..17 epilogue of G0/G1, 20..23 epilogue of G2 (its long wait for the second GEMM must not stall the others),
// 18 loader (residual prefetch into the stg1 sub-tiles as they drain), 19 storer (y and t1' leave
// through tiled 3-D TMA stores of [1 line x 30 pixels x 64 channels]: the two overlap rows of each lane quarter are
// not part of the box and the padding columns are out of bounds, so neither is written; per-thread copy-out loops
// stalled the epilogue warps on the store queue for 40 % of their time).
//
// DS = the FIRST block of the layer (torchvision Bottleneck with a downsample branch, resnet.py:39): there is no identity
// tensor; y = relu(conv3(t2) + b3 + downsample_1x1(x0) + b_ds) is one accumulation over two K segments ([t2 | x0] against
// [W3 | W_ds]) in the same TMEM tile.  The loader warp fetches the 64-channel x0 rows of the tile (16 KB, instead of 64 KB of
// identity rows) into their own A tile; the y staging ring carries no prefetched residual and shrinks to three sub-tiles,
// which pays for the W_ds panel and the x0 tile.  Neither t2 nor the re-read of y touches HBM: 6.6 GB per 512-frame batch
// instead of 10.4 GB for the tap-fused 3x3 kernel + the chained conv3/downsample/conv1 kernel it replaces.
#pragma once
#include "conv3x3_tap3.cuh"
#include "pair_gemm.cuh"

namespace bv {

constexpr int kL1Threads = 24 * 32;
#ifndef BV_L1_STAGES
#define BV_L1_STAGES 3
#endif
// A ring.  A tile needs three stages (one per filter row).  With two, the third load waits for the first MMA group of the SAME
// tile and its latency is exposed once per tile; with three, all loads of tile t+1 are issued while tile t is still in its
// epilogues - paid for with one y / identity staging sub-tile (shared memory is full).
constexpr int kL1Stages = BV_L1_STAGES;
// y / identity staging sub-tiles (L1Cfg::kYBufs): 4 per tile + 1 with an identity stream, so that the identity rows of the
// next tile's sub-tile j only wait for THIS tile's sub-tile j-1 to drain; 3 in the downsample form (nothing is prefetched
// into them)
constexpr int kL1DmaWarp = 18;
constexpr int kL1StoreWarp = 19;
constexpr int kL1E2Warp0 = 20;   // warps 20..23: epilogue of the second GEMM, one warp per TMEM lane quarter
// shared-memory map (bytes)
constexpr int kL1OffA = 0;                                 // kL1Stages x 16 KB A ring
constexpr int kL1OffW2 = kL1OffA + kL1Stages * kABytes;    // 3 filter rows x [96 rows x 128 B]
constexpr int kL1OffW3 = kL1OffW2 + 3 * 96 * 128;          // [128 rows x 128 B]
constexpr int kL1OffW1 = kL1OffW3 + 128 * 128;             // 4 k-blocks x [N2/2 rows x 128 B]

// SH = "shifted taps": the 3x3 GEMM is nine N = 64 MMAs per K step group - the three horizontal taps of a filter row read the
// SAME A stage through descriptors whose start is shifted by 0 / 1 / 2 pixel rows (128 B each; the 128B swizzle follows
// absolute address bits, exactly as kSegWide does in conv_gemm.cuh) and accumulate into ONE 64-column tile.  Against the
// tap-fused N = 192 form: 2.1x the tensor time of G0 (36 x 68 instead of 12 x 96 cycles per tile), but the accumulator
// shrinks from 192 to 64 TMEM columns - which lets it be double-buffered AND leaves room for a 128-wide second GEMM
// (D1 256 + D0 2 x 64 + D2 128 = 512 columns: the layer's LAST block, whose successor conv1 is layer2's) - and the first
// epilogue loses its 32 shuffles and two of its three TMEM loads per thread and tile (the LSU data pipe is 74-79 % busy
// in the N = 192 form, ncu).  Rows 30, 31 of every lane quarter read their taps from the next quarter's pixels (or past
// the stage): garbage in rows that were overlap rows anyway and are never stored.
template <int N2, bool DS = false, bool SH = false>
struct L1Cfg {
    static_assert(N2 == 64 || (N2 == 128 && SH && !DS), "TMEM plan: D1 256 + D0 192 + D2 64, or (shifted taps) D1 256 + D0 2 x 64 + D2 <= 128");
    static constexpr int kYBufs = (DS ? 4 : (N2 == 64 ? 6 : 5)) - (kL1Stages - 2);   // (t1' leaves through direct stores: no staging tile for it)
    static constexpr int kW1Bytes = 4 * (N2 / 2) * 128;
    static constexpr int kWdBytes = DS ? 128 * 128 : 0;          // this CTA's half of the downsample weights [128 rows x 128 B]
    static constexpr int kOffWd = kL1OffW1 + kW1Bytes;
    static constexpr int kOffX0 = kOffWd + kWdBytes;             // DS: x0 A tile [128 rows x 128 B]
    static constexpr int kOffT2 = kOffX0 + (DS ? kABytes : 0);
    static constexpr int kOffStg1 = kOffT2 + kABytes;            // kYBufs sub-tiles x 16 KB
    static constexpr int kOffBars = kOffStg1 + kYBufs * kStagingBytes;
    static constexpr int kNumBars = 2 * kL1Stages + 1 + 8 + 24 + kYBufs + 2 + 2 + 4;
    static constexpr int kSmemBytes = kOffBars + kNumBars * 8 + 16;
    static constexpr uint32_t kWeightBytes = 3 * 96 * 128 + 128 * 128 + kW1Bytes + kWdBytes;
    static_assert(kOffWd % 1024 == 0 && kOffX0 % 1024 == 0 && kOffT2 % 1024 == 0 && kOffStg1 % 1024 == 0,
                  "operand tiles need 1024-byte alignment");
    static_assert(kSmemBytes <= 232448, "exceeds the 227 KB of shared memory a CTA may use");
};

struct L1BlockParams {
    CUtensorMap tmA;    // t1 [B,H,W,64], widened im2col (3x3, pad 1), 32 pixels x 64 channels per load
    CUtensorMap tmRes;  // identity as (256, W, B*H), tiled, box 64 channels x 32 pixels x 1 line
    CUtensorMap tmW2;   // [64, 576]  box 64 x 32
    CUtensorMap tmW3;   // [256, 64]  box 64 x 128
    CUtensorMap tmW1;   // [N2, 256]  box 64 x N2/2
    CUtensorMap tmX0;   // DS: block input x0 as (64, W, B*H), tiled, box 64 channels x 32 pixels x 1 line
    CUtensorMap tmWd;   // DS: downsample weights [256, 64], box 64 x 128
    CUtensorMap tmOut1; // y   as (256, W, B*H), tiled, box 64 channels x 30 pixels x 1 line
    __nv_bfloat16* out2; // t1' [B*H][W][N2]: written with direct 128-byte-per-thread stores by the second-GEMM epilogue
    int lines;           // B * H
    // BatchNorm biases BY VALUE, i.e. in the kernel's constant bank: bias3 (+ downsample bias) [256] | bias2 [64] | bias1 [N2].
    // The epilogues read them with indexed constant loads (LDC): a broadcast LDS.128 costs two wavefronts on the LSU data
    // pipe, which made the bias the largest single consumer of that pipe (ncu source page: 43 M of 118 M shared-memory
    // wavefronts of the identity form) although it carries 1.5 KB.
    float4 bias_c[(256 + 64 + 128) / 4];
    int Ho, Wo;
    int groups_per_line;  // Wo / 30
    int num_groups;       // B * Ho * Wo / 30 lane quarters of work
    int num_tiles;        // ceil(num_groups / 4)
    int num_pair_tiles;   // ceil(num_tiles / 2)
    int l2_prefetch;      // pull the next tile's identity rows into L2 one tile ahead
    long long* dbg;       // optional [pairs][8] cycle counters of the leader's MMA thread (BV_TIMING)
};

// Tiled 3-D store of a [1 line][30 pixels][64 channels] box of an NHWC tensor viewed as (C, W, B*H); pixels whose
// column falls outside [0, W) are out of bounds and are not written - that is how the padding columns of the widened
// row space (and the part of a lane quarter that belongs to the neighbouring image line) are dropped.
// (im2col-mode TMA stores, which would express this directly, raise "illegal instruction" on this driver/GPU.)
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* src, int c, int w, int line) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];\n" ::"l"(
                     reinterpret_cast<uint64_t>(m)),
                 "r"(smem_u32(src)), "r"(c), "r"(w), "r"(line)
                 : "memory");
}

__device__ __forceinline__ void tma_load_3d(const CUtensorMap* m, uint64_t* bar, void* dst, int c, int w, int line,
                                            uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4, %5}], [%2], %6;\n" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c), "r"(w), "r"(line), "l"(policy)
        : "memory");
}
// the same load with its completion bytes credited to the LEADER's barrier (the pair MMA consumes both CTAs' tiles)
__device__ __forceinline__ void tma_load_3d_pair(const CUtensorMap* m, uint64_t* bar, void* dst, int c, int w, int line,
                                                 uint64_t policy) {
    const uint32_t bar_addr = smem_u32(bar) & kPeerBitMask;
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4, %5}], [%2], %6;\n" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_addr), "r"(c), "r"(w), "r"(line), "l"(policy)
        : "memory");
}
// Pull one box into L2 (no smem destination, no completion tracking).
__device__ __forceinline__ void tma_prefetch_l2_3d(const CUtensorMap* m, int c, int w, int line) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];\n" ::"l"(reinterpret_cast<uint64_t>(m)),
                 "r"(c), "r"(w), "r"(line)
                 : "memory");
}

// packed fp32x2 add (sm_100): halves the FADD count of the epilogues
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    float2 d;
    asm("{\n\t.reg .b64 ra, rb, rd;\n\t"
        "mov.b64 ra, {%2, %3};\n\t"
        "mov.b64 rb, {%4, %5};\n\t"
        "add.rn.f32x2 rd, ra, rb;\n\t"
        "mov.b64 {%0, %1}, rd;\n\t}\n"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}

// y sub-tile epilogue for 16 accumulator columns (two 16-byte groups, index g2 = 0..3 inside the 64-column sub-tile) of one
// row: + bias (`bp` = the four float4 of these columns) + residual (in place in the swizzled staging row), ReLU after the bf16
// rounding (max commutes with the rounding), bf16 pack.  RES = the staging row holds the identity values to add (otherwise it
// is only written)
template <bool RES = true>
__device__ __forceinline__ void l1_convert_row16(const uint32_t (&v)[16], const float4 (&bp)[4], uint8_t* row_ptr,
                                                 int l, int g2) {
    const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.0f, 0.0f);
#pragma unroll
    for (int j2 = 0; j2 < 2; ++j2) {
        const int jj = g2 * 2 + j2;
        uint4* sp = reinterpret_cast<uint4*>(row_ptr + ((jj ^ (l & 7)) << 4));
        uint4 rv = make_uint4(0u, 0u, 0u, 0u);
        if constexpr (RES) rv = *sp;
        const uint32_t r[4] = {rv.x, rv.y, rv.z, rv.w};
        const float4 b0 = bp[2 * j2], b1 = bp[2 * j2 + 1];
        const float2 bb[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w), make_float2(b1.x, b1.y), make_float2(b1.z, b1.w)};
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float2 a = make_float2(__uint_as_float(v[8 * j2 + 2 * e]), __uint_as_float(v[8 * j2 + 2 * e + 1]));
            a = add2(a, bb[e]);
            if constexpr (RES) a = add2(a, make_float2(__uint_as_float(r[e] << 16), __uint_as_float(r[e] & 0xFFFF0000u)));
            __nv_bfloat162 h = __floats2bfloat162_rn(a.x, a.y);
            h = __hmax2(h, zero2);
            w[e] = *reinterpret_cast<const uint32_t*>(&h);
        }
        *sp = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

template <int N2, bool DS = false, bool SH = false>
__global__ void __launch_bounds__(kL1Threads, 1) l1_block_kernel(const __grid_constant__ L1BlockParams p) {
    using Cfg = L1Cfg<N2, DS, SH>;
    constexpr int kL1YBufs = Cfg::kYBufs;
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0u) __trap();
    uint8_t* smem_a = smem + kL1OffA;
    uint8_t* smem_w2 = smem + kL1OffW2;
    uint8_t* smem_w3 = smem + kL1OffW3;
    uint8_t* smem_w1 = smem + kL1OffW1;
    uint8_t* smem_wd = smem + Cfg::kOffWd;      // DS only
    uint8_t* x0_tile = smem + Cfg::kOffX0;      // DS only
    uint8_t* t2_tile = smem + Cfg::kOffT2;
    uint8_t* stg1 = smem + Cfg::kOffStg1;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kOffBars);
    uint64_t* full_bar = bars;                    // [3] leader: A stage loaded in BOTH CTAs
    uint64_t* empty_bar = bars + kL1Stages;       // [3] per CTA: stage consumed (multicast commit)
    uint64_t* w_bar = bars + 2 * kL1Stages;       // leader: weights of both CTAs resident
    uint64_t* d0_full = w_bar + 1;                // per CTA (multicast commit)
    uint64_t* d0_empty = w_bar + 2;               // leader, 32 epilogue warps of the pair
    uint64_t* t2_ready = w_bar + 3;               // leader, 32
    uint64_t* t2_free = w_bar + 4;                // per CTA (multicast commit after G1)
    uint64_t* d1_full = w_bar + 5;                // per CTA
    uint64_t* d1_empty = w_bar + 6;               // leader, 32
    uint64_t* d2_full = w_bar + 7;                // per CTA
    uint64_t* d2_empty = w_bar + 8;               // leader, 32
    uint64_t* sub_written = w_bar + 9;            // [4] leader, 32 (16 warps x 2 CTAs): y sub-tile j complete
    // the loader waits on these one tile (five sub-tiles) behind the MMA / storer that complete them: two barrier sets,
    // used by even and odd tiles, keep a waiter from ever being two phases behind the barrier (parity aliasing)
    uint64_t* sub_consumed = sub_written + 4;     // [2][4] per CTA (multicast commit after G2 k-block j)
    uint64_t* store_done = sub_consumed + 8;      // [2][4] per CTA: the TMA stores of y sub-tile j have read smem
    uint64_t* res_ready = store_done + 8;         // [kL1YBufs] per CTA, per staging BUFFER: identity rows landed (TMA tx)
    uint64_t* y_local = res_ready + kL1YBufs;            // [4] per CTA, 16 warps: y sub-tile j written (for the storer)
    uint64_t* e2_local = y_local + 4;             // per CTA, 16 warps: t1' tile written
    uint64_t* stg2_free = e2_local + 1;           // per CTA: the TMA stores of t1' have read smem
    uint64_t* x0_full = stg2_free + 1;            // DS, leader: the x0 tiles of BOTH CTAs have landed (TMA tx)
    uint64_t* x0_free = x0_full + 1;              // DS, per CTA (multicast commit after G1): the x0 tile may be refilled
    uint64_t* d0b_full = x0_free + 1;             // SH: [2] per CTA, double-buffered 64-column conv2 accumulator
    uint64_t* d0b_empty = d0b_full + 2;           // SH: [2] leader, 32
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + Cfg::kNumBars);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1;
    const int num_pairs = gridDim.x >> 1;
    const int T = (p.num_pair_tiles - pair + num_pairs - 1) / num_pairs;  // tiles this CTA processes
    auto tile_of = [&](int t) { return 2 * (pair + t * num_pairs) + static_cast<int>(rank); };

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tmA);
        if constexpr (DS) {
            tma_prefetch_desc(&p.tmX0);
            tma_prefetch_desc(&p.tmWd);
        } else {
            tma_prefetch_desc(&p.tmRes);
        }
        tma_prefetch_desc(&p.tmW2);
        tma_prefetch_desc(&p.tmW3);
        tma_prefetch_desc(&p.tmW1);
        tma_prefetch_desc(&p.tmOut1);
        for (int i = 0; i < kL1Stages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        mbar_init(w_bar, 1);
        mbar_init(d0_full, 1);
        mbar_init(d0_empty, 32);
        mbar_init(t2_ready, 32);
        mbar_init(t2_free, 1);
        mbar_init(d1_full, 1);
        mbar_init(d1_empty, 32);
        mbar_init(d2_full, 1);
        mbar_init(d2_empty, 8);
        mbar_init(e2_local, 4);
        for (int b = 0; b < kL1YBufs; ++b) mbar_init(&res_ready[b], 1);
        mbar_init(stg2_free, 1);
        mbar_init(x0_full, 1);
        mbar_init(x0_free, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&d0b_full[i], 1);
            mbar_init(&d0b_empty[i], 32);
        }
        for (int j = 0; j < 4; ++j) {
            mbar_init(&sub_written[j], 32);
            mbar_init(&sub_consumed[j], 1);
            mbar_init(&sub_consumed[4 + j], 1);
            mbar_init(&store_done[j], 1);
            mbar_init(&store_done[4 + j], 1);
            mbar_init(&y_local[j], 16);
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
    pdl_launch_dependents();   // the next kernel's prologue may overlap this kernel's tail ...
    pdl_wait();                // ... and this kernel touches activations only once its predecessors have completed
    constexpr uint32_t kD1 = 0, kD0 = 256, kD2 = SH ? 384 : 448;   // TMEM column plan (SH: D0 = 2 x 64 columns at 256 / 320)

    // image line (global over the batch) and first output column of the four lane quarters of a tile
    auto group_coords = [&](int tile, int (&gl)[4], int (&gq)[4]) {
        const int g0 = tile * 4;
        gl[0] = g0 / p.groups_per_line;
        gq[0] = (g0 - gl[0] * p.groups_per_line) * kTap3Group;
#pragma unroll
        for (int g = 1; g < 4; ++g) {
            gl[g] = gl[g - 1];
            gq[g] = gq[g - 1] + kTap3Group;
            if (gq[g] >= p.Wo) {
                gq[g] = 0;
                ++gl[g];
            }
        }
    };

    if (T <= 0) {
        // nothing to do (the host never launches more pairs than pair tiles)
    } else if (warp == 0) {
        // ===================== TMA producer: weights once, then the A ring =====================
        if (elect_one()) {
            if (rank == 0) mbar_arrive_expect_tx(w_bar, 2u * Cfg::kWeightBytes);
            if constexpr (SH) {
                for (int tap = 0; tap < 9; ++tap)   // this CTA's 32 of the 64 cout rows of every tap
                    tma_load_2d_pair(&p.tmW2, w_bar, smem_w2 + tap * 4096, tap * kBlockK, 32 * static_cast<int>(rank), kEvictLast);
            } else {
            for (int tr = 0; tr < 3; ++tr)
                for (int b = 0; b < 3; ++b) {   // this CTA's 96 of the 192 (tap, cout) rows of filter row tr
                    const int n = 96 * static_cast<int>(rank) + 32 * b;
                    tma_load_2d_pair(&p.tmW2, w_bar, smem_w2 + tr * 12288 + b * 4096, (tr * 3 + n / 64) * kBlockK, n % 64,
                                     kEvictLast);
                }
            }
            tma_load_2d_pair(&p.tmW3, w_bar, smem_w3, 0, 128 * static_cast<int>(rank), kEvictLast);
            for (int kb = 0; kb < 4; ++kb)
                tma_load_2d_pair(&p.tmW1, w_bar, smem_w1 + kb * (N2 / 2) * 128, kb * kBlockK, (N2 / 2) * static_cast<int>(rank),
                                 kEvictLast);
            if constexpr (DS) tma_load_2d_pair(&p.tmWd, w_bar, smem_wd, 0, 128 * static_cast<int>(rank), kEvictLast);
        }
        __syncwarp();
        int stage = 0;
        uint32_t phase = 0;
        for (int t = 0; t < T; ++t) {
            int gl[4], gq[4], gi[4], gp[4];
            group_coords(tile_of(t), gl, gq);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                gi[g] = gl[g] / p.Ho;
                gp[g] = gl[g] - gi[g] * p.Ho;
            }
            for (int tr = 0; tr < 3; ++tr) {
                mbar_wait(&empty_bar[stage], phase ^ 1u);
                if (elect_one()) {
                    uint8_t* dst = smem_a + stage * kABytes;
                    if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2u * kABytes);
#pragma unroll
                    for (int g = 0; g < 4; ++g)
                        tma_load_im2col_4d_pair(&p.tmA, &full_bar[stage], dst + g * 4096, 0, gq[g] - 1, gp[g] - 1, gi[g], 0,
                                                static_cast<uint16_t>(tr), kEvictNormal);
                }
                __syncwarp();
                if (++stage == kL1Stages) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
            if constexpr (DS) {
                // x0 rows of this tile (row j of lane quarter q = pixel gq[q] + j of line gl[q], the output pixel of TMEM
                // lane 32 q + j).  Issued here, right behind the tile's last filter-row load: the first GEMM-1 of the
                // PREVIOUS tile (which frees the x0 tile) precedes the G0 that freed that ring stage, and G1 of this tile is
                // a whole tile time away - the load has microseconds to land.
                if (t > 0) mbar_wait(x0_free, (t - 1) & 1u);
                if (elect_one()) {
                    if (rank == 0) mbar_arrive_expect_tx(x0_full, 2u * kABytes);
#pragma unroll
                    for (int g = 0; g < 4; ++g)
                        tma_load_3d_pair(&p.tmX0, x0_full, x0_tile + g * 4096, 0, gq[g], gl[g], kEvictFirst);
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (rank == 0) {
            constexpr uint32_t idesc0 = umma_idesc_bf16_f32(256, kTap3N);
            constexpr uint32_t idesc1 = umma_idesc_bf16_f32(256, 256);
            constexpr uint32_t idesc2 = umma_idesc_bf16_f32(256, N2);
            const uint32_t base = smem_u32(smem);
            int stage = 0;
            uint32_t phase = 0;
            long long tw[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // total, d0_empty, full, d1_empty, t2_ready, d2_empty, sub_written
            const long long t_begin = tclock();
            auto timed_wait = [&](uint64_t* bar, uint32_t parity, int slot) {
                const long long a = tclock();
                mbar_wait_cluster(bar, parity);
                tw[slot] += tclock() - a;
            };
            auto trace = [&](int t, int ev) {
                if (kTimingBuild && p.dbg && pair == 0 && t >= 10 && t < 12 && lane == 0) p.dbg[2048 + (t - 10) * 32 + ev] = tclock();
            };
            auto g0 = [&](int t) {
                if constexpr (SH) timed_wait(&d0b_empty[t & 1], ((t >> 1) & 1u) ^ 1u, 1);
                else timed_wait(d0_empty, (t & 1u) ^ 1u, 1);
                tc_fence_after();
                for (int tr = 0; tr < 3; ++tr) {
                    timed_wait(&full_bar[stage], phase, 2);
                    tc_fence_after();
                    if (elect_one()) {
                        if constexpr (SH) {
                            constexpr uint32_t idesc64 = umma_idesc_bf16_f32(256, 64);
                            const uint32_t d_tmem = tmem_base + kD0 + static_cast<uint32_t>((t & 1) * 64);
#pragma unroll
                            for (int ts = 0; ts < 3; ++ts) {
                                // tap (tr, ts): the same stage, read from pixel row ts on
                                const uint64_t adesc = umma_desc_k_sw128(base + static_cast<uint32_t>(kL1OffA + stage * kABytes + ts * 128));
                                const uint64_t bdesc = umma_desc_k_sw128(base + static_cast<uint32_t>(kL1OffW2 + (tr * 3 + ts) * 4096));
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    umma_bf16_ss_pair(d_tmem, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k),
                                                      idesc64, (tr != 0 || ts != 0 || k != 0) ? 1u : 0u);
                            }
                            umma_commit_pair(&empty_bar[stage]);
                            if (tr == 2) umma_commit_pair(&d0b_full[t & 1]);
                        } else {
                        const uint64_t adesc = umma_desc_k_sw128(base + static_cast<uint32_t>(kL1OffA + stage * kABytes));
                        const uint64_t bdesc = umma_desc_k_sw128(base + static_cast<uint32_t>(kL1OffW2 + tr * 12288));
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16_ss_pair(tmem_base + kD0, adesc + static_cast<uint64_t>(2 * k),
                                              bdesc + static_cast<uint64_t>(2 * k), idesc0, (tr != 0 || k != 0) ? 1u : 0u);
                        umma_commit_pair(&empty_bar[stage]);
                        if (tr == 2) umma_commit_pair(d0_full);
                        }
                    }
                    if (tr == 2) trace(t, 0);   // G0 issued
                    __syncwarp();
                    if (++stage == kL1Stages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            };
            auto g1 = [&](int t) {
                timed_wait(d1_empty, (t & 1u) ^ 1u, 3);
                timed_wait(t2_ready, t & 1u, 4);
                if constexpr (DS) timed_wait(x0_full, t & 1u, 4);
                trace(t, 1);   // t2_ready seen
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t adesc = umma_desc_k_sw128(base + static_cast<uint32_t>(Cfg::kOffT2));
                    const uint64_t bdesc = umma_desc_k_sw128(base + static_cast<uint32_t>(kL1OffW3));
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16_ss_pair(tmem_base + kD1, adesc + static_cast<uint64_t>(2 * k),
                                          bdesc + static_cast<uint64_t>(2 * k), idesc1, k != 0 ? 1u : 0u);
                    if constexpr (DS) {   // second K segment: downsample_1x1(x0) accumulates into the same tile
                        const uint64_t xdesc = umma_desc_k_sw128(base + static_cast<uint32_t>(Cfg::kOffX0));
                        const uint64_t wdesc = umma_desc_k_sw128(base + static_cast<uint32_t>(Cfg::kOffWd));
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16_ss_pair(tmem_base + kD1, xdesc + static_cast<uint64_t>(2 * k),
                                              wdesc + static_cast<uint64_t>(2 * k), idesc1, 1u);
                        umma_commit_pair(x0_free);
                    }
                    umma_commit_pair(t2_free);
                    umma_commit_pair(d1_full);
                }
                __syncwarp();
                trace(t, 2);   // G1 issued
            };
            auto g2 = [&](int t) {
                timed_wait(d2_empty, (t & 1u) ^ 1u, 5);
                for (int j = 0; j < 4; ++j) {
                    timed_wait(&sub_written[j], t & 1u, 6);
                    trace(t, 3 + j);   // sub_written[j] seen
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t adesc =
                            umma_desc_k_sw128(base + static_cast<uint32_t>(Cfg::kOffStg1 + ((4 * t + j) % kL1YBufs) * kStagingBytes));
                        const uint64_t bdesc = umma_desc_k_sw128(base + static_cast<uint32_t>(kL1OffW1 + j * (N2 / 2) * 128));
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16_ss_pair(tmem_base + kD2, adesc + static_cast<uint64_t>(2 * k),
                                              bdesc + static_cast<uint64_t>(2 * k), idesc2, (j != 0 || k != 0) ? 1u : 0u);
                        umma_commit_pair(&sub_consumed[(t & 1) * 4 + j]);
                        if (j == 3) umma_commit_pair(d2_full);
                    }
                    __syncwarp();
                    if (j == 3) trace(t, 7);   // G2 issued
                }
            };
            // (Running E0(t+1) before E1(t) - to fill the wait for the identity rows - with G0/G1 one tile further ahead was
            // measured 10 % slower: the identity tile is the critical resource either way and the longer MMA queue
            // delays the G2 that frees it.)
            mbar_wait_cluster(w_bar, 0);
            g0(0);
            for (int t = 0; t < T; ++t) {
                g1(t);
                if (t + 1 < T) g0(t + 1);
                g2(t);
            }
            if (kTimingBuild && p.dbg && lane == 0) {
                tw[0] = tclock() - t_begin;
                for (int i = 0; i < 8; ++i) p.dbg[pair * 8 + i] = tw[i];
            }
        }
    } else if (warp == kL1DmaWarp) {
        // ===================== DMA: residual prefetch into the stg1 sub-tiles =====================
        // The y staging tile is single-buffered, so a residual sub-tile can only be requested once the previous tile's
        // sub-tile has drained; its HBM latency would be exposed every tile.  The NEXT tile's residual is therefore
        // pulled into L2 one tile ahead (short distance: the lines are still there when the real load follows).
        if (lane == 0) {
            for (int t = 0; t < T; ++t) {
                int gl[4], gq[4];
                group_coords(tile_of(t), gl, gq);
                if (!DS && p.l2_prefetch && t + 1 < T) {
                    int nl[4], nq[4];
                    group_coords(tile_of(t + 1), nl, nq);
                    for (int j = 0; j < 4; ++j)
#pragma unroll
                        for (int g = 0; g < 4; ++g) tma_prefetch_l2_3d(&p.tmRes, j * kChunkCols, nq[g], nl[g]);
                }
                for (int j = 0; j < 4; ++j) {
                    // sub-tile g = 4t + j lives in buffer g % 5; its previous tenant was sub-tile g - 5 = (t - 1, j - 1)
                    // (or (t - 2, 3) for j = 0): wait until the second GEMM has read it and its TMA store has drained
                    const int g = 4 * t + j, gprev = g - kL1YBufs;
                    if (gprev >= 0) {
                        const int tp = gprev >> 2, jp = gprev & 3;
                        mbar_wait(&sub_consumed[(tp & 1) * 4 + jp], (tp >> 1) & 1u);
                        mbar_wait(&store_done[(tp & 1) * 4 + jp], (tp >> 1) & 1u);
                    }
                    const int b = g % kL1YBufs;
                    if constexpr (DS) {
                        mbar_arrive(&res_ready[b]);   // nothing to prefetch: the buffer is simply free again
                    } else {
                        mbar_arrive_expect_tx(&res_ready[b], kStagingBytes);
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            tma_load_3d(&p.tmRes, &res_ready[b], stg1 + b * kStagingBytes + q * 4096, j * kChunkCols, gq[q],
                                        gl[q], kEvictFirst);
                    }
                }
            }
        }
    } else if (warp == kL1StoreWarp) {
        // ===================== storer: y sub-tiles and the t1' tile leave through im2col-mode TMA stores =====================
        if (lane == 0) {
            for (int t = 0; t < T; ++t) {
                int gl[4], gq[4];
                group_coords(tile_of(t), gl, gq);
                uint64_t* sd = store_done + (t & 1) * 4;
                for (int j = 0; j < 4; ++j) {
                    mbar_wait(&y_local[j], t & 1u);
#pragma unroll
                    for (int g = 0; g < 4; ++g)
                        tma_store_3d(&p.tmOut1, stg1 + ((4 * t + j) % kL1YBufs) * kStagingBytes + g * 4096, j * kChunkCols, gq[g],
                                     gl[g]);
                    tma_store_commit();
                    if constexpr (DS) {
                        // three staging buffers: sub-tile 3 of THIS tile reuses the buffer of sub-tile 0, so "drained" must
                        // be reported as the stores go, not after the tile's last one (which waits for sub-tile 3)
                        if (j > 0) {
                            tma_store_wait_read<1>();
                            mbar_arrive(&sd[j - 1]);
                        }
                    }
                }
                // bulk groups complete in order: allow the 3 - j most recent ones to be pending
                if constexpr (!DS) {
                    tma_store_wait_read<3>();
                    mbar_arrive(&sd[0]);
                    tma_store_wait_read<2>();
                    mbar_arrive(&sd[1]);
                    tma_store_wait_read<1>();
                    mbar_arrive(&sd[2]);
                }
                tma_store_wait_read<0>();
                mbar_arrive(&sd[3]);
            }
            tma_store_wait_all<0>();
        }
    } else if (warp >= kL1E2Warp0) {
        // ===================== epilogue of the second GEMM: D2 -> + bias, ReLU -> stg2 (one warp per lane quarter) =====================
        const int quarter = warp & 3;
        const int l = quarter * 32 + lane;
        const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
        const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.0f, 0.0f);
        for (int t = 0; t < T; ++t) {
            mbar_wait(d2_full, t & 1u);
            tc_fence_after();
            int gl[4], gq[4];
            group_coords(tile_of(t), gl, gq);
            // this thread's row = pixel gq[quarter] + lane of line gl[quarter]; lanes 30, 31 hold the overlap rows of the lane
            // quarter and the last tile may reach past the last line: neither is stored.  64 channels = one 128-byte line.
            const bool store_ok = lane < kTap3Group && gl[quarter] < p.lines;
            uint4* op = reinterpret_cast<uint4*>(p.out2 + (static_cast<size_t>(gl[quarter]) * p.Wo + gq[quarter] + lane) * N2);
#pragma unroll 1
            for (int half = 0; half < N2 / 64; ++half) {
                uint32_t w[32];
#pragma unroll
                for (int c16 = 0; c16 < 4; ++c16) {
                    uint32_t v[16];
                    tmem_ld_32x16(lane_base + kD2 + static_cast<uint32_t>(half * 64 + c16 * 16), v);
                    tmem_ld_wait();
                    const int bi = (320 + half * 64 + c16 * 16) >> 2;
                    const float4 bq[4] = {p.bias_c[bi], p.bias_c[bi + 1], p.bias_c[bi + 2], p.bias_c[bi + 3]};
#pragma unroll
                    for (int c = 0; c < 16; c += 2) {
                        const float4 b4 = bq[c >> 2];
                        const float2 a = add2(make_float2(__uint_as_float(v[c]), __uint_as_float(v[c + 1])),
                                              (c & 2) ? make_float2(b4.z, b4.w) : make_float2(b4.x, b4.y));
                        __nv_bfloat162 h = __floats2bfloat162_rn(a.x, a.y);
                        h = __hmax2(h, zero2);
                        w[c16 * 8 + (c >> 1)] = *reinterpret_cast<const uint32_t*>(&h);
                    }
                }
                if (half == N2 / 64 - 1) {   // the accumulator has been read completely: the leader may start the next tile's G2
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_leader(d2_empty);
                }
                if (store_ok) {
#pragma unroll
                    for (int ch = 0; ch < 8; ++ch) op[half * 8 + ch] = make_uint4(w[4 * ch], w[4 * ch + 1], w[4 * ch + 2], w[4 * ch + 3]);
                }
            }
        }
        (void)l;
    } else if (warp >= 2 && warp < 18) {
        // ===================== epilogue warps of the first two GEMMs =====================
        const int quarter = warp & 3;
        const int cg = (warp - 2) >> 2;                 // 16-column group (E0, E2) / y sub-tile (E1)
        const int l = quarter * 32 + lane;              // TMEM lane = row of every smem tile
        const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);

        long long te[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        long long tl = tclock();
        const long long te_begin = tl;
        auto lap = [&](int slot) {   // cycles since the previous lap go to `slot`
            const long long now = tclock();
            te[slot] += now - tl;
            tl = now;
        };
        auto gtime = []() { unsigned long long g = 0; if constexpr (kTimingBuild) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)); return static_cast<long long>(g); };
        auto etrace = [&](int t, int ev) {
            if (kTimingBuild && p.dbg && pair == 0 && rank == 0 && warp == 2 && lane == 0 && t >= 10 && t < 12) p.dbg[2048 + (t - 10) * 32 + ev] = tclock();
            // wall-clock (ns) copies for cross-SM comparison: leader at [16 + ev - 8], peer at [24 + ev - 8]
            if (kTimingBuild && p.dbg && pair == 0 && warp == 2 && lane == 0 && t >= 10 && t < 12 && ev >= 8)
                p.dbg[2048 + (t - 10) * 32 + (rank == 0 ? 16 : 24) + (ev - 8)] = gtime();
        };
        auto e0 = [&](int t) {
            lap(11);
            if constexpr (SH) {
                // shifted taps: the accumulator already holds the nine taps summed per output pixel - one TMEM load, no shuffles
                mbar_wait(&d0b_full[t & 1], (t >> 1) & 1u);
                lap(0);
                tc_fence_after();
                uint32_t v[16];
                tmem_ld_32x16(lane_base + kD0 + static_cast<uint32_t>((t & 1) * 64 + cg * 16), v);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_leader(&d0b_empty[t & 1]);
                uint32_t w[8];
                const int bi = (256 + cg * 16) >> 2;
                const float4 bq[4] = {p.bias_c[bi], p.bias_c[bi + 1], p.bias_c[bi + 2], p.bias_c[bi + 3]};
#pragma unroll
                for (int c = 0; c < 16; c += 2) {
                    const float4 b4 = bq[c >> 2];
                    const float2 a = add2(make_float2(__uint_as_float(v[c]), __uint_as_float(v[c + 1])),
                                          (c & 2) ? make_float2(b4.z, b4.w) : make_float2(b4.x, b4.y));
                    __nv_bfloat162 h = __floats2bfloat162_rn(a.x, a.y);
                    h = __hmax2(h, __floats2bfloat162_rn(0.0f, 0.0f));
                    w[c >> 1] = *reinterpret_cast<const uint32_t*>(&h);
                }
                lap(1);
                mbar_wait(t2_free, (t & 1u) ^ 1u);
                lap(2);
                uint8_t* rp = t2_tile + l * 128;
                *reinterpret_cast<uint4*>(rp + (((2 * cg) ^ (l & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
                *reinterpret_cast<uint4*>(rp + (((2 * cg + 1) ^ (l & 7)) << 4)) = make_uint4(w[4], w[5], w[6], w[7]);
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive_leader(t2_ready);
                lap(1);
                return;
            }
            mbar_wait(d0_full, t & 1u);
            etrace(t, 8);    // d0_full seen
            lap(0);
            tc_fence_after();
            uint32_t v0[16], v1[16], v2[16];
            const uint32_t ta = lane_base + kD0 + static_cast<uint32_t>(cg * 16);
            tmem_ld_32x16(ta, v0);
            tmem_ld_32x16(ta + 64u, v1);
            tmem_ld_32x16(ta + 128u, v2);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(d0_empty);
            uint32_t w[8];
            const int bi = (256 + cg * 16) >> 2;
            const float4 bq[4] = {p.bias_c[bi], p.bias_c[bi + 1], p.bias_c[bi + 2], p.bias_c[bi + 3]};
#pragma unroll
            for (int c = 0; c < 16; c += 2) {
                const float4 b4 = bq[c >> 2];
                float2 u1, u2;
                u1.x = __shfl_down_sync(0xffffffffu, __uint_as_float(v1[c]), 1);
                u1.y = __shfl_down_sync(0xffffffffu, __uint_as_float(v1[c + 1]), 1);
                u2.x = __shfl_down_sync(0xffffffffu, __uint_as_float(v2[c]), 2);
                u2.y = __shfl_down_sync(0xffffffffu, __uint_as_float(v2[c + 1]), 2);
                const float2 a = add2(make_float2(__uint_as_float(v0[c]), __uint_as_float(v0[c + 1])), u1);
                const float2 b = add2(u2, (c & 2) ? make_float2(b4.z, b4.w) : make_float2(b4.x, b4.y));
                const float2 f2 = add2(a, b);
                const float f[2] = {f2.x, f2.y};
                __nv_bfloat162 h = __floats2bfloat162_rn(f[0], f[1]);
                h = __hmax2(h, __floats2bfloat162_rn(0.0f, 0.0f));
                w[c >> 1] = *reinterpret_cast<const uint32_t*>(&h);
            }
            lap(1);
            mbar_wait(t2_free, (t & 1u) ^ 1u);   // the first GEMM-1 of the previous tile has read the t2 tile
            lap(2);
            uint8_t* rp = t2_tile + l * 128;
            *reinterpret_cast<uint4*>(rp + (((2 * cg) ^ (l & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
            *reinterpret_cast<uint4*>(rp + (((2 * cg + 1) ^ (l & 7)) << 4)) = make_uint4(w[4], w[5], w[6], w[7]);
            // cta-scope release on purpose: a cluster-scope release would also wait for this thread's outstanding GLOBAL
            // stores of the previous copy-out (measured: 45 % of the MMA thread's time went into this barrier); the data
            // the pair MMA reads is this SM's own shared memory, ordered by the proxy fence above
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(t2_ready);
            etrace(t, 9);    // t2_ready arrived
            lap(1);
        };
        // All 16 warps work on ONE 64-column sub-tile at a time (warp (quarter, cg) converts 16 of its columns), so
        // sub-tile 0 is complete after a quarter of E1: its second-GEMM k-block, its TMA store and - most important -
        // the reload of the NEXT tile's identity rows into it start while sub-tiles 1..3 are still being converted.
        auto e1 = [&](int t) {
            lap(11);
            mbar_wait(d1_full, t & 1u);
            etrace(t, 10);   // d1_full seen
            lap(3);
            tc_fence_after();
#pragma unroll 1
            for (int j = 0; j < 4; ++j) {
                const int yb = (4 * t + j) % kL1YBufs;
                uint8_t* sub = stg1 + yb * kStagingBytes;
                uint32_t v[16];
                tmem_ld_32x16(lane_base + kD1 + static_cast<uint32_t>(j * kChunkCols + cg * 16), v);
                mbar_wait(&res_ready[yb], ((4 * t + j) / kL1YBufs) & 1u);
                if (j == 0) etrace(t, 11);   // res_ready seen
                tmem_ld_wait();
                if (j == 3) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_leader(d1_empty);
                }
                const int bi = (j * kChunkCols + cg * 16) >> 2;
                const float4 bq[4] = {p.bias_c[bi], p.bias_c[bi + 1], p.bias_c[bi + 2], p.bias_c[bi + 3]};
                l1_convert_row16<!DS>(v, bq, sub + l * 128, l, cg);
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive_leader(&sub_written[j]);
                    mbar_arrive(&y_local[j]);
                }
            }
            etrace(t, 12);   // E1 done
            lap(5);
        };
        e0(0);
        for (int t = 0; t < T; ++t) {
            e1(t);
            if (t + 1 < T) e0(t + 1);
        }
        if (kTimingBuild && p.dbg && rank == 0 && warp == 2 && lane == 0) {
            for (int i = 0; i < 12; ++i) p.dbg[1024 + pair * 12 + i] = te[i];
            p.dbg[1024 + 74 * 12 + pair] = tclock() - te_begin;
        }
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, 512);
    }
}

}  // namespace bv
