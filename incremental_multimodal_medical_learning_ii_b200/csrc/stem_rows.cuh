// Row-streaming BioViL stem for 8-bit frames: conv 7x7 stride 2 pad 3 (1 folded input channel -> 64) + BatchNorm +
// ReLU + max-pool 3x3 stride 2 pad 1 -> NHWC bf16 input of layer1, as a warp-specialised pipeline that walks down a
// strip of the frame and never materialises a 2-D im2col tile.
//
// Replaces (reference): ResNetHIML.forward conv1 -> bn1 -> relu -> maxpool,
// health_multimodal/image/model/resnet.py:34-37, on frames made by ToTensor + ExpandChannels
// (health_multimodal/image/data/transforms.py:12-38).  Same arithmetic and rounding points as stem_fused.cuh
// (bf16 operands, fp32 accumulate incl. the BatchNorm bias, one bf16 rounding, max-pool, ReLU).
//
// Idea: with a NON-swizzled K-major shared-memory descriptor the two 16-byte K chunks of one tcgen05.mma K step may
// sit anywhere (their distance is the descriptor's LBO field).  Expand every input row ONCE horizontally,
//     E[y][m] = the 8 pixels under the filter for conv column cx(m)            (16 bytes, 128 GEMM rows -> 2 KB),
// and the im2col operand of conv row c is simply rows 2c-3 .. 2c+3 of E: chunk r of GEMM row m = E[2c-3+r][m].
// Vertical reuse costs nothing - each input row is expanded once (2 x 16 B per conv pixel instead of 7 x 16 B) -
// and two conv rows share one N = 128 MMA: A = [ones | E rows 4g-3 .. 4g+5] (10 chunks = 5 K steps),
// B rows 0..63 = filter rows against chunks 1..7, B rows 64..127 = the same filter rows against chunks 3..9,
// chunk 0 = BatchNorm bias (hi + mid + lo bf16 terms against 1.0).  One pooled output row therefore is
// 5 MMAs (128 x 128 x 16), 4 new E rows, and an epilogue that pools IN REGISTERS: vertical max against the conv row
// carried from the previous step, horizontal max with two warp shuffles.
//
// GEMM row m <-> conv column: lane quarter q = m / 32 covers local conv columns 30 q + (m % 32), i.e. quarters overlap
// by two columns so that every 3-wide pooling window lies inside one warp (TMEM lane quarters cannot be crossed).
// One strip = 60 pooled columns (4 quarters x 15) x the whole frame height; W = 480 is exactly two strips.
//
// Warps: 0 cp.async producer (raw bytes, 4 input rows per stage, zero fill outside the frame), 1-4 expand + u8 -> bf16,
// 5 MMA issue, 6.. epilogue (lane quarter x channel group, 4 x CG warps).  Rings: raw boxes (4), E groups of 4 rows (5),
// TMEM accumulators (4 x 128 columns).
#pragma once
#include "ptx.cuh"
#include "stem_fused.cuh"   // cp_async_8_zfill

namespace bv {

constexpr int kSrStripPx = 60;                    // pooled columns per strip
constexpr int kSrNG = 5;                          // E ring: groups of 4 expanded input rows
constexpr int kSrRowBytes = 128 * 16;             // one expanded input row
constexpr int kSrGroupBytes = 4 * kSrRowBytes;
constexpr int kSrRawStages = 8;                   // 1 KB each (4 / 8 / 16 stages measured equal)
constexpr int kSrRawPitch = 256;                  // bytes per raw input row in a box
constexpr int kSrRawBytes = 4 * kSrRawPitch;
constexpr int sr_tmem_stages(bool c1) { return c1 ? 3 : 4; }   // x 128 columns (two conv rows x 64 channels); the fused
                                                               // conv1 takes the last 128 columns (2 x 64)
constexpr int kSrALbo = 2048 + 64;                // pooled-pixel A operand of the fused conv1: chunk stride (the 64 skews
                                                  // the banks of the even / odd lanes that store chunk pairs)
constexpr int kSrABufBytes = 8 * kSrALbo;         // 128 pooled pixels x 64 channels, chunk-major
constexpr int kSrW1Bytes = 64 * 64 * 2;
constexpr int kSrChunks = 10;                     // K chunks of 8: ones + 9 input rows
constexpr int kSrWBytes = kSrChunks * 128 * 16;   // B operand: chunk-major, 128 rows x 16 B per chunk
__host__ __device__ constexpr int sr_threads(int cg) { return (6 + 4 * cg) * 32; }   // CG = channel groups of the epilogue (64 / CG channels per warp)
constexpr int kSrOnesBytes = 128 * 16;
constexpr int kSrRowBufBytes = 4 * 512;            // 4 rows x 256 bf16 pixels, double-buffered
constexpr int kSrBaseBytes = kSrOnesBytes + kSrNG * kSrGroupBytes + kSrWBytes + kSrRawStages * kSrRawBytes + 512 + 2 * kSrRowBufBytes;
constexpr int sr_smem_bytes(bool c1) { return kSrBaseBytes + (c1 ? 2 * kSrABufBytes + kSrW1Bytes + 256 : 0); }

struct StemRowsParams {
    const uint8_t* frames;        // u8 [B][H][W], 8-byte aligned
    const __nv_bfloat16* w;       // [64][64], k = r*8 + s, chunk 7 = BatchNorm bias hi/mid/lo (bv_weights.stem_u8_k8)
    __nv_bfloat16* out;           // [B][H/4][W/4][64]
    int B, H, W;
    int strips_x;                 // ceil((W/4) / 60)
    // fused layer1.0 conv1 (1x1, 64 -> 64, + folded bn1 + ReLU; resnet.py:39 / torchvision Bottleneck.forward): C1 kernels only
    const __nv_bfloat16* w1;      // [64 out][64 in] bf16
    const float* b1;              // [64]
    __nv_bfloat16* out1;          // [B][H/4][W/4][64]
    int debug;                    // BV_SR_DEBUG bisect bits (1 no loads, 2 no MMA; fused conv1: 4 no conv1 stores, 8 no conv1 epilogue, 16 no A-buffer stores)
};

// K-major descriptor WITHOUT swizzle (cute::UMMA LayoutType::SWIZZLE_NONE, canonical layout
// ((8,m),(8,2)):((16 B, SBO),(2 B, LBO))): rows 16 bytes apart inside an 8-row core matrix, 8-row groups SBO apart,
// the two K chunks of a K step LBO apart.
__device__ __forceinline__ uint64_t umma_desc_k_none(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    return d;
}

// four 8-bit pixels -> four bf16 (exact): 0x4B0000pp is the float 8388608 + pp
__device__ __forceinline__ void u8x4_to_bf16x4(uint32_t w, uint32_t& lo, uint32_t& hi) {
    const float f0 = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7440)) - 8388608.0f;
    const float f1 = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7441)) - 8388608.0f;
    const float f2 = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7442)) - 8388608.0f;
    const float f3 = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7443)) - 8388608.0f;
    // integers below 256 are exact in bf16: the pack is the two high halves
    lo = __byte_perm(__float_as_uint(f0), __float_as_uint(f1), 0x7632);
    hi = __byte_perm(__float_as_uint(f2), __float_as_uint(f3), 0x7632);
}

// two fp32 -> packed bf16x2 (lo in the low half), round to nearest even, ReLU folded into the conversion
__device__ __forceinline__ uint32_t pack_bf16x2_relu(uint32_t lo_f32, uint32_t hi_f32) {
    uint32_t d;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;\n" : "=r"(d) : "f"(__uint_as_float(hi_f32)), "f"(__uint_as_float(lo_f32)));
    return d;
}

__device__ __forceinline__ uint32_t bf16x2_max(uint32_t a, uint32_t b) {
    const __nv_bfloat162 r = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
    return *reinterpret_cast<const uint32_t*>(&r);
}

template <int N>
__device__ __forceinline__ void sr_tmem_ld(uint32_t taddr, uint32_t (&r)[N]) {
    static_assert(N == 16 || N == 32, "16 or 32 columns");
    if constexpr (N == 32) {
        tmem_ld_32x32(taddr, r);
    } else {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
              "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(taddr)
            : "memory");
    }
}

template <int CG, bool C1>
__global__ void __launch_bounds__(sr_threads(CG), 1) stem_rows_kernel(const __grid_constant__ StemRowsParams p) {
    constexpr int kThreads = sr_threads(CG);
    constexpr int kSrTmemStages = C1 ? 3 : 4;
    constexpr int kCh = 64 / CG;        // channels per epilogue warp
    constexpr int kPk = kCh / 2;        // packed bf16x2 registers per conv row
    extern __shared__ __align__(1024) uint8_t smem[];   // keep the shared address space visible (LDS/STS, not LD.E)
    if ((smem_u32(smem) & 127u) != 0u) __trap();
    uint8_t* ones = smem;                                   // 128 x [1,1,1,0,0,0,0,0]: sits BELOW the ring (LBO > 0)
    uint8_t* ring = ones + kSrOnesBytes;                    // kSrNG groups x 4 expanded rows
    uint8_t* wsm = ring + kSrNG * kSrGroupBytes;            // B operand
    uint8_t* raw = wsm + kSrWBytes;                         // raw byte boxes
    uint64_t* bars = reinterpret_cast<uint64_t*>(raw + kSrRawStages * kSrRawBytes);
    uint64_t* raw_full = bars;
    uint64_t* raw_empty = raw_full + kSrRawStages;
    uint64_t* e_full = raw_empty + kSrRawStages;
    uint64_t* e_empty = e_full + kSrNG;
    uint64_t* t_full = e_empty + kSrNG;
    uint64_t* t_empty = t_full + kSrTmemStages;
    uint64_t* a_full = t_empty + 4;          // fused conv1: pooled-pixel A buffers (2), conv1 accumulators (2)
    uint64_t* a_empty = a_full + 2;
    uint64_t* c1_full = a_empty + 2;
    uint64_t* c1_empty = c1_full + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(c1_empty + 2);
    uint8_t* rowbf = raw + kSrRawStages * kSrRawBytes + 512;   // bf16 row buffers of the expanders
    uint8_t* abuf = smem + kSrBaseBytes;     // C1 only
    uint8_t* w1sm = abuf + 2 * kSrABufBytes;
    float* b1sm = reinterpret_cast<float*>(w1sm + kSrW1Bytes);

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const int Hp = p.H / 4, Wp = p.W / 4;
    const int nstrips = p.B * p.strips_x;
    constexpr uint32_t nraw = kSrRawStages;
    const uint32_t groups_per_strip = static_cast<uint32_t>(Hp + 2);

    if (tid == 0) {
        for (int i = 0; i < kSrRawStages; ++i) {
            mbar_init(raw_full + i, 32);
            mbar_init(raw_empty + i, 128);
        }
        for (int i = 0; i < kSrNG; ++i) {
            mbar_init(e_full + i, 128);
            mbar_init(e_empty + i, 1);
        }
        for (int i = 0; i < kSrTmemStages; ++i) {
            mbar_init(t_full + i, 1);
            mbar_init(t_empty + i, 4 * CG);
        }
        if (C1) {
            for (int i = 0; i < 2; ++i) {
                mbar_init(a_full + i, 4 * CG * 32);
                mbar_init(a_empty + i, 1);
                mbar_init(c1_full + i, 1);
                mbar_init(c1_empty + i, 4 * CG);
            }
        }
        fence_barrier_init();
    }
    // constant operands, written once through the generic proxy
    if (tid < 128) *reinterpret_cast<uint4*>(ones + tid * 16) = make_uint4(0x3F803F80u, 0x00003F80u, 0u, 0u);
    for (int i = tid; i < kSrChunks * 128; i += kThreads) {
        const int c = i >> 7, n = i & 127;
        // rows 0..63: conv row 2g (filter row r against chunk 1 + r); rows 64..127: conv row 2g+1 (chunk 3 + r)
        const int r = (n < 64) ? c - 1 : c - 3;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (c == 0) v = *reinterpret_cast<const uint4*>(p.w + (n & 63) * 64 + 56);
        else if (r >= 0 && r < 7) {
            // the expanded chunks start ONE pixel left of the filter window (4-byte aligned in the bf16 row buffer):
            // filter column s sits at chunk position s + 1, position 0 carries a zero weight
            const uint4 w = *reinterpret_cast<const uint4*>(p.w + (n & 63) * 64 + r * 8);
            v = make_uint4(w.x << 16, __funnelshift_l(w.x, w.y, 16), __funnelshift_l(w.y, w.z, 16), __funnelshift_l(w.z, w.w, 16));
        }
        *reinterpret_cast<uint4*>(wsm + c * 2048 + n * 16) = v;
    }
    if (C1) {
        for (int i = tid; i < 8 * 64; i += kThreads) {   // conv1 weights, chunk-major: 64 rows x 16 B per chunk
            const int c = i >> 6, n = i & 63;
            *reinterpret_cast<uint4*>(w1sm + c * 1024 + n * 16) = *reinterpret_cast<const uint4*>(p.w1 + n * 64 + c * 8);
        }
        if (tid < 64) b1sm[tid] = p.b1[tid];
    }
    fence_proxy_async_smem();
    if (warp == 5) {
        tmem_alloc(tmem_ptr, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_launch_dependents();   // the next kernel's prologue may overlap this kernel's tail ...
    pdl_wait();                // ... and this kernel touches activations only once its predecessors have completed

    if (warp == 0) {
        // ---------------- raw byte producer ----------------
        // 4 input rows x 256 bytes per group as 8-byte cp.async pieces (TMA needs a 16-byte aligned box start; the
        // strip origin 240 s - 8 is only 8-byte aligned), zero fill outside the frame; the mbarrier arrival of each
        // lane fires when its copies have landed.  smem byte i of a row <-> input column 240 s - 8 + i.
        uint32_t n = 0;
        for (int strip = blockIdx.x; strip < nstrips; strip += gridDim.x) {
            const int b = strip / p.strips_x, s = strip - b * p.strips_x;
            const int x = 4 * kSrStripPx * s - 8 + 8 * lane;
            const bool x_ok = x >= 0 && x < p.W;
            const uint8_t* fb = p.frames + static_cast<size_t>(b) * p.H * p.W + (x_ok ? x : 0);
            for (int j = -2; j < Hp; ++j, ++n) {
                const uint32_t st = n % nraw;
                mbar_wait(raw_empty + st, ((n / nraw) & 1u) ^ 1u);
                if (!(p.debug & 1)) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int y = 4 * j + 2 + k;
                        const bool ok = x_ok && y >= 0 && y < p.H;
                        cp_async_8_zfill(raw + st * kSrRawBytes + k * kSrRawPitch + lane * 8,
                                         fb + static_cast<size_t>(ok ? y : 0) * p.W, ok);
                    }
                }
                asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(raw_full + st)) : "memory");
            }
        }
    } else if (warp <= 4) {
        // ---------------- horizontal expansion + u8 -> bf16 ----------------
        // phase 1: every pixel of the 4 raw rows is converted ONCE into a bf16 row buffer (thread t: row t / 32, 8 pixels);
        // phase 2: thread m copies the 8 pixels under conv column lcx(m) (element 2 lcx + 2 onwards: 4-byte aligned) of
        // each row into E.  The row buffer is double-buffered, so one named barrier per group orders both hazards.
        const int m = tid - 32;
        const int lcx = 30 * (m >> 5) + (m & 31);
        const int prow = m >> 5, pcol = m & 31;
        uint32_t n = 0;
        for (int strip = blockIdx.x; strip < nstrips; strip += gridDim.x) {
            for (int j = -2; j < Hp; ++j, ++n) {
                const uint32_t rs = n % nraw, es = n % kSrNG;
                uint8_t* rb = rowbf + (n & 1u) * kSrRowBufBytes;
                mbar_wait(raw_full + rs, (n / nraw) & 1u);
                {
                    const uint2 px = *reinterpret_cast<const uint2*>(raw + rs * kSrRawBytes + prow * kSrRawPitch + pcol * 8);
                    uint4 v;
                    u8x4_to_bf16x4(px.x, v.x, v.y);
                    u8x4_to_bf16x4(px.y, v.z, v.w);
                    *reinterpret_cast<uint4*>(rb + prow * 512 + pcol * 16) = v;
                }
                mbar_arrive(raw_empty + rs);
                asm volatile("bar.sync 1, 128;\n" ::: "memory");
                mbar_wait(e_empty + es, ((n / kSrNG) & 1u) ^ 1u);
                const uint8_t* src = rb + 4 * lcx + 4;
                uint8_t* dst = ring + es * kSrGroupBytes + m * 16;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const uint32_t* s32 = reinterpret_cast<const uint32_t*>(src + r * 512);
                    *reinterpret_cast<uint4*>(dst + r * kSrRowBytes) = make_uint4(s32[0], s32[1], s32[2], s32[3]);
                }
                fence_proxy_async_smem();
                mbar_arrive(e_full + es);
            }
        }
    } else if (warp == 5) {
        // ---------------- MMA issue ----------------
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc_bf16_f32(128, 128);
            const uint32_t ones_a = smem_u32(ones), ring_a = smem_u32(ring), w_a = smem_u32(wsm);
            uint32_t gbase = 0, sc = 0, su = 0;   // su: pairs of pooled rows ("super-steps"), counted across strips
            uint32_t next_c1 = 0;                 // first super-step whose conv1 has not been issued yet
            // conv1 of super-step x: [128 pooled pixels x 64] x W1^T, A operand written by the epilogue warps
            auto issue_c1 = [&](uint32_t x) {
                const uint32_t ab = x & 1u, ph = (x >> 1) & 1u;
                mbar_wait(a_full + ab, ph);
                mbar_wait(c1_empty + ab, ph ^ 1u);
                tc_fence_after();
                const uint32_t a_a = smem_u32(abuf) + ab * kSrABufBytes, w1_a = smem_u32(w1sm);
                const uint32_t d1 = tmem_base + 384u + ab * 64u;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16_ss(d1, umma_desc_k_none(a_a + 2 * k * kSrALbo, kSrALbo, 128),
                                 umma_desc_k_none(w1_a + 2 * k * 1024, 1024, 128), umma_idesc_bf16_f32(128, 64), k ? 1u : 0u);
                umma_commit(c1_full + ab);
                umma_commit(a_empty + ab);
            };
            for (int strip = blockIdx.x; strip < nstrips; strip += gridDim.x, gbase += groups_per_strip) {
                for (int g = 0; g < Hp; ++g, ++sc) {
                    // group G = gbase + j + 2 holds input rows 4j+2 .. 4j+5; this step reads j = g-2 (last row), g-1, g
                    const uint32_t G0 = gbase + g, G1 = G0 + 1, G2 = G0 + 2;
                    if (g == 0) {
                        mbar_wait(e_full + G0 % kSrNG, (G0 / kSrNG) & 1u);
                        mbar_wait(e_full + G1 % kSrNG, (G1 / kSrNG) & 1u);
                    }
                    mbar_wait(e_full + G2 % kSrNG, (G2 / kSrNG) & 1u);
                    const uint32_t ts = sc % kSrTmemStages;
                    mbar_wait(t_empty + ts, ((sc / kSrTmemStages) & 1u) ^ 1u);
                    tc_fence_after();
                    const uint32_t a0 = ring_a + (G0 % kSrNG) * kSrGroupBytes;
                    const uint32_t a1 = ring_a + (G1 % kSrNG) * kSrGroupBytes;
                    const uint32_t a2 = ring_a + (G2 % kSrNG) * kSrGroupBytes;
                    const uint32_t d = tmem_base + ts * 128u;
                    // K step 0: [ones | input row 4g-3 (last row of group g-2)]
                    const uint32_t row0 = a0 + 3 * kSrRowBytes;
                    if (!(p.debug & 2)) {
                    umma_bf16_ss(d, umma_desc_k_none(ones_a, row0 - ones_a, 128), umma_desc_k_none(w_a, 2048, 128), idesc, 0u);
                    // K steps 1..4: row pairs (4g-2, 4g-1), (4g, 4g+1), (4g+2, 4g+3), (4g+4, 4g+5)
                    umma_bf16_ss(d, umma_desc_k_none(a1, kSrRowBytes, 128), umma_desc_k_none(w_a + 2 * 2048, 2048, 128), idesc, 1u);
                    umma_bf16_ss(d, umma_desc_k_none(a1 + 2 * kSrRowBytes, kSrRowBytes, 128),
                                 umma_desc_k_none(w_a + 4 * 2048, 2048, 128), idesc, 1u);
                    umma_bf16_ss(d, umma_desc_k_none(a2, kSrRowBytes, 128), umma_desc_k_none(w_a + 6 * 2048, 2048, 128), idesc, 1u);
                    umma_bf16_ss(d, umma_desc_k_none(a2 + 2 * kSrRowBytes, kSrRowBytes, 128),
                                 umma_desc_k_none(w_a + 8 * 2048, 2048, 128), idesc, 1u);
                    }
                    umma_commit(t_full + ts);
                    umma_commit(e_empty + G0 % kSrNG);
                    if (g == Hp - 1) {
                        umma_commit(e_empty + G1 % kSrNG);
                        umma_commit(e_empty + G2 % kSrNG);
                    }
                    if (C1) {
                        // conv1 of a finished super-step is issued as soon as the epilogue has written its A buffer; the
                        // main MMAs never wait for it (they run up to three pooled rows ahead of the epilogue)
                        su += static_cast<uint32_t>(g & 1);
                        if (next_c1 < su && mbar_try_wait(a_full + (next_c1 & 1u), (next_c1 >> 1) & 1u)) issue_c1(next_c1++);
                    }
                }
            }
            if (C1)
                while (next_c1 < su) issue_c1(next_c1++);
        }
        __syncwarp();
    } else {
        // ---------------- epilogue: TMEM -> bf16 -> 3x3/2 max-pool in registers -> ReLU -> global ----------------
        const int q = warp & 3;            // TMEM lane quarter
        const int cg = (warp - 6) >> 2;    // channel group
        uint32_t sc = 0, su = 0;
        int pb = 0, ps = 0, pg0 = 0;       // frame, strip column and first pooled row of the previous super-step
        // second epilogue (fused conv1): accumulator lane m <-> pooled pixel (row pg0 + m / 64, strip column m % 64)
        auto conv1_out = [&](uint32_t x) {
            const uint32_t cb = x & 1u;
            mbar_wait(c1_full + cb, (x >> 1) & 1u);
            tc_fence_after();
            uint32_t r[kCh];
            sr_tmem_ld<kCh>(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + 384u + cb * 64u + static_cast<uint32_t>(cg * kCh), r);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(c1_empty + cb);
            if (p.debug & 8) return;
            const int m = 32 * q + lane, lpx = m & 63;
            const int pxx = kSrStripPx * ps + lpx;
            const bool ok = lpx < kSrStripPx && pxx < Wp;
            __nv_bfloat16* dst = p.out1 + ((static_cast<size_t>(pb) * Hp + pg0 + (m >> 6)) * Wp + pxx) * 64 + cg * kCh;
#pragma unroll
            for (int c = 0; c < kCh / 8; ++c) {
                const float4 ba = *reinterpret_cast<const float4*>(b1sm + cg * kCh + c * 8);
                const float4 bb = *reinterpret_cast<const float4*>(b1sm + cg * kCh + c * 8 + 4);
                const float bias[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
                uint4 o;
                uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const __nv_bfloat162 h = __floats2bfloat162_rn(__uint_as_float(r[c * 8 + 2 * e]) + bias[2 * e],
                                                                   __uint_as_float(r[c * 8 + 2 * e + 1]) + bias[2 * e + 1]);
                    ow[e] = bf16x2_max(*reinterpret_cast<const uint32_t*>(&h), 0u);
                }
                if (ok && !(p.debug & 4)) *reinterpret_cast<uint4*>(dst + c * 8) = o;
            }
        };
        for (int strip = blockIdx.x; strip < nstrips; strip += gridDim.x) {
            const int b = strip / p.strips_x, s = strip - b * p.strips_x;
            // local conv column 0 of strip 0 is conv column -1: max-pool padding, never wins
            const bool pad_warp = (s == 0 && q == 0);
            const int px = kSrStripPx * s + 15 * q + (lane >> 1);
            const bool store_ok = (lane >> 1) < 15 && px < Wp;
            uint32_t carry[kPk];
#pragma unroll
            for (int i = 0; i < kPk; ++i) carry[i] = 0u;   // conv row -1: padding (values are ReLU'd, 0 never wins wrongly)
            for (int g = 0; g < Hp; ++g, ++sc) {
                const uint32_t ts = sc % kSrTmemStages;
                mbar_wait(t_full + ts, (sc / kSrTmemStages) & 1u);
                tc_fence_after();
                uint32_t ra[kCh], rb[kCh];
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + ts * 128u + static_cast<uint32_t>(cg * kCh);
                sr_tmem_ld<kCh>(taddr, ra);
                sr_tmem_ld<kCh>(taddr + 64u, rb);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(t_empty + ts);
                uint32_t v[kPk];
                if (pad_warp) {   // warp-uniform: only the first quarter of strip 0 holds the padding column
#pragma unroll
                    for (int i = 0; i < kCh; ++i) {
                        ra[i] = lane == 0 ? 0xFF800000u : ra[i];   // -inf
                        rb[i] = lane == 0 ? 0xFF800000u : rb[i];
                    }
                }
#pragma unroll
                for (int i = 0; i < kPk; ++i) {
                    // ReLU rides on the conversion (relu and max commute, so this is the reference's relu -> maxpool order)
                    const uint32_t ua = pack_bf16x2_relu(ra[2 * i], ra[2 * i + 1]);
                    const uint32_t ub = pack_bf16x2_relu(rb[2 * i], rb[2 * i + 1]);
                    uint32_t x = bf16x2_max(bf16x2_max(carry[i], ua), ub);   // conv rows 2g-1, 2g, 2g+1
                    carry[i] = ub;
                    const uint32_t x1 = __shfl_down_sync(0xffffffffu, x, 1);
                    const uint32_t x2 = __shfl_down_sync(0xffffffffu, x, 2);
                    v[i] = bf16x2_max(bf16x2_max(x, x1), x2);                // conv columns 2j, 2j+1, 2j+2 (lane 2j)
                }
                // even lane 2j holds pooled pixel j (kCh channels = kCh / 8 chunks of 16 bytes); the odd neighbour takes
                // the odd chunks so that every store instruction writes whole 32-byte sectors
                __nv_bfloat16* dst = p.out + ((static_cast<size_t>(b) * Hp + g) * Wp + px) * 64 + cg * kCh + (lane & 1) * 8;
                uint8_t* arow = nullptr;
                if (C1) {
                    if (!(g & 1)) mbar_wait(a_empty + (su & 1u), ((su >> 1) & 1u) ^ 1u);   // conv1 of super-step su-2 has read it
                    arow = abuf + (su & 1u) * kSrABufBytes + (cg * (kCh / 8) + (lane & 1)) * kSrALbo +
                           (64 * (g & 1) + 15 * q + (lane >> 1)) * 16;
                }
#pragma unroll
                for (int i = 0; i < kCh / 16; ++i) {
                    uint4 o;
                    uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const uint32_t t = __shfl_sync(0xffffffffu, v[(2 * i + 1) * 4 + e], lane & ~1);
                        ow[e] = (lane & 1) ? t : v[2 * i * 4 + e];
                    }
                    if (store_ok) *reinterpret_cast<uint4*>(dst + i * 16) = o;
                    if (C1 && (lane >> 1) < 15 && !(p.debug & 16)) *reinterpret_cast<uint4*>(arow + 2 * i * kSrALbo) = o;
                }
                if (C1 && (g & 1)) {
                    fence_proxy_async_smem();
                    mbar_arrive(a_full + (su & 1u));
                    if (su >= 1) conv1_out(su - 1);
                    pb = b;
                    ps = s;
                    pg0 = g - 1;
                    ++su;
                }
            }
        }
        if (C1 && su >= 1) conv1_out(su - 1);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace bv
