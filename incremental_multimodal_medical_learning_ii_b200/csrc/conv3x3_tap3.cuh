// 3x3 stride-1 convolution with 64 input and 64 output channels (layer1's conv2): the three HORIZONTAL taps of a
// filter row are one tcgen05.mma of N = 192.
//
// A tcgen05.mma with N <= 128 is bound by the 64 B/cycle A-operand read, not by the tensor core: N = 64 costs 68
// cycles, N = 192 only 96 (tools/umma_microbench.cu).  The wide-row kernel (conv_gemm.cuh, kSegWide) issues three
// N = 64 MMAs per filter row and K step, each re-reading the same A tile through a descriptor shifted by ts pixels.
// Here the shift moves to the OUTPUT side instead:
//
//     D[j, ts*64 + c] = sum_{tr, k} A_tr[j, k] * W[tr][ts][c, k]          one MMA per (tr, K step), N = 192
//     out[j, c]       = D[j, c] + D[j+1, 64 + c] + D[j+2, 128 + c]        epilogue: rows j+1 / j+2 via warp shuffles
//
// with j running over consecutive pixels of the zero-padded image row (the "wide" row space of conv_gemm.cuh, the
// TMA im2col box widened by the padding); the two columns q' >= Wo of every image row are padding outputs and are
// dropped, as in the wide kernel.  MMA work per tile drops from 36 x 68+ to 12 x 96 cycles.
//
// Replaces (reference): Bottleneck.conv2 + bn2 + relu (torchvision resnet.py Bottleneck.forward under
// health_multimodal/image/model/resnet.py:39) for the 64-channel blocks of layer1.
//
// The shift stays inside a warp: TMEM lane quarter q (32 rows) is loaded with the 32 pixels that START at output
// 30q of the tile, so lanes 0..29 of every epilogue warp find rows j+1 / j+2 in lanes 1..31 of the same warp (two
// __shfl_down per value, no shared-memory exchange, no inter-warp barrier).  A tile is therefore four 32-pixel TMA
// im2col loads per filter row, yields 4 x 30 = 120 outputs and advances by 120 rows of the widened pixel space.
// (Measured alternatives, same parity: 128 consecutive pixels with a cross-warp exchange of the two boundary rows
// through smem + pair barriers is epilogue-issue-bound and 20 % slower; per-thread row stores instead of the staged
// copy-out touch 30 cache lines per instruction and are 3x slower.)
//
// Warp roles (576 threads, 1 CTA / SM, persistent): 0 TMA producer, 1 MMA issuer (+TMEM alloc), 2..17 epilogue in
// two groups of eight warps.  Group g owns TMEM accumulator g and therefore every other tile: the epilogue of a tile
// is a chain of latencies (mbarrier poll, tcgen05.ld, shuffles, barrier, stores) of about two MMA tile times, and two
// groups working on alternate tiles hide it.  Inside a group there are two warps per lane quarter, each converting 32
// of the 64 channels in two passes of 16.  Phase 1 leaves the bf16 tile in a swizzled smem buffer (one per group);
// phase 2 (after a 256-thread named barrier) writes full 128-byte rows, four rows per warp instruction.
#pragma once
#include "conv_gemm.cuh"

namespace bv {

constexpr int kTap3EpiWarps = 16;
constexpr int kTap3Threads = (2 + kTap3EpiWarps) * 32;
constexpr int kTap3Group = 30;               // outputs per lane quarter
constexpr int kTap3Rows = 4 * kTap3Group;    // outputs per tile
constexpr int kTap3Stages = 6;               // A ring: one 4 x 32-pixel x 64-channel tile per filter row
constexpr int kTap3N = 192;
constexpr int kTap3WBytes = 9 * 64 * 128;    // resident weights, slot (tr*3 + ts) = [64 cout][64 cin] bf16
constexpr int kTap3StageOut = kTap3Rows * 128;   // bf16 output tile staged in smem (16-byte chunks XOR-swizzled by row)
constexpr int kTap3NumBars = 2 * kTap3Stages + 4 + 1;
constexpr int kTap3SmemBytes = kTap3Stages * kABytes + kTap3WBytes + 2 * kTap3StageOut + 2 * kBlockM * 4 +
                               kTap3NumBars * 8 + 16;

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(threads) : "memory");
}


__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}

// p.M = rows of the widened pixel space (B * Ho * (Wo + 2)), p.Wwide = Wo + 2, p.num_m_blocks = ceil(M / 120);
// p.tmA[0] = im2col map over the widened bounding box with 32 pixels per load.
__global__ void __launch_bounds__(kTap3Threads, 1) conv3x3_tap3_kernel(const __grid_constant__ ConvGemmParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0u) __trap();
    uint8_t* smem_a = smem;
    uint8_t* smem_w = smem + kTap3Stages * kABytes;
    uint8_t* stage_out = smem_w + kTap3WBytes;                                   // 2 x [120 rows][128 B]
    int* rowoff = reinterpret_cast<int*>(stage_out + 2 * kTap3StageOut);         // 2 x [128] real output row or -1
    uint64_t* bars = reinterpret_cast<uint64_t*>(rowoff + 2 * kBlockM);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + kTap3Stages;
    uint64_t* tmem_full = bars + 2 * kTap3Stages;
    uint64_t* tmem_empty = tmem_full + 2;
    uint64_t* w_bar = tmem_full + 4;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + kTap3NumBars);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_tiles = p.num_m_blocks;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tmA[0]);
        tma_prefetch_desc(&p.tmB[0]);
        for (int i = 0; i < kTap3Stages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], kTap3EpiWarps / 2);
        }
        mbar_init(w_bar, 1);
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

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            mbar_arrive_expect_tx(w_bar, kTap3WBytes);
            for (int slot = 0; slot < 9; ++slot)
                tma_load_2d(&p.tmB[0], w_bar, smem_w + slot * 8192, slot * kBlockK, 0, kEvictLast);
        }
        __syncwarp();
        int stage = 0;
        uint32_t phase = 0;
        const int hww = p.Ho * p.Wwide;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            // window origins of the four 32-pixel groups (group g starts 30 g widened pixels into the tile)
            const int m0 = tile * kTap3Rows;
            int gi[4], gp[4], gq[4];
            gi[0] = m0 / hww;
            const int rem = m0 - gi[0] * hww;
            gp[0] = rem / p.Wwide;
            gq[0] = rem - gp[0] * p.Wwide;
#pragma unroll
            for (int g = 1; g < 4; ++g) {
                gi[g] = gi[g - 1];
                gp[g] = gp[g - 1];
                gq[g] = gq[g - 1] + kTap3Group;
                while (gq[g] >= p.Wwide) {
                    gq[g] -= p.Wwide;
                    if (++gp[g] == p.Ho) {
                        gp[g] = 0;
                        ++gi[g];
                    }
                }
            }
            for (int tr = 0; tr < 3; ++tr) {
                mbar_wait(&empty_bar[stage], phase ^ 1u);
                if (elect_one()) {
                    uint8_t* dst = smem_a + stage * kABytes;
                    mbar_arrive_expect_tx(&full_bar[stage], kABytes);
#pragma unroll
                    for (int g = 0; g < 4; ++g)
                        tma_load_im2col_4d(&p.tmA[0], &full_bar[stage], dst + g * 4096, 0, gq[g] - 1, gp[g] - 1, gi[g], 0,
                                           static_cast<uint16_t>(tr), kEvictNormal);
                }
                __syncwarp();
                if (++stage == kTap3Stages) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc = umma_idesc_bf16_f32(kBlockM, kTap3N);
        const uint32_t a_base = smem_u32(smem_a);
        const uint32_t w_base = smem_u32(smem_w);
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        long long t_acc = 0, t_full = 0;
        const long long t_begin = tclock();
        mbar_wait(w_bar, 0);
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            long long tw = tclock();
            mbar_wait(&tmem_empty[acc], ((it >> 1) & 1u) ^ 1u);
            t_acc += tclock() - tw;
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * 256);
            for (int tr = 0; tr < 3; ++tr) {
                tw = tclock();
                mbar_wait(&full_bar[stage], phase);
                t_full += tclock() - tw;
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t adesc = umma_desc_k_sw128(a_base + static_cast<uint32_t>(stage * kABytes));
                    const uint64_t bdesc = umma_desc_k_sw128(w_base + static_cast<uint32_t>(tr * 3 * 8192));
#pragma unroll
                    for (int k = 0; k < kBlockK / 16; ++k)
                        umma_bf16_ss(d_tmem, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k),
                                     idesc, (tr != 0 || k != 0) ? 1u : 0u);
                    umma_commit(&empty_bar[stage]);
                    if (tr == 2) umma_commit(&tmem_full[acc]);
                }
                __syncwarp();
                if (++stage == kTap3Stages) {
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
    } else {
        // ===================== epilogue (warps 2..17) =====================
        const int quarter = warp & 3;
        const int group = (warp - 2) >> 3;           // accumulator / staging buffer / tile parity of this warp
        const int hc = ((warp - 2) >> 2) & 1;        // which 32 of the 64 output channels
        const int e = (warp - 2) & 7;                // phase 2: rows 16e .. 16e+15 of the staged tile
        const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                                   static_cast<uint32_t>(group * 256 + hc * 32);
        uint8_t* out = reinterpret_cast<uint8_t*>(p.out);
        uint8_t* so = stage_out + group * kTap3StageOut;
        int* ro = rowoff + group * kBlockM;
        float bias[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) bias[c] = __ldg(p.bias[0] + hc * 32 + c);
        const bool relu = p.relu != 0;
        const int r_tile = quarter * kTap3Group + lane;   // row of the tile this thread produces (lane < 30)
        // widened row of this thread's output, kept as (line, qq) and advanced without divisions
        const int first = static_cast<int>(blockIdx.x) + group * static_cast<int>(gridDim.x);
        int row = first * kTap3Rows + r_tile;
        int line = row / p.Wwide;
        int qq = row - line * p.Wwide;
        const int step = 2 * static_cast<int>(gridDim.x) * kTap3Rows;
        const int step_lines = step / p.Wwide, step_q = step - step_lines * p.Wwide;
        uint32_t ph = 0;
        for (int tile = first; tile < num_tiles; tile += 2 * gridDim.x, ph ^= 1u) {
            mbar_wait(&tmem_full[group], ph);
            tc_fence_after();
#pragma unroll
            for (int sub = 0; sub < 2; ++sub) {
                uint32_t v0[16], v1[16], v2[16];
                const uint32_t t = lane_base + static_cast<uint32_t>(sub * 16);
                tmem_ld_32x16(t, v0);
                tmem_ld_32x16(t + 64u, v1);
                tmem_ld_32x16(t + 128u, v2);
                tmem_ld_wait();
                if (sub == 1) {  // the accumulator is in registers: release it before the arithmetic
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tmem_empty[group]);
                }
                uint32_t w[8];
#pragma unroll
                for (int c = 0; c < 16; c += 2) {
                    float f[2];
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const float u1 = __shfl_down_sync(0xffffffffu, __uint_as_float(v1[c + u]), 1);
                        const float u2 = __shfl_down_sync(0xffffffffu, __uint_as_float(v2[c + u]), 2);
                        f[u] = (__uint_as_float(v0[c + u]) + u1) + (u2 + bias[sub * 16 + c + u]);
                    }
                    __nv_bfloat162 h = __floats2bfloat162_rn(f[0], f[1]);
                    if (relu) h = __hmax2(h, __floats2bfloat162_rn(0.0f, 0.0f));  // max commutes with the rounding
                    w[c >> 1] = *reinterpret_cast<const uint32_t*>(&h);
                }
                if (lane < kTap3Group) {
                    uint8_t* rp = so + r_tile * 128;
                    const int ch = hc * 4 + sub * 2;  // 16-byte chunk index of these 16 channels
                    *reinterpret_cast<uint4*>(rp + ((ch ^ (r_tile & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
                    *reinterpret_cast<uint4*>(rp + (((ch + 1) ^ (r_tile & 7)) << 4)) = make_uint4(w[4], w[5], w[6], w[7]);
                }
            }
            if (hc == 0 && lane < kTap3Group) ro[r_tile] = (row < p.M && qq < p.Wo) ? line * p.Wo + qq : -1;
            row += step;
            line += step_lines;
            qq += step_q;
            if (qq >= p.Wwide) {
                qq -= p.Wwide;
                ++line;
            }
            named_bar_sync(1 + group, 256);      // the group's staged tile is complete
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                const int r = e * 16 + h * 4 + (lane >> 3);
                const int chunk = lane & 7;
                if (r < kTap3Rows) {
                    const int dst_row = ro[r];
                    const uint4 val = *reinterpret_cast<const uint4*>(so + r * 128 + ((chunk ^ (r & 7)) << 4));
                    if (dst_row >= 0) *reinterpret_cast<uint4*>(out + static_cast<size_t>(dst_row) * 128 + chunk * 16) = val;
                }
            }
            named_bar_sync(3 + group, 256);      // copy-out done: the buffer may be overwritten by the group's next tile
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace bv
