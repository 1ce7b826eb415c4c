// TILE variant of the fused 8-bit stem (superseded as the default by the row-streaming kernel in stem_rows.cuh, which is
// 2x faster; kept behind BV_STEM_V1 and bv_stem_u8_nhwc(variant 1) as the second implementation the tests compare with).
// Fused BioViL stem for 8-bit frames: conv 7x7 stride 2 pad 3 (1 folded input channel -> 64) + BatchNorm + ReLU +
// max-pool 3x3 stride 2 pad 1, written straight as the NHWC bf16 input of layer1.  The 240x240x64 conv output
// (7.4 MB per frame in bf16, the largest tensor of the network) never reaches HBM.
//
// Replaces (reference): ResNetHIML.forward conv1 -> bn1 -> relu -> maxpool,
// health_multimodal/image/model/resnet.py:34-37, on frames made by ToTensor + ExpandChannels
// (health_multimodal/image/data/transforms.py:12-38: uint8/255, three identical channels -> one summed channel,
// 1/255 folded into the weights, so the integer pixels are exact bf16 operands).
//
// One tile = 8x8 pooled pixels of one frame = 17x17 conv pixels (289 GEMM rows, three 128-row tcgen05 blocks)
// = a 39x40 patch of the input.  Per tile, 256 threads:
//   1. cp.async the NEXT tile's raw bytes (zero-filled outside the frame) while working on this one
//   2. u8 -> bf16 into a small smem image
//   3. build the im2col A operand in the 128B-swizzled K-major layout tcgen05 reads: row = conv pixel,
//      k = r*8 + s (7 rows x 8 columns of the window; column 7 has zero weights); k = 56..58 hold 1.0 against the
//      BatchNorm bias split into three bf16 terms, so the bias is added by the MMA itself
//   4. one thread issues 3 x 4 tcgen05.mma (128x64x16) into three TMEM accumulators
//   5. all warps: TMEM -> bf16 conv tile in smem (aliasing the dead A operand); no arithmetic
//   6. 3x3/2 max-pool from smem, THEN ReLU (max, ReLU and the bf16 rounding are monotonic, so
//      relu(max(round(x))) == max(round(relu(x))) exactly, on 64 pooled instead of 289 conv pixels),
//      16-byte coalesced stores of the pooled pixels
// (Converting the next tile's pixels while this tile's MMAs run was measured slower: the two extra CTA barriers
// inside the MMA window cost more than the conversion they hide.)
// Persistent CTAs, two per SM (TMEM: 256 columns each) so one CTA's load/convert phases overlap the other's math.
#pragma once
#include "ptx.cuh"

namespace bv {

constexpr int kStemThreads = 256;
constexpr int kStemPool = 8;                       // pooled tile edge
constexpr int kStemConv = 2 * kStemPool + 1;       // 17 conv pixels per edge
constexpr int kStemRows = kStemConv * kStemConv;   // 289 GEMM rows
constexpr int kStemIn = 2 * (kStemConv - 1) + 7;   // 39 input rows / columns touched
constexpr int kStemRawPitch = 48;                  // raw bytes per input row (8-byte aligned superset of 39+1)
constexpr int kStemInPitch = 40;                   // bf16 elements per converted row
constexpr int kStemABytes = 3 * 128 * 128;         // three 128-row blocks of 128-byte rows (also holds the conv tile)
constexpr int kStemBBytes = 64 * 128;
constexpr int kStemRawBytes = 2048;                // >= 39*48, per buffer
constexpr int kStemInBytes = 3200;                 // >= 39*40*2
constexpr int kStemTabBytes = kStemRows * 8 * 4;   // per (row, 16-byte chunk): smem source / destination offsets
constexpr int kStemSmemBytes =
    1024 + kStemABytes + kStemBBytes + 2 * kStemRawBytes + kStemInBytes + kStemTabBytes + 64;
constexpr int kStemSmemRequest = 100 * 1024;       // > 227/3 KB so that at most two CTAs share an SM (TMEM 2 x 256)

struct StemParams {
    CUtensorMap tmW;              // weights [64 out][64 k] bf16, k = r*8+s, 1/255 and BatchNorm folded
    const uint8_t* frames;        // [B][H][W]
    const float* bias;            // [64]
    __nv_bfloat16* out;           // [B][H/4][W/4][64]
    int B, H, W;
    int tiles_x, tiles_y;         // pooled tiles per frame
};

__device__ __forceinline__ void cp_async_8_zfill(void* dst, const void* src, bool valid) {
    const uint32_t n = valid ? 8u : 0u;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(smem_u32(dst)), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

__global__ void __launch_bounds__(kStemThreads, 2) stem_fused_kernel(const __grid_constant__ StemParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];   // keep the shared address space visible (LDS/STS, not LD.E)
    if ((smem_u32(smem) & 1023u) != 0u) __trap();
    uint8_t* smem_a = smem;                                  // A operand, later the conv tile
    uint8_t* smem_b = smem_a + kStemABytes;                  // weights
    uint8_t* raw = smem_b + kStemBBytes;                     // 2 raw input buffers
    __nv_bfloat16* in_s = reinterpret_cast<__nv_bfloat16*>(raw + 2 * kStemRawBytes);
    uint32_t* tab = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(in_s) + kStemInBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(tab) + kStemTabBytes);
    uint64_t* w_bar = bars;
    uint64_t* mma_bar = bars + 1;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2);

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const int Hc = p.H / 2, Wc = p.W / 2, Hp = p.H / 4, Wp = p.W / 4;
    const int tiles_per_frame = p.tiles_x * p.tiles_y;
    const int num_tiles = p.B * tiles_per_frame;

    if (tid == 0) {
        tma_prefetch_desc(&p.tmW);
        mbar_init(w_bar, 1);
        mbar_init(mma_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_ptr, 256);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    if (tid == 0) {
        mbar_arrive_expect_tx(w_bar, kStemBBytes);
        tma_load_2d(&p.tmW, w_bar, smem_b, 0, 0, kEvictLast);
    }

    auto tile_origin = [&](int tile, int& b, int& Y0, int& X0) {
        b = tile / tiles_per_frame;
        const int t = tile - b * tiles_per_frame;
        const int ty = t / p.tiles_x;
        Y0 = ty * kStemPool;
        X0 = (t - ty * p.tiles_x) * kStemPool;
    };
    // raw[buf][row][0..47] = frame[b][4*Y0-5+row][4*X0-8 .. 4*X0+39], zero outside the frame
    auto prefetch = [&](int tile, int buf) {
        if (tile < num_tiles && tid < kStemIn * 6) {
            int b, Y0, X0;
            tile_origin(tile, b, Y0, X0);
            const int row = tid / 6, w8 = tid - row * 6;
            const int iy = 4 * Y0 - 5 + row;
            const int ix = 4 * X0 - 8 + w8 * 8;
            const bool ok = iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
            const uint8_t* src = p.frames + (static_cast<size_t>(b) * p.H + (ok ? iy : 0)) * p.W + (ok ? ix : 0);
            cp_async_8_zfill(raw + buf * kStemRawBytes + row * kStemRawPitch + w8 * 8, src, ok);
        }
        cp_async_commit();
    };

    constexpr uint32_t idesc = umma_idesc_bf16_f32(128, 64);
    const int quarter = warp & 3;   // TMEM lane quarter of this warp
    const int half = warp >> 2;     // which 32 of the 64 output channels this warp converts

    // Tile-independent index tables (the divisions by 17 are done once per CTA, not once per tile):
    //   tab[i] for im2col item i = (row m, chunk r): low 16 bits = byte offset of the 16 source bytes in in_s
    //   (0xFFFF = zero chunk), high 16 bits = byte offset of the destination chunk in the A operand.
    for (int i = tid; i < kStemRows * 8; i += kStemThreads) {
        const int m = i >> 3, r = i & 7;
        const int cy = m / kStemConv, cx = m - cy * kStemConv;
        const uint32_t src = (r < 7) ? static_cast<uint32_t>(((2 * cy + r) * kStemInPitch + 2 * cx) * 2) : 0xFFFFu;
        const uint32_t dst = static_cast<uint32_t>(m * 128 + ((r ^ (m & 7)) << 4));
        tab[i] = src | (dst << 16);
    }
    // conv-tile coordinates of the three rows this thread converts in the epilogue (packed cy | cx << 8)
    uint32_t epi_yx[3];
#pragma unroll
    for (int blk = 0; blk < 3; ++blk) {
        const int m = blk * 128 + quarter * 32 + lane;
        const int cy = m / kStemConv, cx = m - cy * kStemConv;
        epi_yx[blk] = static_cast<uint32_t>(cy) | (static_cast<uint32_t>(cx) << 8);
    }
    // the two pooled (pixel, 8-channel chunk) items of this thread
    const int pool_c = tid & 7;
    int pool_py[2], pool_px[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int pp = (tid + j * kStemThreads) >> 3;
        pool_py[j] = pp / kStemPool;
        pool_px[j] = pp - pool_py[j] * kStemPool;
    }
    static_assert(kStemPool * kStemPool * 8 == 2 * kStemThreads, "two pooled items per thread");

    prefetch(blockIdx.x, 0);
    mbar_wait(w_bar, 0);
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        int b, Y0, X0;
        tile_origin(tile, b, Y0, X0);
        prefetch(tile + gridDim.x, buf ^ 1);
        cp_async_wait<1>();   // this tile's bytes have landed (the group just committed may still be in flight)
        __syncthreads();

        // ---- 2. u8 -> bf16, dropping the 3 alignment bytes: in_s[row][j] = raw[row][j+3] ----
        {
            const uint8_t* rb = raw + buf * kStemRawBytes;
            for (int i = tid; i < kStemIn * (kStemInPitch / 2); i += kStemThreads) {
                const int row = i / (kStemInPitch / 2);
                const int j = (i - row * (kStemInPitch / 2)) * 2;
                const float a = static_cast<float>(rb[row * kStemRawPitch + j + 3]);
                const float c = static_cast<float>(rb[row * kStemRawPitch + j + 4]);
                *reinterpret_cast<__nv_bfloat162*>(in_s + row * kStemInPitch + j) = __floats2bfloat162_rn(a, c);
            }
        }
        __syncthreads();

        // ---- 3. im2col rows: A[m][r*8 .. r*8+7] = in_s[2*cy + r][2*cx .. 2*cx+7], chunk 7 = 0 ----
#pragma unroll 2
        for (int i = tid; i < kStemRows * 8; i += kStemThreads) {
            const uint32_t e = tab[i];
            uint4 v = make_uint4(0x3F803F80u, 0x00003F80u, 0u, 0u);   // chunk 7: k = 56..58 = 1.0 (bias columns)
            if ((e & 0xFFFFu) != 0xFFFFu) {
                const uint32_t* src =
                    reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(in_s) + (e & 0xFFFFu));
                v = make_uint4(src[0], src[1], src[2], src[3]);
            }
            *reinterpret_cast<uint4*>(smem_a + (e >> 16)) = v;
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncthreads();

        // ---- 4. MMA: three 128-row blocks, K = 64 ----
        if (warp == 0 && elect_one()) {  // warp-uniform branch + election keeps the MMA operands in uniform registers
            tc_fence_after();
            const uint64_t bdesc = umma_desc_k_sw128(smem_u32(smem_b));
            // k outer, block inner: consecutive MMAs go to different accumulators (same-accumulator MMAs serialise)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
#pragma unroll
                for (int blk = 0; blk < 3; ++blk) {
                    const uint64_t adesc = umma_desc_k_sw128(smem_u32(smem_a + blk * 128 * 128));
                    umma_bf16_ss(tmem_base + static_cast<uint32_t>(blk * 64), adesc + static_cast<uint64_t>(2 * k),
                                 bdesc + static_cast<uint64_t>(2 * k), idesc, k != 0 ? 1u : 0u);
                }
            }
            umma_commit(mma_bar);
        }
        mbar_wait(mma_bar, it & 1u);
        tc_fence_after();

        // ---- 5. TMEM (conv + bias) -> bf16 conv tile (overwrites A: the MMAs have finished reading it) ----
#pragma unroll
        for (int blk = 0; blk < 3; ++blk) {
            const int m = blk * 128 + quarter * 32 + lane;
            uint32_t v[32];
            tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                              static_cast<uint32_t>(blk * 64 + half * 32), v);
            tmem_ld_wait();
            if (m < kStemRows) {
                const int cy = static_cast<int>(epi_yx[blk] & 0xFFu), cx = static_cast<int>(epi_yx[blk] >> 8);
                const int yc = 2 * Y0 - 1 + cy, xc = 2 * X0 - 1 + cx;
                // conv pixels outside the conv image are the max-pool's padding: -inf never wins (every window holds
                // a real pixel), which is the reference's padding semantics
                const bool inside = yc >= 0 && yc < Hc && xc >= 0 && xc < Wc;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint32_t w[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const __nv_bfloat162 h2 =
                            __floats2bfloat162_rn(__uint_as_float(v[8 * j + 2 * e]), __uint_as_float(v[8 * j + 2 * e + 1]));
                        w[e] = inside ? *reinterpret_cast<const uint32_t*>(&h2) : 0xFF80FF80u;
                    }
                    *reinterpret_cast<uint4*>(smem_a + m * 128 + (((half * 4 + j) ^ (m & 7)) << 4)) =
                        make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
        }
        tc_fence_before();
        __syncthreads();

        // ---- 6. max-pool 3x3 stride 2 over the conv tile, 16-byte stores ----
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int c = pool_c, py = pool_py[j], px = pool_px[j];
            __nv_bfloat162 mx[4];
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    const int m = (2 * py + dy) * kStemConv + 2 * px + dx;
                    const uint4 v = *reinterpret_cast<const uint4*>(smem_a + m * 128 + ((c ^ (m & 7)) << 4));
                    const __nv_bfloat162* hv = reinterpret_cast<const __nv_bfloat162*>(&v);
                    if (dy == 0 && dx == 0) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) mx[e] = hv[e];
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e) mx[e] = __hmax2(mx[e], hv[e]);
                    }
                }
            }
            const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.0f, 0.0f);
#pragma unroll
            for (int e = 0; e < 4; ++e) mx[e] = __hmax2(mx[e], zero2);   // ReLU after the pool
            uint4 o;
            o.x = *reinterpret_cast<uint32_t*>(&mx[0]);
            o.y = *reinterpret_cast<uint32_t*>(&mx[1]);
            o.z = *reinterpret_cast<uint32_t*>(&mx[2]);
            o.w = *reinterpret_cast<uint32_t*>(&mx[3]);
            __nv_bfloat16* dst = p.out + ((static_cast<size_t>(b) * Hp + Y0 + py) * Wp + X0 + px) * 64 + c * 8;
            *reinterpret_cast<uint4*>(dst) = o;
        }
        __syncthreads();   // the conv tile is overwritten by the next tile's im2col rows
    }
    cp_async_wait<0>();

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}

}  // namespace bv
