// CTA-pair building blocks (tcgen05 cta_group::2) and a minimal pair GEMM used to validate them.
//
// Two CTAs of a cluster (the two SMs of a TPC) execute ONE tcgen05.mma of M = 256: each CTA supplies its own 128 rows
// of A and HALF of the B tile (N/2 rows of the K-major weight matrix) from its own shared memory and receives its 128
// rows of D in its own TMEM.  The point for this code base is shared-memory capacity and L2->SM traffic: resident
// weights cost half as much per SM.
//
// Protocol (rank 0 = leader):
//   * TMA loads are issued by BOTH CTAs with .cta_group::2 and the mbarrier address of the LEADER's full barrier (peer
//     bit cleared), so one barrier collects the bytes of both CTAs; the leader's producer posts expect_tx for both.
//   * only the leader's MMA thread issues tcgen05.mma.cta_group::2; tcgen05.commit.cta_group::2 with multicast mask
//     0b11 arrives on the barrier at the same offset in both CTAs (stage-empty, accumulator-full).
//   * "accumulator drained" goes the other way: the epilogue warps of both CTAs arrive on the LEADER's barrier
//     (mbarrier.arrive.shared::cluster on the peer-bit-cleared address).
//   * TMEM is allocated with tcgen05.alloc.cta_group::2 by the same warp of both CTAs; a cluster barrier separates
//     barrier initialisation from first use and the last use from deallocation.
#pragma once
#include "ptx.cuh"

namespace bv {

constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address (rank 0 of the pair)

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// address of a local shared-memory object in the shared::cluster window of THIS CTA
__device__ __forceinline__ uint32_t cluster_addr_of(const void* p, uint32_t rank) {
    uint32_t a;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(a) : "r"(smem_u32(p)), "r"(rank));
    return a;
}
// arrive on the barrier at the same offset in the LEADER CTA (works from either CTA)
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
    const uint32_t a = cluster_addr_of(bar, 0);
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];\n" ::"r"(a) : "memory");
}

__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(dst_smem)),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols) : "memory");
}

// D[tmem, both CTAs] (+)= A[smem, per CTA] * B[smem, half per CTA]^T; issued by ONE thread of the leader CTA.
__device__ __forceinline__ void umma_bf16_ss_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                  uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives on the barrier at this offset in BOTH CTAs once all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(
            smem_u32(bar)),
        "h"(static_cast<uint16_t>(3))
        : "memory");
}

// TMA loads whose completion bytes are credited to the LEADER's barrier at the same offset as `bar`
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1,
                                                 uint64_t policy) {
    const uint32_t bar_addr = cluster_addr_of(bar, cluster_ctarank()) & kPeerBitMask;
    const uint32_t dst_addr = cluster_addr_of(dst, cluster_ctarank());
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4}], [%2], %5;\n" ::"r"(dst_addr),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_addr), "r"(c0), "r"(c1), "l"(policy)
        : "memory");
}

__device__ __forceinline__ void mbar_arrive_leader_release(uint64_t* bar) {
    const uint32_t a = cluster_addr_of(bar, 0);
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(a) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_try_wait_cluster_hint(uint64_t* bar, uint32_t parity) {
    if constexpr (kMbarHintNs == 0) return mbar_try_wait_cluster(bar, parity);
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(kMbarHintNs)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait_cluster(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait_cluster_hint(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;\n" ::: "memory"); }

// im2col-mode load whose completion bytes are credited to the LEADER's barrier (pair MMAs consume it)
__device__ __forceinline__ void tma_load_im2col_4d_pair(const CUtensorMap* m, uint64_t* bar, void* dst, int c, int w,
                                                        int h, int n, uint16_t off_w, uint16_t off_h, uint64_t policy) {
    const uint32_t bar_addr = smem_u32(bar) & kPeerBitMask;
    asm volatile(
        "cp.async.bulk.tensor.4d.im2col.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8}, %9;\n" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_addr), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h),
        "l"(policy)
        : "memory");
}

// ------------------------------------------------------------------------------------------------------------
// Minimal pair GEMM: out[M, N] (fp32) = A[M, K] * W[N, K]^T, bf16 operands.  M is tiled in 256-row pair tiles
// (128 rows per CTA), the whole N (<= 256, multiple of 32) is one tile, K in 64-wide blocks.  Validation vehicle.
// ------------------------------------------------------------------------------------------------------------
struct PairGemmParams {
    CUtensorMap tmA;  // [M, K], box 64 x 128
    CUtensorMap tmB;  // [N, K], box 64 x N/2
    float* out;
    int M, N, K;
    int num_pair_tiles;
};

constexpr int kPairStages = 4;
constexpr int kPairThreads = 192;  // warp 0 producer, warp 1 MMA, warps 2..5 epilogue
constexpr int kPairStageBytes = kPairStages * 0 + 16384 + 16384;  // A tile + B half tile (N/2 <= 128 rows)
constexpr int kPairSmemBytes = kPairStages * kPairStageBytes + (2 * kPairStages + 4) * 8 + 16;

__global__ void __launch_bounds__(kPairThreads, 1) pair_gemm_kernel(const __grid_constant__ PairGemmParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0u) __trap();
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kPairStages * kPairStageBytes);
    uint64_t* full_bar = bars;                        // used in the leader only
    uint64_t* empty_bar = bars + kPairStages;         // per CTA (multicast commit)
    uint64_t* tmem_full = bars + 2 * kPairStages;     // per CTA (multicast commit), 2 accumulators
    uint64_t* tmem_empty = tmem_full + 2;             // used in the leader only: 8 epilogue warps of the pair
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1;
    const int num_pairs = gridDim.x >> 1;
    const int kblocks = p.K / 64;
    const int half_n = p.N / 2;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tmA);
        tma_prefetch_desc(&p.tmB);
        for (int i = 0; i < kPairStages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 8);
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

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        int stage = 0;
        uint32_t phase = 0;
        for (int t = pair; t < p.num_pair_tiles; t += num_pairs) {
            const int m0 = t * 256 + static_cast<int>(rank) * 128;
            for (int kb = 0; kb < kblocks; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1u);
                if (elect_one()) {
                    uint8_t* dst = smem + stage * kPairStageBytes;
                    // the leader posts the bytes of BOTH CTAs on its full barrier
                    if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2u * (16384u + static_cast<uint32_t>(half_n) * 128u));
                    tma_load_2d_pair(&p.tmA, &full_bar[stage], dst, kb * 64, m0, kEvictNormal);
                    tma_load_2d_pair(&p.tmB, &full_bar[stage], dst + 16384, kb * 64, static_cast<int>(rank) * half_n,
                                     kEvictLast);
                }
                __syncwarp();
                if (++stage == kPairStages) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (rank == 0) {
            const uint32_t idesc = umma_idesc_bf16_f32(256, p.N);
            const uint32_t base = smem_u32(smem);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int t = pair; t < p.num_pair_tiles; t += num_pairs, ++it) {
                const int acc = it & 1;
                mbar_wait(&tmem_empty[acc], ((it >> 1) & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * 256);
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t adesc = umma_desc_k_sw128(base + static_cast<uint32_t>(stage * kPairStageBytes));
                        const uint64_t bdesc = umma_desc_k_sw128(base + static_cast<uint32_t>(stage * kPairStageBytes + 16384));
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16_ss_pair(d_tmem, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k),
                                              idesc, (kb != 0 || k != 0) ? 1u : 0u);
                        umma_commit_pair(&empty_bar[stage]);
                        if (kb == kblocks - 1) umma_commit_pair(&tmem_full[acc]);
                    }
                    __syncwarp();
                    if (++stage == kPairStages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else {
        // ===================== epilogue (warps 2..5, both CTAs): TMEM -> fp32 rows =====================
        const int quarter = warp & 3;
        int it = 0;
        for (int t = pair; t < p.num_pair_tiles; t += num_pairs, ++it) {
            const int acc = it & 1;
            mbar_wait(&tmem_full[acc], (it >> 1) & 1u);
            tc_fence_after();
            const int row = t * 256 + static_cast<int>(rank) * 128 + quarter * 32 + lane;
            for (int c = 0; c < p.N; c += 32) {
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(acc * 256 + c), v);
                tmem_ld_wait();
                if (row < p.M) {
                    float4* op = reinterpret_cast<float4*>(p.out + static_cast<size_t>(row) * p.N + c);
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        op[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                            __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&tmem_empty[acc]);
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
