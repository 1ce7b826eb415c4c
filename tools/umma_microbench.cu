// Raw tcgen05.mma throughput on one SM: N, number of round-robin accumulators, dependent vs independent issue.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_microbench umma_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../incremental_multimodal_medical_learning_ii_b200/csrc/ptx.cuh"
using namespace bv;

template <int N>
__global__ void __launch_bounds__(128, 1) k(int iters, int nacc, int ablocks, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tptr;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) { tmem_alloc(&tptr, 512); tmem_relinquish(); }
    fence_proxy_async_smem();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tb = tptr;
    if (warp == 0) {
        constexpr uint32_t idesc = umma_idesc_bf16_f32(128, N);
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 128 * 1024);
        long long t0 = 0, t1 = 0;
        if (elect_one()) {
            t0 = clock64();
            for (int i = 0; i < iters; ++i) {
                const uint64_t ad = umma_desc_k_sw128(a0 + (i % ablocks) * 16384);
                const uint64_t bd = umma_desc_k_sw128(b0 + (i & 1) * 32768);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    umma_bf16_ss(tb + ((i * 4 + kk) % nacc) * N, ad + 2 * kk, bd + 2 * kk, idesc, 1u);
            }
            umma_commit(&bar);
        }
        __syncwarp();
        mbar_wait(&bar, 0);
        t1 = clock64();
        if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

template <int N>
void run(int nacc, int ablocks, int grid) {
    long long* d; cudaMalloc(&d, 8 * 256);
    cudaFuncSetAttribute(k<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int iters = 2048;
    k<N><<<grid, 128, 200 * 1024>>>(iters, nacc, ablocks, d);
    k<N><<<grid, 128, 200 * 1024>>>(iters, nacc, ablocks, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[256]; cudaMemcpy(h, d, 8 * grid, cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
    const double cyc = (double)mx / (iters * 4);
    printf("N=%3d nacc=%d ablocks=%d grid=%3d: %6.1f cycles/MMA  -> %5.0f MAC/cycle/SM (%s)\n", N, nacc, ablocks, grid, cyc,
           128.0 * N * 16 / cyc, cudaGetErrorString(e));
    cudaFree(d);
}

int main() {
    for (int grid : {1, 148}) {
        run<64>(1, 1, grid); run<64>(4, 1, grid); run<64>(4, 8, grid);
        run<128>(1, 1, grid); run<128>(2, 1, grid); run<128>(2, 8, grid);
        run<192>(1, 1, grid); run<192>(2, 8, grid);
        run<256>(1, 1, grid); run<256>(2, 8, grid);
    }
    return 0;
}
