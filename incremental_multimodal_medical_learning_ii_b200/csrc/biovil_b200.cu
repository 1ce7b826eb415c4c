// Host side of the C ABI declared in include/biovil_b200.h: weight table, per-shape launch plan (activation
// buffers carved from the caller's workspace, TMA tensor maps, kernel parameters) and the launches.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "../../include/biovil_b200.h"
#include "aux_kernels.cuh"
#include "chain_gemm.cuh"
#include "conv3x3_tap3.cuh"
#include "conv_gemm.cuh"
#include "jpeg_stage.h"
#include "l1_block.cuh"
#include "pair_chain.cuh"
#include "pair_gemm.cuh"
#include "resize.cuh"
#include "stem_fused.cuh"
#include "stem_rows.cuh"

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define BV_CUDA(expr)                                                                                   \
    do {                                                                                                \
        cudaError_t e__ = (expr);                                                                       \
        if (e__ != cudaSuccess)                                                                         \
            return fail(BV_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, \
                        __LINE__);                                                                      \
    } while (0)

// ---- driver entry points for tensor-map encoding (resolved lazily so the library loads without libcuda) ----
PFN_cuTensorMapEncodeTiled_v12000 g_encode_tiled = nullptr;
PFN_cuTensorMapEncodeIm2col_v12000 g_encode_im2col = nullptr;

int resolve_driver() {
    if (g_encode_tiled && g_encode_im2col) return BV_OK;
    cudaDriverEntryPointQueryResult qres;
    void* fn = nullptr;
    BV_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) return fail(BV_ERR_CUDA, "cuTensorMapEncodeTiled not available");
    g_encode_tiled = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    fn = nullptr;
    BV_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) return fail(BV_ERR_CUDA, "cuTensorMapEncodeIm2col not available");
    g_encode_im2col = reinterpret_cast<PFN_cuTensorMapEncodeIm2col_v12000>(fn);
    return BV_OK;
}

int make_tmap_2d(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t outer, uint32_t box_inner,
                 uint32_t box_outer) {
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {inner * 2};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box,
                                estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail(BV_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) inner=%llu outer=%llu box=%ux%u", (int)r,
                    (unsigned long long)inner, (unsigned long long)outer, box_inner, box_outer);
    return BV_OK;
}

// NHWC tensor viewed as (C, W, lines = N*H), tiled mode, box = 64 channels x `px` pixels x 1 line, 128B swizzle.
int make_tmap_lines(CUtensorMap* tm, const void* base, int lines, int W, int C, int px) {
    cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)lines};
    cuuint64_t strides[2] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)px, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = g_encode_tiled(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail(BV_ERR_CUDA, "cuTensorMapEncodeTiled(3d) failed (%d) lines=%d W=%d C=%d", (int)r, lines, W, C);
    return BV_OK;
}

// NHWC activation tensor [N][H][W][C] seen by TMA as (C, W, H, N); 128 output pixels x 64 channels per load.
int make_tmap_im2col(CUtensorMap* tm, const void* base, int N, int H, int W, int C, int R, int S, int stride,
                     int pad, bool wide = false, int pixels = 0, bool widen_right = false) {
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    int lower[2] = {-pad, -pad};
    int upper[2] = {pad - (S - 1), pad - (R - 1)};
    if (wide) upper[0] = pad;  // window origins -pad .. W-1+pad: the zero-padded image row, linear in memory order
    if (widen_right) {         // 1x1 view of an OUTPUT-sized tensor in the same widened row space: pixels 0 .. W+1 per row
        lower[0] = 0;
        upper[0] = 2;
    }
    cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
    CUresult r = g_encode_im2col(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides,
                                 lower, upper, bv::kBlockK, pixels > 0 ? pixels : (wide ? bv::kWideRows : bv::kBlockM), estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail(BV_ERR_CUDA, "cuTensorMapEncodeIm2col failed (%d) N=%d H=%d W=%d C=%d R=%d S=%d stride=%d pad=%d",
                    (int)r, N, H, W, C, R, S, stride, pad);
    return BV_OK;
}

bool env_flag(const char* name) {
    const char* v = getenv(name);
    return v && v[0] && v[0] != '0';
}

// Every kernel of the forward goes through here: optional thread-block cluster and - opt-in, BV_PDL=1 - programmatic
// dependent launch (ptx.cuh: pdl_launch_dependents / pdl_wait; the next kernel's prologue would overlap this kernel's
// tail).  PDL is OFF by default: same-box A/B showed no gain at batch 512 (21.28k / 21.51k img/s with, 21.37k without:
// one CTA per SM with ~200 KB of shared memory leaves nothing to overlap but the prologue), and a stress loop of 5000
// forwards (tools/stress_forward.py) hit an "unspecified launch failure" twice with it on and never with it off.
bool pdl_enabled() {
    static const bool on = env_flag("BV_PDL");
    return on;
}
template <typename... KArgs, typename... Args>
cudaError_t launch_ex(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, int cluster,
                      Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid, 1, 1);
    cfg.blockDim = dim3(block, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[2];
    int n = 0;
    if (cluster > 1) {
        at[n].id = cudaLaunchAttributeClusterDimension;
        at[n].val.clusterDim.x = cluster;
        at[n].val.clusterDim.y = 1;
        at[n].val.clusterDim.z = 1;
        ++n;
    }
    if (pdl_enabled()) {
        at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    cfg.attrs = at;
    cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// Row-streaming stem: 8 epilogue warps of 32 channels (0.566 ms per 512 frames alone); BV_SR_CG4 = 16 warps of 16
// channels (0.594 ms: the kernel is bound by instruction throughput, not latency, and more warps add per-warp overhead).
// With p.out1 set the kernel also applies layer1.0's conv1 (1x1, 64 -> 64, + bn1 + ReLU) to every pooled pixel.
void launch_stem_rows(const bv::StemRowsParams& p, int grid, cudaStream_t st) {
    const bool c1 = p.out1 != nullptr;
    if (env_flag("BV_SR_CG4")) {
        if (c1) launch_ex(bv::stem_rows_kernel<4, true>, grid, bv::sr_threads(4), bv::sr_smem_bytes(true), st, 1, p);
        else launch_ex(bv::stem_rows_kernel<4, false>, grid, bv::sr_threads(4), bv::sr_smem_bytes(false), st, 1, p);
    } else {
        if (c1) launch_ex(bv::stem_rows_kernel<2, true>, grid, bv::sr_threads(2), bv::sr_smem_bytes(true), st, 1, p);
        else launch_ex(bv::stem_rows_kernel<2, false>, grid, bv::sr_threads(2), bv::sr_smem_bytes(false), st, 1, p);
    }
}

struct ConvOperand {
    const void* x;  // NHWC bf16 [B][H][W][cin]
    int H, W;
    bv_conv c;
};

// Kernel configurations <BN, STAGES, NBUF> (see ConvGemmCfg): picked per layer by arithmetic intensity.
enum ConvCfg { kCfg256Deep = 0, kCfg256Res, kCfg128Res, kCfg128Deep, kCfg64, kCfg64BRes, kCfg64Wide, kCfg64Tap3, kCfg128M2, kCfg128TR,
               kCfg256PairDeep, kCfg256PairRes, kCfg256PairDeepE16, kCfg256PairRes37, kNumCfg };
// kCfg64Tap3 is a different kernel (conv3x3_tap3.cuh): the three horizontal taps of a filter row in one N = 192 MMA
// kCfg256Pair*: the 256-wide tile on CTA PAIRS (cta_group::2, M = 256 over two SMs, half of every weight stage per CTA):
// 32 KB instead of 48 KB per ring stage -> 6 stages (long K) or 4 stages + 6 staging tiles (residual stream)
#define BV_FOR_EACH_CFG(X)                                                                                        \
    X(kCfg256Deep, 256, 4, 2, false, false, 1, 8, false, false) X(kCfg256Res, 256, 3, 5, false, false, 1, 16, false, false)   \
    X(kCfg128Res, 128, 3, 7, false, false, 1, 16, false, false) X(kCfg128Deep, 128, 6, 2, false, false, 1, 16, false, false)  \
    X(kCfg64, 64, 8, 2, false, false, 1, 16, false, false) X(kCfg64BRes, 64, 6, 2, true, false, 1, 16, false, false)          \
    X(kCfg64Wide, 64, 6, 2, true, true, 1, 16, false, false) X(kCfg128M2, 128, 4, 2, false, false, 2, 16, false, false)       \
    X(kCfg128TR, 128, 4, 2, false, false, 2, 16, true, false)                                                                \
    X(kCfg256PairDeep, 256, 6, 2, false, false, 1, 8, false, true) X(kCfg256PairRes, 256, 4, 6, false, false, 1, 16, false, true)   \
    X(kCfg256PairDeepE16, 256, 6, 2, false, false, 1, 16, false, true) X(kCfg256PairRes37, 256, 3, 7, false, false, 1, 16, false, true)
const int kCfgBN[kNumCfg] = {256, 256, 128, 128, 64, 64, 64, 64, 128, 128, 256, 256, 256, 256};
const int kCfgMT[kNumCfg] = {1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 1, 1, 1, 1};
const bool kCfgPair[kNumCfg] = {false, false, false, false, false, false, false, false, false, false, true, true, true, true};

struct ConvLaunch {
    bv::ConvGemmParams p;
    int cfg;
    int bn;
    int grid;
};

// Chained conv3 -> next conv1 kernel: <N2, STAGES, NB1> per width of the second GEMM (see ChainCfg).
// cfg 3: the downsample tail streams two A k-blocks per chunk and carries no residual -> deeper ring, fewer staging tiles
// cfg 4: four chunks per tile (N1 = 512) want staging depth more than ring depth (experiment)
#define BV_FOR_EACH_CHAIN(X) X(0, 64, 3, 7) X(1, 128, 3, 6) X(2, 256, 3, 4) X(3, 64, 4, 5) X(4, 128, 2, 8)

struct ChainLaunch {
    bv::ChainParams p;
    int n2;
    int cfg;
    int grid;
    int k1;  // total K of the first GEMM (for cost accounting)
};

const int kL1ShiftedDefault = 1;   // shifted-tap 3x3 GEMM in the 64-wide forms of the layer1 block kernel (BV_L1_SH)
const int kL1LastBlockDefault = 1; // the layer's last block (128-wide successor) on the block kernel (BV_L1_LAST)

// Fused layer1 block on CTA pairs (l1_block.cuh): conv2 3x3 + conv3 + identity + next conv1.
struct L1Launch {
    bv::L1BlockParams p;
    int grid;
    bool ds;   // first block of the layer: downsample branch as a second K segment instead of an identity tensor
    bool sh;   // shifted-tap form of the 3x3 GEMM (nine N = 64 MMAs into a double-buffered 64-column accumulator)
    int n2;    // width of the chained next conv1: 64, or 128 (the layer's last block; shifted taps only)
};

// Chained conv3 + identity -> next conv1 on CTA pairs (pair_chain.cuh): <N2, KB1, STAGES, NSTG> per shape class.
// (measured, layer3 class: 5 stages + 5 sub-tiles 0.66 ms, 4 + 6 0.61 ms, 3 + 7 0.69 ms per launch - profiles/r2g_*)
#define BV_FOR_EACH_PAIR_CHAIN(X) X(0, 256, 4, 4, 6) X(1, 256, 2, 6, 6) X(2, 128, 2, 5, 7) X(3, 256, 4, 5, 5) X(4, 256, 4, 3, 7)
struct PairChainLaunch {
    bv::PairChainParams p;
    int cfg;
    int grid;
    int n2, k1;
};

// One step of the forward plan: a single fused convolution, a chained pair, a fused layer1 block, or a pair-chained tail.
struct PlanStep {
    bool chain = false;
    bool l1 = false;
    bool pc = false;
    ConvLaunch conv;
    ChainLaunch ch;
    L1Launch l1b;
    PairChainLaunch pch;
};

// Per-device state: the dynamic-shared-memory opt-ins (cudaFuncSetAttribute) apply to the device that is current when
// they are made, the SM count and the cluster-launch capability differ per device, and one process may drive several
// devices (an ImageModel on cuda:0 and a ZeroShotScorer on cuda:1).  Every entry point calls device_setup(), which
// (re)binds the calling thread to the state of ITS current device.
constexpr int kMaxDevices = 64;
struct DeviceState {
    bool ready = false;
    int num_sms = 0;
    int l1_launchable = -1;        // -1 = not probed yet
    int pair_launchable = -1;
    long long* dbg = nullptr;      // BV_TIMING builds: per-CTA wait-cycle counters of the most recent launch
};
DeviceState g_dev[kMaxDevices];
std::mutex g_dev_mutex;
thread_local DeviceState* t_dev = nullptr;   // state of the calling thread's current device (set by device_setup)
thread_local int g_num_sms = 0;              // == t_dev->num_sms

int set_kernel_attributes() {
#define BV_SET_ATTR(id, BN, ST, NB, BR, WD, MT, EP, TR, PR)                                         \
    BV_CUDA(cudaFuncSetAttribute(bv::conv_gemm_kernel<BN, ST, NB, BR, WD, MT, EP, TR, PR>,          \
                                 cudaFuncAttributeMaxDynamicSharedMemorySize,                       \
                                 bv::ConvGemmCfg<BN, ST, NB, BR, WD, MT, EP, TR, PR>::kSmemBytes));
    BV_FOR_EACH_CFG(BV_SET_ATTR)
#undef BV_SET_ATTR
#define BV_SET_CHAIN_ATTR(id, N2, ST, NB)                                                           \
    BV_CUDA(cudaFuncSetAttribute(bv::chain_gemm_kernel<N2, ST, NB>,                                 \
                                 cudaFuncAttributeMaxDynamicSharedMemorySize,                       \
                                 bv::ChainCfg<N2, ST, NB>::kSmemBytes));
    BV_FOR_EACH_CHAIN(BV_SET_CHAIN_ATTR)
#undef BV_SET_CHAIN_ATTR
    BV_CUDA(cudaFuncSetAttribute(bv::conv3x3_tap3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 bv::kTap3SmemBytes));
    BV_CUDA(cudaFuncSetAttribute(bv::l1_block_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 bv::L1Cfg<64>::kSmemBytes));
    BV_CUDA(cudaFuncSetAttribute(bv::l1_block_kernel<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 bv::L1Cfg<64, true>::kSmemBytes));
    BV_CUDA(cudaFuncSetAttribute(bv::l1_block_kernel<64, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 bv::L1Cfg<64, false, true>::kSmemBytes));
    BV_CUDA(cudaFuncSetAttribute(bv::l1_block_kernel<64, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 bv::L1Cfg<64, true, true>::kSmemBytes));
    BV_CUDA(cudaFuncSetAttribute(bv::l1_block_kernel<128, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 bv::L1Cfg<128, false, true>::kSmemBytes));
    BV_CUDA(cudaFuncSetAttribute(bv::head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bv::kHeadSmemBytes));
    BV_CUDA(cudaFuncSetAttribute(bv::stem_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 bv::kStemSmemRequest));
    BV_CUDA(cudaFuncSetAttribute(bv::stem_rows_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 bv::sr_smem_bytes(false)));
    BV_CUDA(cudaFuncSetAttribute(bv::stem_rows_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 bv::sr_smem_bytes(false)));
    BV_CUDA(cudaFuncSetAttribute(bv::stem_rows_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 bv::sr_smem_bytes(true)));
    BV_CUDA(cudaFuncSetAttribute(bv::stem_rows_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 bv::sr_smem_bytes(true)));
    BV_CUDA(cudaFuncSetAttribute(bv::pair_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bv::kPairSmemBytes));
#define BV_SET_PC_ATTR(id, N2, KB, ST, NS)                                                          \
    BV_CUDA(cudaFuncSetAttribute(bv::pair_chain_kernel<N2, KB, ST, NS>,                             \
                                 cudaFuncAttributeMaxDynamicSharedMemorySize,                       \
                                 bv::PairChainCfg<N2, KB, ST, NS>::kSmemBytes));
    BV_FOR_EACH_PAIR_CHAIN(BV_SET_PC_ATTR)
#undef BV_SET_PC_ATTR
    return BV_OK;
}

int device_setup() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        cudaGetLastError();
        return fail(BV_ERR_NO_DEVICE, "no CUDA device available (this library has no CPU fallback)");
    }
    if (dev < 0 || dev >= kMaxDevices) return fail(BV_ERR_INVALID, "device index %d out of range", dev);
    DeviceState* d = &g_dev[dev];
    if (!d->ready) {
        std::lock_guard<std::mutex> lock(g_dev_mutex);
        if (!d->ready) {
            cudaDeviceProp prop;
            BV_CUDA(cudaGetDeviceProperties(&prop, dev));
            if (prop.major != 10)
                return fail(BV_ERR_NO_DEVICE, "device %d (%s) is sm_%d%d; this library is built for sm_100a only", dev,
                            prop.name, prop.major, prop.minor);
            int rc = resolve_driver();
            if (rc) return rc;
            if ((rc = set_kernel_attributes())) return rc;
            d->num_sms = prop.multiProcessorCount;
            d->ready = true;
        }
    }
    t_dev = d;
    g_num_sms = d->num_sms;
    return BV_OK;
}

int conv_out_dim(int in, int k, int stride, int pad) { return (in + 2 * pad - k) / stride + 1; }

// Can this device co-schedule CTA pairs of the 256-wide convolution kernel (cluster launch with ~225 KB of shared memory
// per CTA)?  Queried once per device; when it cannot (MIG slice, odd SM count per TPC) the single-CTA form is used.
bool pair_launchable() {
    if (!t_dev) return false;
    int& cached = t_dev->pair_launchable;
    if (cached >= 0) return cached == 1;
    auto kernel = bv::conv_gemm_kernel<256, 6, 2, false, false, 1, 8, false, true>;
    using Cfg = bv::ConvGemmCfg<256, 6, 2, false, false, 1, 8, false, true>;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2, 1, 1);
    cfg.blockDim = dim3(Cfg::kThreads, 1, 1);
    cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int clusters = 0;
    const cudaError_t e = cudaOccupancyMaxActiveClusters(&clusters, kernel, &cfg);
    if (e != cudaSuccess) cudaGetLastError();
    cached = (e == cudaSuccess && clusters >= 1) ? 1 : 0;
    return cached == 1;
}

// Build the kernel parameters (tensor maps included) for one fused convolution.
// fp32 bias vectors on the HOST (kernels that take their biases by value, i.e. in the constant bank): the handle copies every
// convolution's bias once at bv_create - outside any stream capture, where a synchronous copy would be illegal - and the
// plan builder looks them up by device pointer; the unit-test entry points have no handle and copy here.
using HostBiasMap = std::map<const float*, std::vector<float>>;
thread_local const HostBiasMap* t_bias_cache = nullptr;   // the cache of the handle whose plan is being built (set by build_plan)
int fetch_bias(const HostBiasMap* cache, const float* dev, int n, float* dst) {
    if (!dev) return fail(BV_ERR_INVALID, "convolution without a bias vector");
    if (!cache) cache = t_bias_cache;
    if (cache) {
        auto it = cache->find(dev);
        if (it != cache->end() && static_cast<int>(it->second.size()) >= n) {
            memcpy(dst, it->second.data(), sizeof(float) * n);
            return BV_OK;
        }
    }
    BV_CUDA(cudaMemcpy(dst, dev, sizeof(float) * n, cudaMemcpyDeviceToHost));
    return BV_OK;
}

// sum of the operands' bias vectors (the fused downsample branch adds its own), as the epilogues used to add them on the device
int fetch_summed_bias(const ConvOperand* ops, int nops, int n, float* dst) {
    int rc = fetch_bias(nullptr, ops[0].c.bias, n, dst);
    if (rc) return rc;
    for (int i = 1; i < nops; ++i) {
        std::vector<float> b(n);
        if ((rc = fetch_bias(nullptr, ops[i].c.bias, n, b.data()))) return rc;
        for (int j = 0; j < n; ++j) dst[j] += b[j];
    }
    return BV_OK;
}

int build_conv(ConvLaunch* L, int B, const ConvOperand* ops, int nops, const void* residual, int relu, void* out,
               int out_fp32) {
    if (nops < 1 || nops > 2) return fail(BV_ERR_INVALID, "conv needs 1 or 2 operand pairs");
    memset(&L->p, 0, sizeof(L->p));
    bv::ConvGemmParams& p = L->p;
    const bv_conv& c0 = ops[0].c;
    const int Ho = conv_out_dim(ops[0].H, c0.r, c0.stride, c0.pad);
    const int Wo = conv_out_dim(ops[0].W, c0.s, c0.stride, c0.pad);
    const int N = c0.cout;
    if (N % 64 != 0) return fail(BV_ERR_INVALID, "cout=%d must be a multiple of 64", N);
    int total_kblocks = 0;
    for (int i = 0; i < nops; ++i) total_kblocks += ops[i].c.r * ops[i].c.s * (ops[i].c.cin / 64);
    // Compute-bound shapes (long K loop) take the 256-wide tile; short-K shapes are HBM-bound and take the
    // 128-wide tile with either a deep A ring (no residual) or many staging tiles (residual stream).
    int cfg;
    const bool res_stream = residual && !out_fp32;
    if (N % 128 != 0) cfg = (N == 64 && total_kblocks <= bv::kMaxResidentKB && !env_flag("BV_NO_BRES")) ? kCfg64BRes : kCfg64;
    else if (N % 256 != 0) cfg = kCfg128Deep;
    else if (total_kblocks >= 8) cfg = res_stream ? kCfg256Res : kCfg256Deep;
    // short K: HBM-bound, wants staging depth.  From four k-blocks on, the A/B re-reads of 128-wide tiles saturate the
    // L2 -> SM path before HBM does (layer3 conv3: 160 KB of operand + residual loads per 32 KB of output), and the
    // 256-wide tile, which moves 20 % less, is faster (0.478 vs 0.525 ms, profiles/r1w notes).
    else cfg = res_stream ? (total_kblocks >= 4 ? kCfg256Res : kCfg128Res) : kCfg256Res;
    const bool wide_ok = nops == 1 && c0.r == 3 && c0.s == 3 && c0.stride == 1 && c0.pad == 1 && !residual &&
                         !out_fp32 && N == 64 && total_kblocks <= bv::kMaxResidentKB;
    if (cfg == kCfg64BRes && wide_ok && !env_flag("BV_NO_WIDE")) cfg = kCfg64Wide;
    // long-K 128-wide layers without residual (layer2's 3x3): two m-tiles share every weight stage
    if (cfg == kCfg128Deep && total_kblocks >= 8 && !residual && !out_fp32 && !env_flag("BV_NO_M2")) cfg = kCfg128M2;
    // ... and when the layer is exactly 128 wide the tile is computed transposed (weights as the A operand, N = 256 pixels)
    const bool tr_ok = N == 128 && nops == 1 && !residual && !out_fp32;
    if (cfg == kCfg128M2 && tr_ok && !env_flag("BV_NO_TR")) cfg = kCfg128TR;
    const bool tap3_ok = wide_ok && c0.cin == 64;
    if (cfg == kCfg64Wide && tap3_ok && !env_flag("BV_NO_TAP3")) cfg = kCfg64Tap3;
    // 256-wide tiles on CTA pairs where the device can co-schedule them.  Same-box A/B (profiles/r2b_*): every long-K
    // tile gains 6-14 % (layer3/4 3x3 and 1x1: the per-SM weight stream halves), residual-stream tiles 7-11 % from six
    // k-blocks on (layer4 conv3 + identity, layer2.0 conv3 + downsample) and 0-2 % below (layer3 conv3 + identity, K = 256:
    // HBM-bound on the identity rows either way).  End to end +3.9 % (long-K only) / +5.5 % (everything) on one box.
    // BV_PAIR (bit 0: long-K, bit 1: residual stream; default 3) and BV_PAIR_RES_MINKB (default 0) are the A/B switches.
    {
        static const int pair_mask = getenv("BV_PAIR") ? atoi(getenv("BV_PAIR")) : 3;
        static const int res_minkb = getenv("BV_PAIR_RES_MINKB") ? atoi(getenv("BV_PAIR_RES_MINKB")) : 0;
        static const int variant = getenv("BV_PAIR_VARIANT") ? atoi(getenv("BV_PAIR_VARIANT")) : 0;   // experiment: bit 0 / bit 1
        if (!out_fp32 && pair_launchable()) {
            if (cfg == kCfg256Deep && (pair_mask & 1)) cfg = (variant & 1) ? kCfg256PairDeepE16 : kCfg256PairDeep;
            else if (cfg == kCfg256Res && (pair_mask & 2) && total_kblocks >= res_minkb)
                cfg = (variant & 2) ? kCfg256PairRes37 : kCfg256PairRes;
        }
    }
    if (const char* force = getenv("BV_FORCE_CFG")) {
        const int f = atoi(force);
        if (f >= 0 && f < kNumCfg && N % kCfgBN[f] == 0 && (!kCfgPair[f] || (!out_fp32 && pair_launchable())) &&
            (f != kCfg64BRes || (N == 64 && total_kblocks <= bv::kMaxResidentKB)) && (f != kCfg128M2 || (!residual && !out_fp32)) && (f != kCfg128TR || tr_ok))
            cfg = f;
    }
    if (cfg == kCfg64Wide && !wide_ok) cfg = kCfg64;
    if (cfg == kCfg64Tap3 && !tap3_ok) cfg = kCfg64;
    const bool tap3 = cfg == kCfg64Tap3;
    const bool wide = cfg == kCfg64Wide || tap3;
    const int bn = kCfgBN[cfg];
    const long long M = wide ? (long long)B * Ho * (Wo + 2) : (long long)B * Ho * Wo;
    if (M <= 0 || M > 0x7fffffffLL - 256) return fail(BV_ERR_INVALID, "M=%lld out of range", M);
    const bool force_im2col = env_flag("BV_FORCE_IM2COL");
    for (int i = 0; i < nops; ++i) {
        const bv_conv& c = ops[i].c;
        if (c.cout != N) return fail(BV_ERR_INVALID, "operand %d cout mismatch", i);
        if (c.cin % 64 != 0) return fail(BV_ERR_INVALID, "cin=%d must be a multiple of 64", c.cin);
        if (conv_out_dim(ops[i].H, c.r, c.stride, c.pad) != Ho || conv_out_dim(ops[i].W, c.s, c.stride, c.pad) != Wo)
            return fail(BV_ERR_INVALID, "operand %d output size mismatch", i);
        bv::ConvSeg& sg = p.seg[i];
        sg.cblocks = c.cin / 64;
        sg.kblocks = c.r * c.s * sg.cblocks;
        sg.S = c.s;
        sg.stride = c.stride;
        sg.lower = -c.pad;
        const bool plain = (c.r == 1 && c.s == 1 && c.stride == 1 && c.pad == 0);
        sg.mode = wide ? bv::kSegWide : ((plain && !force_im2col) ? bv::kSegTiled : bv::kSegIm2col);
        int rc;
        if (sg.mode == bv::kSegTiled) {
            rc = make_tmap_2d(&p.tmA[i], ops[i].x, (uint64_t)c.cin, (uint64_t)M, bv::kBlockK, bv::kBlockM);
        } else {
            rc = make_tmap_im2col(&p.tmA[i], ops[i].x, B, ops[i].H, ops[i].W, c.cin, c.r, c.s, c.stride, c.pad, wide,
                                  tap3 ? 32 : 0);
        }
        if (rc) return rc;
        // (CTA pairs: each CTA loads half of the n-block's weight rows per stage)
        rc = make_tmap_2d(&p.tmB[i], c.w, (uint64_t)c.r * c.s * c.cin, (uint64_t)N, bv::kBlockK,
                          (uint32_t)(kCfgPair[cfg] ? bn / 2 : bn));
        if (rc) return rc;
        p.bias[i] = c.bias;
    }
    if (N > bv::kMaxBiasN) return fail(BV_ERR_INVALID, "more than %d output channels", bv::kMaxBiasN);
    {
        int rc = fetch_summed_bias(ops, nops, N, reinterpret_cast<float*>(p.bias_c));
        if (rc) return rc;
    }
    p.Wwide = wide ? Wo + 2 : 0;
    if (!out_fp32 && !wide) {
        int rc = make_tmap_2d(&p.tmOut, out, (uint64_t)N, (uint64_t)M, bv::kChunkCols, bv::kBlockM);
        if (rc) return rc;
        if (residual) {
            rc = make_tmap_2d(&p.tmRes, residual, (uint64_t)N, (uint64_t)M, bv::kChunkCols, bv::kBlockM);
            if (rc) return rc;
        }
    }
    p.nseg = nops;
    p.Ho = Ho;
    p.Wo = Wo;
    p.M = (int)M;
    p.N = N;
    p.num_m_blocks = tap3 ? (int)((M + bv::kTap3Rows - 1) / bv::kTap3Rows) : (int)((M + bv::kBlockM - 1) / bv::kBlockM);
    p.num_n_blocks = N / bn;
    p.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
    p.out = out;
    p.relu = relu;
    p.out_fp32 = out_fp32;
    {
        static const int res_prefetch = getenv("BV_RES_PREFETCH") ? atoi(getenv("BV_RES_PREFETCH")) : 0;
        p.res_prefetch = residual ? res_prefetch : 0;
    }
    L->bn = bn;
    L->cfg = cfg;
    // Partial TMEM accumulators (K steps dealt round-robin, summed in the epilogue).  Measured with
    // tools/umma_microbench.cu: tcgen05.mma throughput does not depend on accumulator reuse (68 / 78 / 128 cycles for
    // N = 64 / 128 / 256 either way: the floor is the 64 B/cycle A-operand read), so one accumulator is used.
    int nacc = 1;
    if (const char* f = getenv("BV_FORCE_NACC")) nacc = std::max(1, std::min(256 / bn, atoi(f)));
    p.nacc = nacc;
    const int mt = kCfgPair[cfg] ? 2 : kCfgMT[cfg];
    const long long tiles = (long long)((p.num_m_blocks + mt - 1) / mt) * p.num_n_blocks;
    if (mt > 1) nacc = 1;
    p.nacc = nacc;
    L->grid = kCfgPair[cfg] ? 2 * (int)std::min<long long>(tiles, g_num_sms / 2) : (int)std::min<long long>(tiles, g_num_sms);
    return BV_OK;
}


int launch_conv(const ConvLaunch& L0, cudaStream_t st) {
    ConvLaunch L = L0;
    if (bv::kTimingBuild && env_flag("BV_TIMING")) {
        long long*& g_dbg = t_dev->dbg;
        if (!g_dbg) cudaMalloc(&g_dbg, 4 * 8 * 1024);
        cudaMemsetAsync(g_dbg, 0, 4 * 8 * 1024, st);
        L.p.dbg = g_dbg;
    }
    switch (L.cfg) {
#define BV_LAUNCH(id, BN, ST, NB, BR, WD, MT, EP, TR, PR)                                                 \
    case id:                                                                                              \
        launch_ex(bv::conv_gemm_kernel<BN, ST, NB, BR, WD, MT, EP, TR, PR>, L.grid,                       \
                  bv::ConvGemmCfg<BN, ST, NB, BR, WD, MT, EP, TR, PR>::kThreads,                          \
                  bv::ConvGemmCfg<BN, ST, NB, BR, WD, MT, EP, TR, PR>::kSmemBytes, st, PR ? 2 : 1, L.p);  \
        break;
        BV_FOR_EACH_CFG(BV_LAUNCH)
#undef BV_LAUNCH
        case kCfg64Tap3:
            launch_ex(bv::conv3x3_tap3_kernel, L.grid, bv::kTap3Threads, bv::kTap3SmemBytes, st, 1, L.p);
            break;
        default:
            return fail(BV_ERR_INVALID, "unknown conv configuration %d", L.cfg);
    }
    BV_CUDA(cudaGetLastError());
    if (L.p.dbg) {
        static long long host[4 * 1024];
        cudaStreamSynchronize(st);
        cudaMemcpy(host, t_dev->dbg, sizeof(long long) * 4 * L.grid, cudaMemcpyDeviceToHost);
        double a[4] = {0, 0, 0, 0};
        for (int i = 0; i < L.grid; ++i)
            for (int j = 0; j < 4; ++j) a[j] += (double)host[i * 4 + j] / L.grid;
        fprintf(stderr, "[timing] cfg %d M=%d N=%d kb=%d: mma-warp total %.0f cyc, wait tmem_empty %.1f%%, wait full "
                        "%.1f%%; producer wait empty %.1f%%\n",
                L.cfg, L.p.M, L.p.N, L.p.seg[0].kblocks + (L.p.nseg > 1 ? L.p.seg[1].kblocks : 0), a[0],
                100 * a[1] / a[0], 100 * a[2] / a[0], 100 * a[3] / a[0]);
    }
    return BV_OK;
}

// conv3 (+ downsample | + residual) + ReLU chained with the next block's 1x1 conv1 + ReLU (chain_gemm.cuh).
// ops[0] = conv3 operand, ops[1] = optional downsample operand; `next` = the following conv1 (1x1, stride 1).
bool chain_supported(const ConvOperand* ops, int nops, const bv_conv& next) {
    const int n1 = ops[0].c.cout;
    if (n1 % bv::kChainBN1 != 0 || n1 > 1024) return false;
    if (next.r != 1 || next.s != 1 || next.stride != 1 || next.pad != 0 || next.cin != n1) return false;
    if (next.cout != 64 && next.cout != 128 && next.cout != 256) return false;
    for (int i = 0; i < nops; ++i)
        if (ops[i].c.cin % 64 != 0 || ops[i].c.cout != n1) return false;
    return true;
}

int build_chain(ChainLaunch* L, int B, const ConvOperand* ops, int nops, const void* residual, void* out1,
                const bv_conv& next, void* out2) {
    if (nops < 1 || nops > 2) return fail(BV_ERR_INVALID, "chain needs 1 or 2 operand pairs");
    if (nops == 2 && residual) return fail(BV_ERR_INVALID, "chain takes a downsample operand or a residual, not both");
    if (!chain_supported(ops, nops, next)) return fail(BV_ERR_INVALID, "unsupported shapes for the chained kernel");
    memset(&L->p, 0, sizeof(L->p));
    bv::ChainParams& p = L->p;
    const bv_conv& c0 = ops[0].c;
    const int Ho = conv_out_dim(ops[0].H, c0.r, c0.stride, c0.pad);
    const int Wo = conv_out_dim(ops[0].W, c0.s, c0.stride, c0.pad);
    const int N1 = c0.cout, N2 = next.cout;
    const long long M = (long long)B * Ho * Wo;
    if (M <= 0 || M > 0x7fffffffLL - 256) return fail(BV_ERR_INVALID, "M=%lld out of range", M);
    int k1 = 0;
    for (int i = 0; i < nops; ++i) {
        const bv_conv& c = ops[i].c;
        if (conv_out_dim(ops[i].H, c.r, c.stride, c.pad) != Ho || conv_out_dim(ops[i].W, c.s, c.stride, c.pad) != Wo)
            return fail(BV_ERR_INVALID, "operand %d output size mismatch", i);
        bv::ConvSeg& sg = p.seg[i];
        sg.cblocks = c.cin / 64;
        sg.kblocks = c.r * c.s * sg.cblocks;
        sg.S = c.s;
        sg.stride = c.stride;
        sg.lower = -c.pad;
        const bool plain = (c.r == 1 && c.s == 1 && c.stride == 1 && c.pad == 0);
        sg.mode = plain ? bv::kSegTiled : bv::kSegIm2col;
        int rc;
        if (plain)
            rc = make_tmap_2d(&p.tmA[i], ops[i].x, (uint64_t)c.cin, (uint64_t)M, bv::kBlockK, bv::kBlockM);
        else
            rc = make_tmap_im2col(&p.tmA[i], ops[i].x, B, ops[i].H, ops[i].W, c.cin, c.r, c.s, c.stride, c.pad);
        if (rc) return rc;
        rc = make_tmap_2d(&p.tmB1[i], c.w, (uint64_t)c.r * c.s * c.cin, (uint64_t)N1, bv::kBlockK, bv::kChainBN1);
        if (rc) return rc;
        k1 += sg.kblocks * 64;
    }
    int rc = make_tmap_2d(&p.tmB2, next.w, (uint64_t)N1, (uint64_t)N2, bv::kBlockK, (uint32_t)N2);
    if (rc) return rc;
    if ((rc = make_tmap_2d(&p.tmOut1, out1, (uint64_t)N1, (uint64_t)M, bv::kChunkCols, bv::kBlockM))) return rc;
    if ((rc = make_tmap_2d(&p.tmOut2, out2, (uint64_t)N2, (uint64_t)M, bv::kChunkCols, bv::kBlockM))) return rc;
    if (residual && (rc = make_tmap_2d(&p.tmRes, residual, (uint64_t)N1, (uint64_t)M, bv::kChunkCols, bv::kBlockM)))
        return rc;
    if ((rc = fetch_summed_bias(ops, nops, N1, reinterpret_cast<float*>(p.bias1_c)))) return rc;
    if ((rc = fetch_bias(nullptr, next.bias, N2, reinterpret_cast<float*>(p.bias2_c)))) return rc;
    p.nseg = nops;
    p.Ho = Ho;
    p.Wo = Wo;
    p.M = (int)M;
    p.N1 = N1;
    p.num_m_blocks = (int)((M + bv::kBlockM - 1) / bv::kBlockM);
    p.has_res = residual ? 1 : 0;
    // Measured (ncu dram__bytes_read, profiles/r1n): pulling the next tile into L2 one tile ahead makes the DRAM read
    // traffic 14-40 % LARGER (lines are evicted again before the smem pipeline reaches them) and the kernel slower.
    // BV_L2_PREFETCH = n > 1: only the residual, n sub-tiles (16 KB each) ahead of the one being loaded into smem
    p.l2_prefetch = getenv("BV_L2_PREFETCH") ? atoi(getenv("BV_L2_PREFETCH")) : 0;
    L->n2 = N2;
    L->cfg = (N2 == 64) ? ((nops == 2 && !env_flag("BV_CHAIN_NO_DEEP")) ? 3 : 0) : (N2 == 128 ? 1 : 2);
    if (N2 == 128 && N1 == 512 && env_flag("BV_CHAIN_L2_DEEPSTG")) L->cfg = 4;
    L->k1 = k1;
    L->grid = std::min(p.num_m_blocks, g_num_sms);
    return BV_OK;
}

bool l1_block_supported(const bv_conv& c2, const bv_conv& c3, const bv_conv& next, int W) {
    return W % bv::kTap3Group == 0 && c2.r == 3 && c2.s == 3 && c2.stride == 1 && c2.pad == 1 && c2.cin == 64 && c2.cout == 64 && c3.r == 1 &&
           c3.s == 1 && c3.stride == 1 && c3.pad == 0 && c3.cin == 64 && c3.cout == 256 && next.r == 1 && next.s == 1 &&
           next.stride == 1 && next.pad == 0 && next.cin == 256 && (next.cout == 64 || next.cout == 128);
}

bool l1_ds_supported(const bv_conv& ds) {
    return ds.w != nullptr && ds.r == 1 && ds.s == 1 && ds.stride == 1 && ds.pad == 0 && ds.cin == 64 && ds.cout == 256;
}

// t1 [B,H,W,64] -> out1 = relu(conv3(relu(conv2(t1))) + residual) [B,H,W,256], out2 = relu(next(out1)) [B,H,W,64]
// With a downsample branch (x0 [B,H,W,64], ds 64 -> 256 1x1) instead of a residual: out1 = relu(conv3(..) + ds(x0)).
int build_l1_block(L1Launch* L, int B, int H, int W, const void* t1, const bv_conv& c2, const bv_conv& c3,
                   const void* residual, void* out1, const bv_conv& next, void* out2, const void* x0 = nullptr,
                   const bv_conv* ds = nullptr, const HostBiasMap* bias_cache = nullptr) {
    if (!l1_block_supported(c2, c3, next, W))
        return fail(BV_ERR_INVALID, "unsupported shapes for the fused layer1 block (64->64 3x3, 64->256, 256->64, width %% 30 == 0)");
    const bool use_ds = x0 != nullptr && ds != nullptr;
    if (next.cout == 128 && use_ds) return fail(BV_ERR_INVALID, "the 128-wide successor form takes an identity residual");
    if (use_ds && (residual || !l1_ds_supported(*ds)))
        return fail(BV_ERR_INVALID, "the downsample form of the fused layer1 block takes a 64 -> 256 1x1 stride-1 branch and no residual");
    if (!use_ds && !residual) return fail(BV_ERR_INVALID, "the fused layer1 block needs an identity residual or a downsample branch");
    memset(&L->p, 0, sizeof(L->p));
    bv::L1BlockParams& p = L->p;
    const long long groups = (long long)B * H * (W / bv::kTap3Group);
    if (groups <= 0 || groups > 0x7fffffffLL / 64) return fail(BV_ERR_INVALID, "too many rows (%lld lane quarters)", groups);
    int rc;
    if ((rc = make_tmap_im2col(&p.tmA, t1, B, H, W, 64, 3, 3, 1, 1, true, 32))) return rc;
    if (use_ds) {
        if ((rc = make_tmap_lines(&p.tmX0, x0, B * H, W, 64, 32))) return rc;
        if ((rc = make_tmap_2d(&p.tmWd, ds->w, 64, 256, bv::kBlockK, 128))) return rc;
    } else {
        if ((rc = make_tmap_lines(&p.tmRes, residual, B * H, W, 256, 32))) return rc;
    }
    if ((rc = make_tmap_2d(&p.tmW2, c2.w, 576, 64, bv::kBlockK, 32))) return rc;
    if ((rc = make_tmap_2d(&p.tmW3, c3.w, 64, 256, bv::kBlockK, 128))) return rc;
    if ((rc = make_tmap_2d(&p.tmW1, next.w, 256, (uint64_t)next.cout, bv::kBlockK, (uint32_t)(next.cout / 2)))) return rc;
    {   // biases by value: bias3 (+ downsample bias) [256] | bias2 [64] | bias1 [next.cout]
        float* bc = reinterpret_cast<float*>(p.bias_c);
        if ((rc = fetch_bias(bias_cache, c3.bias, 256, bc))) return rc;
        if (use_ds) {
            float bd[256];
            if ((rc = fetch_bias(bias_cache, ds->bias, 256, bd))) return rc;
            for (int i = 0; i < 256; ++i) bc[i] += bd[i];
        }
        if ((rc = fetch_bias(bias_cache, c2.bias, 64, bc + 256))) return rc;
        if ((rc = fetch_bias(bias_cache, next.bias, next.cout, bc + 320))) return rc;
    }
    if ((rc = make_tmap_lines(&p.tmOut1, out1, B * H, W, 256, bv::kTap3Group))) return rc;
    // shifted-tap 3x3 GEMM: required by the 128-wide successor; for the 64-wide forms an A/B switch (BV_L1_SH, default on/off below)
    const int sh_mode = getenv("BV_L1_SH") ? atoi(getenv("BV_L1_SH")) : kL1ShiftedDefault;
    L->n2 = next.cout;
    L->sh = next.cout == 128 || sh_mode != 0;
    p.out2 = reinterpret_cast<__nv_bfloat16*>(out2);
    p.lines = B * H;
    p.Ho = H;
    p.Wo = W;
    p.groups_per_line = W / bv::kTap3Group;
    p.num_groups = (int)groups;
    p.num_tiles = (int)((groups + 3) / 4);
    p.num_pair_tiles = (p.num_tiles + 1) / 2;
    L->grid = 2 * std::min(p.num_pair_tiles, g_num_sms / 2);
    L->ds = use_ds;
    return BV_OK;
}

// Can the device co-schedule CTA pairs of the fused layer1 kernel at all (cluster launch, 224 KB of smem per CTA)?
// Queried once; when it cannot (MIG slice, odd SM count per TPC) the plan keeps the two-kernel path for that block.
bool l1_block_launchable() {
    if (!t_dev) return false;
    int& cached = t_dev->l1_launchable;
    if (cached >= 0) return cached == 1;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2, 1, 1);
    cfg.blockDim = dim3(bv::kL1Threads, 1, 1);
    cfg.dynamicSmemBytes = bv::L1Cfg<64>::kSmemBytes;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int clusters = 0;
    const cudaError_t e = cudaOccupancyMaxActiveClusters(&clusters, bv::l1_block_kernel<64>, &cfg);
    if (e != cudaSuccess) cudaGetLastError();
    cached = (e == cudaSuccess && clusters >= 1) ? 1 : 0;
    return cached == 1;
}

int launch_l1_block(const L1Launch& L, cudaStream_t st) {
    bv::L1BlockParams prm = L.p;
    // measured (ncu): the L2 prefetch adds 0.6 GB of DRAM reads per launch and does not change the duration
    prm.l2_prefetch = env_flag("BV_L1_PREFETCH") ? 1 : 0;
    if (bv::kTimingBuild && env_flag("BV_TIMING")) {
        long long*& g_dbg = t_dev->dbg;
        if (!g_dbg) cudaMalloc(&g_dbg, 4 * 8 * 1024);
        cudaMemsetAsync(g_dbg, 0, 4 * 8 * 1024, st);
        prm.dbg = g_dbg;
    }
    if (L.n2 == 128)
        BV_CUDA(launch_ex(bv::l1_block_kernel<128, false, true>, L.grid, bv::kL1Threads, bv::L1Cfg<128, false, true>::kSmemBytes, st, 2, prm));
    else if (L.ds && L.sh)
        BV_CUDA(launch_ex(bv::l1_block_kernel<64, true, true>, L.grid, bv::kL1Threads, bv::L1Cfg<64, true, true>::kSmemBytes, st, 2, prm));
    else if (L.ds)
        BV_CUDA(launch_ex(bv::l1_block_kernel<64, true>, L.grid, bv::kL1Threads, bv::L1Cfg<64, true>::kSmemBytes, st, 2, prm));
    else if (L.sh)
        BV_CUDA(launch_ex(bv::l1_block_kernel<64, false, true>, L.grid, bv::kL1Threads, bv::L1Cfg<64, false, true>::kSmemBytes, st, 2, prm));
    else
        BV_CUDA(launch_ex(bv::l1_block_kernel<64>, L.grid, bv::kL1Threads, bv::L1Cfg<64>::kSmemBytes, st, 2, prm));
    if (prm.dbg) {
        static long long host[8 * 128];
        const int pairs = L.grid / 2;
        cudaStreamSynchronize(st);
        cudaMemcpy(host, t_dev->dbg, sizeof(long long) * 8 * pairs, cudaMemcpyDeviceToHost);
        double a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int i = 0; i < pairs; ++i)
            for (int j = 0; j < 8; ++j) a[j] += (double)host[i * 8 + j] / pairs;
        fprintf(stderr, "[timing] l1_block: mma thread %.0f cyc; waits: d0_empty %.1f%% full %.1f%% d1_empty %.1f%% t2_ready %.1f%% "
                        "d2_empty %.1f%% sub_written %.1f%%\n",
                a[0], 100 * a[1] / a[0], 100 * a[2] / a[0], 100 * a[3] / a[0], 100 * a[4] / a[0], 100 * a[5] / a[0],
                100 * a[6] / a[0]);
        static long long h2[74 * 13];
        cudaMemcpy(h2, t_dev->dbg + 1024, sizeof(h2), cudaMemcpyDeviceToHost);
        double e[12] = {0}, tot = 0;
        for (int i = 0; i < pairs; ++i) {
            for (int j = 0; j < 12; ++j) e[j] += (double)h2[i * 12 + j] / pairs;
            tot += (double)h2[74 * 12 + i] / pairs;
        }
        const char* nm[12] = {"wait d0_full", "e0 compute+write", "wait t2_free", "wait d1_full", "wait res_ready", "e1 compute",
                              "e1 group barrier", "e1 copy-out", "wait d2_full", "e2 compute+copy", "e2 barrier", "between"};
        static long long tr[64];
        cudaMemcpy(tr, t_dev->dbg + 2048, sizeof(tr), cudaMemcpyDeviceToHost);
        const char* en[13] = {"G0 issued", "t2_ready seen", "G1 issued", "sub_written0", "sub_written1", "sub_written2", "sub_written3",
                              "G2 issued", "E0: d0_full seen", "E0: t2_ready arrived", "E1: d1_full seen", "E1: res_ready seen", "E1 done"};
        const long long base = tr[1];
        for (int t = 0; t < 2; ++t)
            for (int e = 0; e < 13; ++e) fprintf(stderr, "[trace] tile %d %-22s %8lld\n", 10 + t, en[e], tr[t * 32 + e] - base);
        for (int t = 0; t < 2; ++t)
            for (int e = 0; e < 5; ++e)
                fprintf(stderr, "[trace-ns] tile %d %-22s leader %8lld  peer %8lld\n", 10 + t, en[8 + e], tr[t * 32 + 16 + e] - tr[16],
                        tr[t * 32 + 24 + e] - tr[16]);
        fprintf(stderr, "[timing] l1_block epilogue warp 2 (%.0f cyc):", tot);
        for (int j = 0; j < 12; ++j) fprintf(stderr, " %s %.1f%%;", nm[j], 100 * e[j] / tot);
        fprintf(stderr, "\n");
    }
    return BV_OK;
}

int launch_chain(const ChainLaunch& L, cudaStream_t st) {
    switch (L.cfg) {
#define BV_LAUNCH_CHAIN(id, N2, ST, NB)                                                                    \
    case id:                                                                                               \
        launch_ex(bv::chain_gemm_kernel<N2, ST, NB>, L.grid, bv::kChainThreads,                            \
                  bv::ChainCfg<N2, ST, NB>::kSmemBytes, st, 1, L.p);                                       \
        break;
        BV_FOR_EACH_CHAIN(BV_LAUNCH_CHAIN)
#undef BV_LAUNCH_CHAIN
        default:
            return fail(BV_ERR_INVALID, "unknown chain configuration %d", L.cfg);
    }
    BV_CUDA(cudaGetLastError());
    return BV_OK;
}

// conv3 (1x1, K1 = 128 / 256) + identity + ReLU chained with the next block's conv1 (1x1, N2 = 128 / 256) on CTA pairs
int pair_chain_cfg(const bv_conv& c3, const bv_conv& next) {
    const bool plain3 = c3.r == 1 && c3.s == 1 && c3.stride == 1 && c3.pad == 0;
    const bool plain1 = next.r == 1 && next.s == 1 && next.stride == 1 && next.pad == 0;
    if (!plain3 || !plain1 || next.cin != c3.cout || c3.cout % bv::kChainBN1 != 0) return -1;
    if (next.cout == 256 && c3.cin == 256) return 0;
    if (next.cout == 256 && c3.cin == 128) return 1;
    if (next.cout == 128 && c3.cin == 128) return 2;
    return -1;
}

int build_pair_chain(PairChainLaunch* L, int B, int H, int W, const void* x, const bv_conv& c3, const void* residual,
                     void* out1, const bv_conv& next, void* out2) {
    const int cfg = pair_chain_cfg(c3, next);
    if (cfg < 0) return fail(BV_ERR_INVALID, "unsupported shapes for the pair-chained kernel (1x1 K=128/256 -> N1 %% 128 == 0 -> 1x1 N2=128/256)");
    if (!residual) return fail(BV_ERR_INVALID, "the pair-chained kernel needs an identity residual");
    if (!pair_launchable()) return fail(BV_ERR_INVALID, "this device cannot co-schedule CTA pairs");
    memset(&L->p, 0, sizeof(L->p));
    bv::PairChainParams& p = L->p;
    const long long M = (long long)B * H * W;
    if (M <= 0 || M > 0x7fffffffLL - 256) return fail(BV_ERR_INVALID, "M=%lld out of range", M);
    const int N1 = c3.cout, N2 = next.cout, K1 = c3.cin;
    int rc;
    if ((rc = make_tmap_2d(&p.tmA, x, (uint64_t)K1, (uint64_t)M, bv::kBlockK, bv::kBlockM))) return rc;
    if ((rc = make_tmap_2d(&p.tmB1, c3.w, (uint64_t)K1, (uint64_t)N1, bv::kBlockK, 64))) return rc;
    if ((rc = make_tmap_2d(&p.tmB2, next.w, (uint64_t)N1, (uint64_t)N2, bv::kBlockK, (uint32_t)(N2 / 2)))) return rc;
    if ((rc = make_tmap_2d(&p.tmRes, residual, (uint64_t)N1, (uint64_t)M, bv::kChunkCols, bv::kBlockM))) return rc;
    if ((rc = make_tmap_2d(&p.tmOut1, out1, (uint64_t)N1, (uint64_t)M, bv::kChunkCols, bv::kBlockM))) return rc;
    if (N1 > 1024 || N2 > 256) return fail(BV_ERR_INVALID, "pair-chained kernel: N1 <= 1024, N2 <= 256");
    if ((rc = fetch_bias(nullptr, c3.bias, N1, reinterpret_cast<float*>(p.bias1_c)))) return rc;
    if ((rc = fetch_bias(nullptr, next.bias, N2, reinterpret_cast<float*>(p.bias2_c)))) return rc;
    p.out2 = reinterpret_cast<__nv_bfloat16*>(out2);
    p.M = (int)M;
    p.N1 = N1;
    p.num_m_blocks = (int)((M + bv::kBlockM - 1) / bv::kBlockM);
    p.num_pair_tiles = (p.num_m_blocks + 1) / 2;
    p.res_prefetch = getenv("BV_RES_PREFETCH") ? atoi(getenv("BV_RES_PREFETCH")) : 0;
    L->cfg = cfg;
    if (cfg == 0) {   // experiment: ring / staging split of the layer3 class
        static const int v = getenv("BV_PC_VARIANT") ? atoi(getenv("BV_PC_VARIANT")) : 0;
        if (v == 1) L->cfg = 3;
        if (v == 2) L->cfg = 4;
    }
    L->n2 = N2;
    L->k1 = K1;
    L->grid = 2 * std::min(p.num_pair_tiles, g_num_sms / 2);
    return BV_OK;
}

int launch_pair_chain(const PairChainLaunch& L, cudaStream_t st) {
    switch (L.cfg) {
#define BV_LAUNCH_PC(id, N2, KB, ST, NS)                                                                   \
    case id:                                                                                               \
        launch_ex(bv::pair_chain_kernel<N2, KB, ST, NS>, L.grid, bv::kPcThreads,                           \
                  bv::PairChainCfg<N2, KB, ST, NS>::kSmemBytes, st, 2, L.p);                               \
        break;
        BV_FOR_EACH_PAIR_CHAIN(BV_LAUNCH_PC)
#undef BV_LAUNCH_PC
        default:
            return fail(BV_ERR_INVALID, "unknown pair-chain configuration %d", L.cfg);
    }
    BV_CUDA(cudaGetLastError());
    return BV_OK;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Activation buffers inside the caller's workspace.
struct Layout {
    size_t buf_a, buf_b, t1, t2, tds, hid, total;
};

const int kLayerBlocks[4] = {3, 4, 6, 3};
// Shape classes of the pair-chained kernel that are on by default: NONE.  Measured (profiles/r2f_*, r2g_*): layer3
// 0.61-0.66 ms per block against 0.415 + 0.212 ms for the pair conv3 + identity kernel followed by the pair conv1 kernel;
// layer2 1.17 ms against 1.01 ms for the single-CTA chain; layer2 -> layer3 1.42 ms against 0.79 + 0.47 ms.  The kernel
// removes the HBM re-read of the block output and 60 % of the L2 -> SM operand stream, but every staging sub-tile is held
// from the identity prefetch until both the TMA store and the second GEMM have read it, and 5-6 sub-tiles (all that fits
// next to the resident 64 KB input tile) do not cover that span.  Kept (bit-exact, tests/test_chain_gpu.py) behind
// BV_PAIR_CHAIN for the next attempt.
const int kPairChainDefault = 1;

const int kLayerWidth[4] = {64, 128, 256, 512};

Layout make_layout(int B, int C, int H, int W) {
    Layout l{};
    const size_t hw = (size_t)H * W;
    const size_t K = (C == 3) ? 192 : 64;
    const size_t patches = (size_t)B * (hw / 4) * K * 2;      // [B*H/2*W/2, K] bf16
    const size_t stem_out = (size_t)B * (hw / 4) * 64 * 2;    // [B,H/2,W/2,64]
    const size_t block_out = (size_t)B * (hw / 16) * 256 * 2;  // layer1 output, the largest block output
    const size_t t1 = (size_t)B * (hw / 16) * 128 * 2;         // layer2.0 conv1 output
    const size_t t2 = t1;                                      // layer1 conv2 output; same size as t1: the fused layer1
                                                               // block swaps the roles of the two scratch buffers
    const size_t tds = block_out;                              // un-fused downsample output (debug path)
    const size_t hid = (size_t)B * (hw / 1024) * 128 * 4;      // projector hidden, fp32
    size_t off = 0;
    l.buf_a = off; off += align_up(std::max(patches, block_out), 1024);
    l.buf_b = off; off += align_up(std::max(stem_out, block_out), 1024);
    l.t1 = off; off += align_up(t1, 1024);
    l.t2 = off; off += align_up(t2, 1024);
    l.hid = off; off += align_up(hid, 1024);
    l.tds = off;
    if (env_flag("BV_NO_FUSE_DS")) off += align_up(tds, 1024);
    l.total = off + 1024;  // slack to align the caller's base pointer
    return l;
}

}  // namespace

struct bv_handle {
    bv_weights w;
    int device;
    HostBiasMap host_bias;    // host copies of the bias vectors (see fetch_bias)
    // prompts
    float* yn = nullptr;      // [L][2][P][128] unit vectors
    float* heat_t = nullptr;  // [L][128]
    size_t yn_cap = 0, heat_cap = 0;   // capacities in 128-float rows
    int L = 0, P = 0;
    // cached plan
    struct Key {
        void* workspace;
        int dtype, B, C, H, W;
        bool operator==(const Key& o) const {
            return workspace == o.workspace && dtype == o.dtype && B == o.B && C == o.C &&
                   H == o.H && W == o.W;
        }
    } key{};
    bool plan_valid = false;
    std::vector<PlanStep> steps;    // in execution order (stem GEMM first)
    bv::StemParams stem{};          // fused 8-bit stem (valid when use_fused_stem)
    bool use_fused_stem = false;
    bv::StemRowsParams stem_rows{}; // row-streaming 8-bit stem (valid when use_stem_rows)
    bool use_stem_rows = false;
    const void* trunk = nullptr;    // final [B,h,w,2048] bf16
    int last_launches = 0;
    // optional per-launch timing (cudaEvents on the launch stream)
    bool profile = false;
    std::vector<cudaEvent_t> events;
    std::vector<bv_launch_info> infos;
    int n_events = 0;
    // bv_forward_graph: instantiated CUDA graphs of whole forwards, keyed by every argument that is baked into the nodes
    struct GraphEntry {
        std::vector<uintptr_t> key;
        cudaGraphExec_t exec;
        int launches;
    };
    std::vector<GraphEntry> graphs;
    cudaStream_t capture_stream = nullptr;
};

namespace {
// Record a timestamp on the launch stream when profiling is on (one event before the first launch, one after each).
void prof_mark(bv_handle* h, cudaStream_t st, const char* name, double flops, double bytes) {
    if (!h->profile) return;
    if ((int)h->events.size() <= h->n_events) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        h->events.push_back(e);
    }
    cudaEventRecord(h->events[h->n_events++], st);
    if (name) {
        bv_launch_info info{};
        snprintf(info.name, sizeof(info.name), "%s", name);
        info.flops = flops;
        info.bytes = bytes;
        info.ms = 0.f;
        h->infos.push_back(info);
    }
}

void conv_cost(const ConvLaunch& L, double* flops, double* bytes, char* name, size_t n) {
    const bv::ConvGemmParams& p = L.p;
    double k = 0, abytes = 0;
    for (int i = 0; i < p.nseg; ++i) {
        k += (double)p.seg[i].kblocks * 64;
        // algorithmic A traffic: every input element once (taps and strides re-read through L2, not HBM)
        const double rows = (p.seg[i].mode == bv::kSegTiled) ? (double)p.M : (double)p.M * p.seg[i].stride * p.seg[i].stride;
        abytes += rows * p.seg[i].cblocks * 64 * 2;
    }
    *flops = 2.0 * p.M * p.N * k;
    *bytes = abytes + (double)p.M * p.N * (p.out_fp32 ? 4 : 2) + (p.residual ? (double)p.M * p.N * 2 : 0) + k * p.N * 2;
    snprintf(name, n, "conv_gemm<%d/c%d> M=%d N=%d K=%d%s%s", L.bn, L.cfg, p.M, p.N, (int)k, p.nseg > 1 ? " +ds" : "",
             p.residual ? " +res" : "");
}
void chain_cost(const ChainLaunch& L, double* flops, double* bytes, char* name, size_t n) {
    const bv::ChainParams& p = L.p;
    double abytes = 0;
    for (int i = 0; i < p.nseg; ++i) {
        const double rows = (p.seg[i].mode == bv::kSegTiled) ? (double)p.M : (double)p.M * p.seg[i].stride * p.seg[i].stride;
        abytes += rows * p.seg[i].cblocks * 64 * 2;
    }
    *flops = 2.0 * p.M * p.N1 * L.k1 + 2.0 * p.M * L.n2 * p.N1;
    *bytes = abytes + (double)p.M * p.N1 * 2 * (p.has_res ? 2 : 1) + (double)p.M * L.n2 * 2 +
             ((double)L.k1 * p.N1 + (double)p.N1 * L.n2) * 2;
    snprintf(name, n, "chain_gemm<%d> M=%d N1=%d K1=%d%s", L.n2, p.M, p.N1, L.k1, p.nseg > 1 ? " +ds" : " +res");
}

void l1_cost(const L1Launch& L, double* flops, double* bytes, char* name, size_t n) {
    const double Mr = (double)L.p.num_groups * bv::kTap3Group;   // output pixels
    *flops = 2.0 * Mr * 64 * 576 + 2.0 * Mr * 256 * 64 * (L.ds ? 2 : 1) + 2.0 * Mr * L.n2 * 256;
    *bytes = Mr * 64 * 2 + (L.ds ? Mr * 64 * 2 + Mr * 256 * 2 : Mr * 256 * 2 * 2) + Mr * L.n2 * 2 +
             (576.0 * 64 + 64 * 256 * (L.ds ? 2 : 1) + 256.0 * L.n2) * 2;
    snprintf(name, n, "l1_block<%d%s> M=%.0f 3x3(64)+1x1(256)+%s+1x1(%d)", L.n2, L.sh ? "/sh" : "", Mr, L.ds ? "ds" : "res", L.n2);
}

void pair_chain_cost(const PairChainLaunch& L, double* flops, double* bytes, char* name, size_t n) {
    const bv::PairChainParams& p = L.p;
    *flops = 2.0 * p.M * p.N1 * L.k1 + 2.0 * p.M * L.n2 * p.N1;
    *bytes = (double)p.M * L.k1 * 2 + (double)p.M * p.N1 * 2 * 2 + (double)p.M * L.n2 * 2 +
             ((double)L.k1 * p.N1 + (double)p.N1 * L.n2) * 2;
    snprintf(name, n, "pair_chain<%d> M=%d N1=%d K1=%d +res", L.n2, p.M, p.N1, L.k1);
}

void step_cost(const PlanStep& s, double* flops, double* bytes, char* name, size_t n) {
    if (s.pc) pair_chain_cost(s.pch, flops, bytes, name, n);
    else if (s.l1) l1_cost(s.l1b, flops, bytes, name, n);
    else if (s.chain) chain_cost(s.ch, flops, bytes, name, n);
    else conv_cost(s.conv, flops, bytes, name, n);
}

int launch_step(const PlanStep& s, cudaStream_t st) {
    if (s.pc) return launch_pair_chain(s.pch, st);
    if (s.l1) return launch_l1_block(s.l1b, st);
    return s.chain ? launch_chain(s.ch, st) : launch_conv(s.conv, st);
}
}  // namespace

extern "C" {

const char* bv_last_error(void) { return g_err.c_str(); }
const char* bv_version(void) { return "biovil_b200 0.1 (sm_100a, tcgen05 implicit-GEMM)"; }

int32_t bv_patch_grid(int32_t size) { return size / 32; }

size_t bv_workspace_bytes(int32_t B, int32_t C, int32_t H, int32_t W) {
    if (B <= 0 || H <= 0 || W <= 0 || H % 32 || W % 32 || (C != 1 && C != 3)) return 0;
    return make_layout(B, C, H, W).total;
}

int32_t bv_create(bv_handle** out, const bv_weights* w, int32_t device) {
    if (!out || !w) return fail(BV_ERR_INVALID, "null argument");
    if (cudaSetDevice(device) != cudaSuccess) {
        cudaGetLastError();
        return fail(BV_ERR_NO_DEVICE, "cudaSetDevice(%d) failed: no usable CUDA device (no CPU fallback)", device);
    }
    int rc = device_setup();
    if (rc) return rc;
    bv_handle* h = new bv_handle();
    h->w = *w;
    h->device = device;
    // host copies of the bias vectors (the convolution kernels take them by value): the packing kernels that wrote them may
    // still be running on any stream of the caller
    {
        bool any = false;
        for (int i = 0; i < BV_NUM_BLOCKS && !any; ++i) any = w->conv2[i].bias != nullptr;
        if (any) {
            if (cudaDeviceSynchronize() != cudaSuccess) {
                delete h;
                return fail(BV_ERR_CUDA, "cudaDeviceSynchronize failed: %s", cudaGetErrorString(cudaGetLastError()));
            }
            std::vector<const bv_conv*> convs = {&w->stem_u8, &w->stem_f1, &w->stem_f3, &w->proj0};
            for (int i = 0; i < BV_NUM_BLOCKS; ++i)
                for (const bv_conv* c : {&w->conv1[i], &w->conv2[i], &w->conv3[i], &w->downsample[i]}) convs.push_back(c);
            for (const bv_conv* c : convs) {
                if (!c->bias || c->cout <= 0) continue;
                std::vector<float>& v = h->host_bias[c->bias];
                v.resize(c->cout);
                if (cudaMemcpy(v.data(), c->bias, sizeof(float) * c->cout, cudaMemcpyDeviceToHost) != cudaSuccess) {
                    delete h;
                    return fail(BV_ERR_CUDA, "copying a bias vector to the host failed: %s", cudaGetErrorString(cudaGetLastError()));
                }
            }
        }
    }
    *out = h;
    return BV_OK;
}

void bv_destroy(bv_handle* h) {
    if (!h) return;
    if (h->yn) cudaFree(h->yn);
    if (h->heat_t) cudaFree(h->heat_t);
    for (cudaEvent_t e : h->events) cudaEventDestroy(e);
    for (auto& g : h->graphs) cudaGraphExecDestroy(g.exec);
    if (h->capture_stream) cudaStreamDestroy(h->capture_stream);
    delete h;
}

int32_t bv_set_prompts(bv_handle* h, const float* prompts, int32_t L, int32_t P, const float* heat_text,
                       bv_stream stream) {
    if (!h || !prompts || L <= 0 || P <= 0) return fail(BV_ERR_INVALID, "bad prompt arguments");
    int rc = device_setup();
    if (rc) return rc;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int NP = L * 2 * P;
    // The prompt buffers are reused while they are large enough: cudaFree synchronises the whole device, and callers such
    // as Trainer.myCosineSimilarity's drop-in install prompts 2 x L times per batch.  Growing allocates the new buffers
    // first, so a failed allocation leaves the handle's previous prompt set intact.
    if ((size_t)NP > h->yn_cap || (size_t)L > h->heat_cap) {
        float *yn = nullptr, *ht = nullptr;
        const size_t yn_cap = std::max<size_t>(NP, 2 * h->yn_cap), heat_cap = std::max<size_t>(L, 2 * h->heat_cap);
        BV_CUDA(cudaMalloc(&yn, yn_cap * bv::kEmbDim * sizeof(float)));
        cudaError_t e = cudaMalloc(&ht, heat_cap * bv::kEmbDim * sizeof(float));
        if (e != cudaSuccess) {
            cudaFree(yn);
            return fail(BV_ERR_CUDA, "cudaMalloc of the heat-map text buffer failed: %s", cudaGetErrorString(e));
        }
        if (h->yn) cudaFree(h->yn);          // (synchronises: kernels still reading the old set finish first)
        if (h->heat_t) cudaFree(h->heat_t);
        h->yn = yn;
        h->heat_t = ht;
        h->yn_cap = yn_cap;
        h->heat_cap = heat_cap;
    }
    bv::prompt_normalize_kernel<<<(NP + 3) / 4, 128, 0, st>>>(prompts, h->yn, NP);
    BV_CUDA(cudaGetLastError());
    if (heat_text) {
        bv::prompt_normalize_kernel<<<(L + 3) / 4, 128, 0, st>>>(heat_text, h->heat_t, L);
        BV_CUDA(cudaGetLastError());
    } else {
        // positive prompt 0 of each label: rows l*2*P of yn
        BV_CUDA(cudaMemcpy2DAsync(h->heat_t, bv::kEmbDim * sizeof(float), h->yn,
                                  (size_t)2 * P * bv::kEmbDim * sizeof(float), bv::kEmbDim * sizeof(float), L,
                                  cudaMemcpyDeviceToDevice, st));
    }
    h->L = L;
    h->P = P;
    return BV_OK;
}

int32_t bv_pairwise_cosine(const float* x, const float* y, int32_t B, int32_t P, int32_t reduce_max, float* out,
                           bv_stream stream) {
    if (!x || !y || !out || B <= 0 || P <= 0) return fail(BV_ERR_INVALID, "bad cosine arguments");
    int rc = device_setup();
    if (rc) return rc;
    const int blocks = std::min((B + 7) / 8, g_num_sms * 8);
    bv::pairwise_cosine_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, y, B, P, reduce_max, out);
    BV_CUDA(cudaGetLastError());
    return BV_OK;
}

int32_t bv_score(bv_handle* h, const float* emb, int32_t B, float* sim, float* prob, uint8_t* pred, float* score,
                 bv_stream stream) {
    if (!h || !emb || B <= 0) return fail(BV_ERR_INVALID, "bad score arguments");
    if (!h->yn) return fail(BV_ERR_INVALID, "bv_set_prompts has not been called");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    bv::ScoreOut o{sim, prob, pred, score};
    int rc = device_setup();
    if (rc) return rc;
    const int blocks = std::min((B + 7) / 8, g_num_sms * 8);
    bv::score_kernel<<<blocks, 256, 0, st>>>(emb, h->yn, B, h->L, h->P, o);
    BV_CUDA(cudaGetLastError());
    return BV_OK;
}

int32_t bv_last_forward_launches(const bv_handle* h) { return h ? h->last_launches : 0; }

int32_t bv_set_profile(bv_handle* h, int32_t enable) {
    if (!h) return fail(BV_ERR_INVALID, "null handle");
    h->profile = enable != 0;
    return BV_OK;
}

int32_t bv_get_profile(bv_handle* h, bv_launch_info* out, int32_t capacity) {
    if (!h) return fail(BV_ERR_INVALID, "null handle");
    const int n = (int)h->infos.size();
    if (h->n_events != n + 1 && n > 0) return fail(BV_ERR_INVALID, "profile is incomplete");
    if (n > 0) BV_CUDA(cudaEventSynchronize(h->events[n]));
    for (int i = 0; i < n && i < capacity; ++i) {
        float ms = 0.f;
        BV_CUDA(cudaEventElapsedTime(&ms, h->events[i], h->events[i + 1]));
        out[i] = h->infos[i];
        out[i].ms = ms;
    }
    return n;
}

static int build_plan(bv_handle* h, const void* frames, int dtype, int B, int C, int H, int W, uint8_t* ws) {
    const Layout lay = make_layout(B, C, H, W);
    h->steps.clear();
    struct BiasCacheScope {   // the builders below look the bias vectors up in this handle's host copies
        explicit BiasCacheScope(const HostBiasMap* m) { t_bias_cache = m; }
        ~BiasCacheScope() { t_bias_cache = nullptr; }
    } bias_scope(&h->host_bias);
    const bool fuse_ds = !env_flag("BV_NO_FUSE_DS");
    const bool use_chain = !env_flag("BV_NO_CHAIN");
    uint8_t* buf_a = ws + lay.buf_a;
    uint8_t* buf_b = ws + lay.buf_b;
    uint8_t* t1 = ws + lay.t1;
    uint8_t* t2 = ws + lay.t2;
    uint8_t* tds = ws + lay.tds;
    const int H2 = H / 2, W2 = W / 2;
    auto push_conv = [&](const ConvOperand* ops, int nops, const void* residual, int relu, void* out,
                         int out_fp32) -> int {
        PlanStep s;
        s.chain = false;
        int rc = build_conv(&s.conv, B, ops, nops, residual, relu, out, out_fp32);
        if (rc) return rc;
        h->steps.push_back(s);
        return BV_OK;
    };
    int rc;
    // stem GEMM: patches [B*H2*W2, K] (buf_a) x stem weights -> stem_out (buf_b), bias + ReLU
    {
        const bv_conv& sc = (dtype == BV_DTYPE_U8) ? h->w.stem_u8 : (C == 3 ? h->w.stem_f3 : h->w.stem_f1);
        bv_conv g = sc;  // viewed as a 1x1 conv over the gathered patch matrix
        g.r = g.s = 1;
        g.stride = 1;
        g.pad = 0;
        ConvOperand op{buf_a, H2, W2, g};
        if ((rc = push_conv(&op, 1, nullptr, 1, buf_b, 0))) return rc;
    }
    // max-pool writes x0 into buf_a; blocks ping-pong buf_a <-> buf_b
    uint8_t* cur = buf_a;
    uint8_t* nxt = buf_b;
    int ch = H / 4, cw = W / 4;
    int blk = 0;
    // 8-bit frames: the row-streaming stem kernel can also produce layer1.0's conv1 output (BV_STEM_C1).  Off by default:
    // measured 0.91 ms against 0.565 + 0.30 ms for the two kernels (the fused form gives up one of four TMEM stages,
    // and its second epilogue competes with the first for the instruction slots that bound the kernel)
    const bool stem_rows = (dtype == BV_DTYPE_U8) && h->w.stem_u8_k8.w != nullptr && !env_flag("BV_NO_FUSED_STEM") &&
                           !env_flag("BV_STEM_V1");
    const bv_conv& l1c1 = h->w.conv1[0];
    const bool stem_c1 = stem_rows && env_flag("BV_STEM_C1") && l1c1.cin == 64 && l1c1.cout == 64 && l1c1.r == 1 &&
                         l1c1.s == 1 && l1c1.stride == 1 && l1c1.pad == 0 && (H / 4) % 2 == 0;
    uint8_t* stem_t1 = t1;
    bool t1_ready = stem_c1;  // this block's conv1 output was already produced by the previous kernel (chain / stem)
    for (int li = 0; li < 4; ++li) {
        for (int bi = 0; bi < kLayerBlocks[li]; ++bi, ++blk) {
            const bv_conv& c1 = h->w.conv1[blk];
            const bv_conv& c2 = h->w.conv2[blk];
            const bv_conv& c3 = h->w.conv3[blk];
            const bv_conv& ds = h->w.downsample[blk];
            const int oh = conv_out_dim(ch, c2.r, c2.stride, c2.pad);
            const int ow = conv_out_dim(cw, c2.s, c2.stride, c2.pad);
            if (!t1_ready) {
                ConvOperand o1{cur, ch, cw, c1};
                if ((rc = push_conv(&o1, 1, nullptr, 1, t1, 0))) return rc;
            }
            // layer1 blocks with an identity residual whose successor's conv1 is 64 wide: the whole tail of the block
            // (conv2 3x3, conv3 + identity, next conv1) is one CTA-pair kernel
            const bool l1_ds = ds.w != nullptr && l1_ds_supported(ds) && !env_flag("BV_NO_L1_DS");
            const int l1_last = getenv("BV_L1_LAST") ? atoi(getenv("BV_L1_LAST")) : kL1LastBlockDefault;
            const bool l1_wide_next = blk + 1 < BV_NUM_BLOCKS && h->w.conv1[blk + 1].cout == 128;
            if (!env_flag("BV_NO_L1_FUSED") && l1_block_launchable() && (ds.w == nullptr || l1_ds) && blk + 1 < BV_NUM_BLOCKS &&
                (!l1_wide_next || (l1_last && ds.w == nullptr)) && l1_block_supported(c2, c3, h->w.conv1[blk + 1], cw)) {
                PlanStep s;
                s.l1 = true;
                if (l1_ds) rc = build_l1_block(&s.l1b, B, ch, cw, t1, c2, c3, nullptr, nxt, h->w.conv1[blk + 1], t2, cur, &ds, &h->host_bias);
                else rc = build_l1_block(&s.l1b, B, ch, cw, t1, c2, c3, cur, nxt, h->w.conv1[blk + 1], t2, nullptr, nullptr, &h->host_bias);
                if (rc) return rc;
                h->steps.push_back(s);
                // the next block's conv1 output went to t2: swap the roles of the two scratch buffers
                std::swap(t1, t2);
                t1_ready = true;
                std::swap(cur, nxt);
                continue;
            }
            ConvOperand o2{t1, ch, cw, c2};
            if ((rc = push_conv(&o2, 1, nullptr, 1, t2, 0))) return rc;
            // conv3 + (downsample | identity) + ReLU, chained with the next block's conv1 where the shapes allow
            ConvOperand o3[2] = {{t2, oh, ow, c3}, {cur, ch, cw, ds}};
            int n3 = 1;
            const void* res = cur;
            if (ds.w != nullptr) {
                if (fuse_ds) {
                    n3 = 2;
                    res = nullptr;
                } else {
                    ConvOperand od{cur, ch, cw, ds};
                    if ((rc = push_conv(&od, 1, nullptr, 0, tds, 0))) return rc;
                    res = tds;
                }
            }
            t1_ready = false;
            // Deep-layer identity blocks: conv3 + identity chained with the next conv1 on CTA pairs (pair_chain.cuh).
            // BV_PAIR_CHAIN selects the shape classes (bit 0: K1 = 256 -> N2 = 256 [layer3], bit 1: K1 = 128 -> N2 = 256
            // [layer2 -> layer3], bit 2: K1 = 128 -> N2 = 128 [layer2]).
            {
                static const int pc_mask = getenv("BV_PAIR_CHAIN") ? atoi(getenv("BV_PAIR_CHAIN")) : kPairChainDefault;
                const int pcc = (n3 == 1 && res != nullptr && blk + 1 < BV_NUM_BLOCKS) ? pair_chain_cfg(c3, h->w.conv1[blk + 1]) : -1;
                if (pcc >= 0 && ((pc_mask >> pcc) & 1) && pair_launchable()) {
                    PlanStep s;
                    s.pc = true;
                    if ((rc = build_pair_chain(&s.pch, B, oh, ow, t2, c3, res, nxt, h->w.conv1[blk + 1], t1))) return rc;
                    h->steps.push_back(s);
                    t1_ready = true;
                    std::swap(cur, nxt);
                    ch = oh;
                    cw = ow;
                    continue;
                }
            }
            // Measured losses stay unchained: the strided-downsample tail of layer2.0 (six A k-blocks re-streamed per
            // chunk) and the 256-wide second GEMM into layer3 (its staging leaves too little residual prefetch depth).
            const bool chain_l3 = env_flag("BV_CHAIN_L3") && c3.cout == 1024 && n3 == 1;
            const bool chain_256 = env_flag("BV_CHAIN_256") && n3 == 1 && c3.cout == 512;   // layer2 -> layer3 transition
            const bool chain_pays = env_flag("BV_CHAIN_ALL") || chain_l3 || chain_256 ||
                                    (!(n3 == 2 && ds.cin * ds.r * ds.s > 64) && h->w.conv1[blk + 1 < BV_NUM_BLOCKS ? blk + 1 : blk].cout <= 128);
            if (use_chain && chain_pays && blk + 1 < BV_NUM_BLOCKS && chain_supported(o3, n3, h->w.conv1[blk + 1])) {
                PlanStep s;
                s.chain = true;
                if ((rc = build_chain(&s.ch, B, o3, n3, res, nxt, h->w.conv1[blk + 1], t1))) return rc;
                h->steps.push_back(s);
                t1_ready = true;
            } else {
                if ((rc = push_conv(o3, n3, res, 1, nxt, 0))) return rc;
            }
            std::swap(cur, nxt);
            ch = oh;
            cw = ow;
        }
    }
    h->trunk = cur;
    // projector conv 2048 -> 128 (+BN, ReLU), fp32 output for the fp32 tail
    {
        ConvOperand op{cur, ch, cw, h->w.proj0};
        if ((rc = push_conv(&op, 1, nullptr, 1, ws + lay.hid, 1))) return rc;
    }
    h->use_fused_stem = (dtype == BV_DTYPE_U8) && h->w.stem_u8_k8.w != nullptr && !env_flag("BV_NO_FUSED_STEM");
    if (h->use_fused_stem) {
        static_assert(bv::kStemSmemBytes <= bv::kStemSmemRequest, "stem smem request too small");
        memset(&h->stem, 0, sizeof(h->stem));
        if ((rc = make_tmap_2d(&h->stem.tmW, h->w.stem_u8_k8.w, 64, 64, 64, 64))) return rc;
        h->stem.bias = h->w.stem_u8_k8.bias;
        h->stem.out = reinterpret_cast<__nv_bfloat16*>(buf_a);
        h->stem.B = B;
        h->stem.H = H;
        h->stem.W = W;
        h->stem.tiles_x = (W / 4) / bv::kStemPool;
        h->stem.tiles_y = (H / 4) / bv::kStemPool;
    }
    // Row-streaming stem (stem_rows.cuh): default for 8-bit frames; BV_STEM_V1 keeps the tile kernel above.
    h->use_stem_rows = h->use_fused_stem && !env_flag("BV_STEM_V1");
    if (h->use_stem_rows) {
        memset(&h->stem_rows, 0, sizeof(h->stem_rows));
        h->stem_rows.w = reinterpret_cast<const __nv_bfloat16*>(h->w.stem_u8_k8.w);
        h->stem_rows.out = reinterpret_cast<__nv_bfloat16*>(buf_a);
        h->stem_rows.B = B;
        h->stem_rows.H = H;
        h->stem_rows.W = W;
        h->stem_rows.strips_x = (W / 4 + bv::kSrStripPx - 1) / bv::kSrStripPx;
        if (stem_c1) {
            h->stem_rows.w1 = reinterpret_cast<const __nv_bfloat16*>(l1c1.w);
            h->stem_rows.b1 = l1c1.bias;
            h->stem_rows.out1 = reinterpret_cast<__nv_bfloat16*>(stem_t1);
        }
    }
    (void)frames;
    return BV_OK;
}

}  // extern "C"

namespace {
int forward_impl(bv_handle* h, const void* frames, int32_t dtype, int32_t B, int32_t C, int32_t H, int32_t W,
                 void* workspace, size_t workspace_bytes, const bv_outputs* out, bv_stream stream) {
    if (!h || !frames || !workspace || !out) return fail(BV_ERR_INVALID, "null argument");
    if (!h->w.proj3_wt || !h->w.proj0.w || !h->w.stem_u8.w)
        return fail(BV_ERR_INVALID, "this handle was created without image-model weights (scoring only)");
    if (B <= 0 || H <= 0 || W <= 0 || H % 32 || W % 32)
        return fail(BV_ERR_INVALID, "frames must be [B,C,H,W] with H, W multiples of 32 (got %dx%dx%dx%d)", B, C, H, W);
    if (!((dtype == BV_DTYPE_U8 && C == 1) || (dtype == BV_DTYPE_F32 && (C == 1 || C == 3))))
        return fail(BV_ERR_INVALID, "unsupported input: dtype=%d channels=%d (u8 x1, f32 x1, f32 x3)", dtype, C);
    const Layout lay = make_layout(B, C, H, W);
    uint8_t* ws = reinterpret_cast<uint8_t*>(align_up(reinterpret_cast<size_t>(workspace), 1024));
    if (workspace_bytes < lay.total)
        return fail(BV_ERR_WORKSPACE, "workspace too small: %zu < %zu", workspace_bytes, lay.total);
    if ((out->sim || out->prob || out->pred || out->score || out->heat) && !h->yn)
        return fail(BV_ERR_INVALID, "scores requested but bv_set_prompts has not been called");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int rc = device_setup();
    if (rc) return rc;
    bv_handle::Key key{ws, dtype, B, C, H, W};
    if (!h->plan_valid || !(h->key == key)) {
        h->plan_valid = false;
        rc = build_plan(h, frames, dtype, B, C, H, W, ws);
        if (rc) return rc;
        h->key = key;
        h->plan_valid = true;
    }
    int launches = 0;
    h->n_events = 0;
    h->infos.clear();
    prof_mark(h, st, nullptr, 0, 0);
    char pname[64];
    double pflops = 0, pbytes = 0;
    const int H2 = H / 2, W2 = W / 2, H4 = H / 4, W4 = W / 4;
    __nv_bfloat16* patches = reinterpret_cast<__nv_bfloat16*>(ws + lay.buf_a);
    if (h->use_fused_stem && (reinterpret_cast<uintptr_t>(frames) & 7u))
        return fail(BV_ERR_INVALID, "8-bit frames must be 8-byte aligned");
    if (h->use_stem_rows) {
        h->stem_rows.frames = reinterpret_cast<const uint8_t*>(frames);
        const int strips = B * h->stem_rows.strips_x;
        const int grid = std::min(strips, g_num_sms);
        launch_stem_rows(h->stem_rows, grid, st);
        BV_CUDA(cudaGetLastError());
        ++launches;
        const bool c1 = h->stem_rows.out1 != nullptr;
        prof_mark(h, st, c1 ? "stem_rows conv7x7+bn+relu+maxpool+conv1x1(64)" : "stem_rows conv7x7+bn+relu+maxpool",
                  2.0 * B * H2 * W2 * 64 * 49 + (c1 ? 2.0 * B * H4 * W4 * 64 * 64 : 0.0),
                  (double)B * H * W + (double)B * H4 * W4 * 64 * 2 * (c1 ? 2 : 1));
    } else if (h->use_fused_stem) {
        // 1-3 fused: conv7x7/2 + BN + ReLU + max-pool in one kernel, frame bytes in, layer1 input out
        h->stem.frames = reinterpret_cast<const uint8_t*>(frames);
        const int tiles = B * h->stem.tiles_x * h->stem.tiles_y;
        const int grid = std::min(tiles, 2 * g_num_sms);
        bv::stem_fused_kernel<<<grid, bv::kStemThreads, bv::kStemSmemRequest, st>>>(h->stem);
        BV_CUDA(cudaGetLastError());
        ++launches;
        prof_mark(h, st, "stem_fused conv7x7+bn+relu+maxpool", 2.0 * B * H2 * W2 * 64 * 49,
                  (double)B * H * W + (double)B * H4 * W4 * 64 * 2);
    } else {
        // 1. stem patch gather
        {
            const long long total = (long long)B * H2 * W2 * ((C == 3) ? 24 : 8);
            const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)g_num_sms * 16);
            if (dtype == BV_DTYPE_U8)
                bv::stem_patch_kernel<uint8_t, 1><<<blocks, 256, 0, st>>>(reinterpret_cast<const uint8_t*>(frames), patches,
                                                                        B, H, W, H2, W2);
            else if (C == 1)
                bv::stem_patch_kernel<float, 1><<<blocks, 256, 0, st>>>(reinterpret_cast<const float*>(frames), patches, B,
                                                                      H, W, H2, W2);
            else
                bv::stem_patch_kernel<float, 3><<<blocks, 256, 0, st>>>(reinterpret_cast<const float*>(frames), patches, B,
                                                                      H, W, H2, W2);
            BV_CUDA(cudaGetLastError());
            ++launches;
            const double px = (double)B * H2 * W2;
            prof_mark(h, st, "stem_patch_gather", 0, (double)B * C * H * W * (dtype == BV_DTYPE_U8 ? 1 : 4) + px * ((C == 3) ? 192 : 64) * 2);
        }
        // 2. stem GEMM (+bias, ReLU)
        if ((rc = launch_step(h->steps[0], st))) return rc;
        ++launches;
        step_cost(h->steps[0], &pflops, &pbytes, pname, sizeof(pname));
        prof_mark(h, st, pname, pflops, pbytes);
        // 3. max-pool into buf_a
        {
            const long long total = (long long)B * H4 * W4 * 8;
            const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)g_num_sms * 16);
            bv::maxpool3x3s2_nhwc_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(ws + lay.buf_b),
                                                                 reinterpret_cast<__nv_bfloat16*>(ws + lay.buf_a), B, H2,
                                                                 W2, 64, H4, W4);
            BV_CUDA(cudaGetLastError());
            ++launches;
            prof_mark(h, st, "maxpool3x3s2", 0, (double)B * H2 * W2 * 64 * 2 + (double)B * H4 * W4 * 64 * 2);
        }
    }
    // 4. bottleneck convs + projector conv
    for (size_t i = 1; i < h->steps.size(); ++i) {
        if ((rc = launch_step(h->steps[i], st))) return rc;
        ++launches;
        if (h->profile) {
            step_cost(h->steps[i], &pflops, &pbytes, pname, sizeof(pname));
            prof_mark(h, st, pname, pflops, pbytes);
        }
    }
    const int gh = H / 32, gw = W / 32, P = gh * gw;
    // 5. optional trunk outputs
    if (out->pooled) {
        const long long total = (long long)B * (2048 / 8);
        bv::avgpool_nhwc_kernel<<<(int)((total + 255) / 256), 256, 0, st>>>(
            reinterpret_cast<const __nv_bfloat16*>(h->trunk), out->pooled, B, P, 2048);
        BV_CUDA(cudaGetLastError());
        ++launches;
        prof_mark(h, st, "avgpool", 0, (double)B * P * 2048 * 2);
    }
    if (out->trunk_nhwc_bf16) {
        BV_CUDA(cudaMemcpyAsync(out->trunk_nhwc_bf16, h->trunk, (size_t)B * P * 2048 * 2, cudaMemcpyDeviceToDevice, st));
    }
    // 6. projector tail + embeddings + scores
    {
        bv::HeadParams hp{};
        hp.hid = reinterpret_cast<const float*>(ws + lay.hid);
        hp.w2t = h->w.proj3_wt;
        hp.b2 = h->w.proj3_b;
        hp.B = B;
        hp.P = P;
        hp.global_out = out->global_emb;
        hp.patch_out = out->patch_emb;
        hp.normalize_patch = out->normalize_patch;
        const bool want_score = out->sim || out->prob || out->pred || out->score;
        hp.yn = want_score ? h->yn : nullptr;
        hp.L = h->L;
        hp.NPP = h->P;
        hp.score = bv::ScoreOut{out->sim, out->prob, out->pred, out->score};
        hp.heat_out = out->heat;
        hp.heat_t = h->heat_t;
        const size_t smem = bv::kHeadSmemBytes;
        if (!out->patch_emb && !out->heat)
            launch_ex(bv::head_global_kernel, B, 128, 0, st, 1, hp);
        else
            launch_ex(bv::head_kernel, B, 256, smem, st, 1, hp);
        BV_CUDA(cudaGetLastError());
        ++launches;
        prof_mark(h, st, "head_score", 2.0 * B * P * 128 * 128, (double)B * P * 128 * 4);
    }
    h->last_launches = launches;
    return BV_OK;
}
}  // namespace

extern "C" {

int32_t bv_forward(bv_handle* h, const void* frames, int32_t dtype, int32_t B, int32_t C, int32_t H, int32_t W,
                   void* workspace, size_t workspace_bytes, const bv_outputs* out, bv_stream stream) {
    return forward_impl(h, frames, dtype, B, C, H, W, workspace, workspace_bytes, out, stream);
}

int32_t bv_forward_graph(bv_handle* h, const void* frames, int32_t dtype, int32_t B, int32_t C, int32_t H, int32_t W,
                         void* workspace, size_t workspace_bytes, const bv_outputs* out, bv_stream stream) {
    if (!h || !frames || !workspace || !out) return fail(BV_ERR_INVALID, "null argument");
    if (h->profile) return forward_impl(h, frames, dtype, B, C, H, W, workspace, workspace_bytes, out, stream);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    auto up = [](const void* p) { return reinterpret_cast<uintptr_t>(p); };
    const std::vector<uintptr_t> key = {up(frames), (uintptr_t)dtype, (uintptr_t)B, (uintptr_t)C, (uintptr_t)H, (uintptr_t)W,
                                        up(workspace), (uintptr_t)workspace_bytes, up(out->global_emb), up(out->patch_emb),
                                        (uintptr_t)out->normalize_patch, up(out->pooled), up(out->trunk_nhwc_bf16),
                                        up(out->sim), up(out->prob), up(out->pred), up(out->score), up(out->heat),
                                        up(h->yn), up(h->heat_t), (uintptr_t)h->L, (uintptr_t)h->P};
    for (auto& g : h->graphs) {
        if (g.key == key) {
            BV_CUDA(cudaGraphLaunch(g.exec, st));
            h->last_launches = g.launches;
            return BV_OK;
        }
    }
    // First use of this argument set: record the launches of one forward into a graph (nothing executes while the
    // stream is capturing), instantiate it and launch it.  Tensor maps and kernel parameters are built on the host
    // exactly as for a direct call; the capture only sees kernel / memcpy nodes and their programmatic-launch edges.
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    BV_CUDA(cudaStreamIsCapturing(st, &cs));
    if (cs != cudaStreamCaptureStatusNone)   // the caller is capturing already: just add our nodes to its graph
        return forward_impl(h, frames, dtype, B, C, H, W, workspace, workspace_bytes, out, stream);
    // The capture runs on a private stream: the caller's stream may be the legacy default stream (torch's default),
    // which cannot be captured.  Nothing executes during capture; the instantiated graph is launched on the caller's stream.
    if (!h->capture_stream) BV_CUDA(cudaStreamCreateWithFlags(&h->capture_stream, cudaStreamNonBlocking));
    BV_CUDA(cudaStreamBeginCapture(h->capture_stream, cudaStreamCaptureModeThreadLocal));
    const int rc = forward_impl(h, frames, dtype, B, C, H, W, workspace, workspace_bytes, out, h->capture_stream);
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(h->capture_stream, &graph);
    if (rc != BV_OK) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        return rc;
    }
    if (ce != cudaSuccess || !graph) {
        cudaGetLastError();
        return fail(BV_ERR_CUDA, "stream capture of the forward failed: %s", cudaGetErrorString(ce));
    }
    cudaGraphExec_t exec = nullptr;
    const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) return fail(BV_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(ie));
    if (h->graphs.size() >= 8) {   // bounded cache: drop the oldest entry
        cudaGraphExecDestroy(h->graphs.front().exec);
        h->graphs.erase(h->graphs.begin());
    }
    h->graphs.push_back({key, exec, h->last_launches});
    BV_CUDA(cudaGraphLaunch(exec, st));
    return BV_OK;
}

int32_t bv_jpeg_info(const uint8_t* host_data, size_t length, int32_t* width, int32_t* height, int32_t* components) {
    if (!host_data || length == 0 || !width || !height) return fail(BV_ERR_INVALID, "bad JPEG arguments");
    int rc = device_setup();
    if (rc) return rc;
    const char* why = nullptr;
    const jpeg_stage::Api* a = jpeg_stage::api(&why);
    if (!a) return fail(BV_ERR_CUDA, "nvJPEG unavailable: %s", why);
    int dev = 0;
    BV_CUDA(cudaGetDevice(&dev));
    jpeg_stage::Decoder* d = jpeg_stage::decoder(dev, a, &why);
    if (!d) return fail(BV_ERR_CUDA, "nvJPEG: %s", why);
    int ncomp = 0, ws[NVJPEG_MAX_COMPONENT] = {0}, hs[NVJPEG_MAX_COMPONENT] = {0};
    nvjpegChromaSubsampling_t ss;
    const nvjpegStatus_t st = a->GetImageInfo(d->handle, host_data, length, &ncomp, &ss, ws, hs);
    if (st != NVJPEG_STATUS_SUCCESS) return fail(BV_ERR_INVALID, "not a decodable JPEG stream (nvjpegGetImageInfo status %d)", (int)st);
    *width = ws[0];
    *height = hs[0];
    if (components) *components = ncomp;
    return BV_OK;
}

int32_t bv_jpeg_decode_gray_u8(const uint8_t* host_data, size_t length, uint8_t* out, int32_t width, int32_t height,
                               int32_t pitch, bv_stream stream) {
    if (!host_data || length == 0 || !out || width <= 0 || height <= 0 || pitch < width)
        return fail(BV_ERR_INVALID, "bad JPEG arguments");
    int rc = device_setup();
    if (rc) return rc;
    const char* why = nullptr;
    const jpeg_stage::Api* a = jpeg_stage::api(&why);
    if (!a) return fail(BV_ERR_CUDA, "nvJPEG unavailable: %s", why);
    int dev = 0;
    BV_CUDA(cudaGetDevice(&dev));
    jpeg_stage::Decoder* d = jpeg_stage::decoder(dev, a, &why);
    if (!d) return fail(BV_ERR_CUDA, "nvJPEG: %s", why);
    int ncomp = 0, ws[NVJPEG_MAX_COMPONENT] = {0}, hs[NVJPEG_MAX_COMPONENT] = {0};
    nvjpegChromaSubsampling_t ss;
    nvjpegStatus_t st = a->GetImageInfo(d->handle, host_data, length, &ncomp, &ss, ws, hs);
    if (st != NVJPEG_STATUS_SUCCESS) return fail(BV_ERR_INVALID, "not a decodable JPEG stream (status %d)", (int)st);
    if (ws[0] != width || hs[0] != height)
        return fail(BV_ERR_INVALID, "JPEG is %dx%d but the output buffer was sized for %dx%d", ws[0], hs[0], width, height);
    nvjpegImage_t img{};
    img.channel[0] = out;                    // NVJPEG_OUTPUT_Y: the luma plane only (a grey JPEG's single component)
    img.pitch[0] = static_cast<size_t>(pitch);
    st = a->Decode(d->handle, d->state, host_data, length, NVJPEG_OUTPUT_Y, &img, reinterpret_cast<cudaStream_t>(stream));
    if (st != NVJPEG_STATUS_SUCCESS) return fail(BV_ERR_CUDA, "nvjpegDecode failed (status %d)", (int)st);
    return BV_OK;
}

int32_t bv_jpeg_decode_batch_gray_u8(const uint8_t* const* host_datas, const size_t* lengths, int32_t n, uint8_t* const* outs,
                                     const int32_t* pitches, int32_t backend, bv_stream stream) {
    if (!host_datas || !lengths || !outs || !pitches || n <= 0) return fail(BV_ERR_INVALID, "bad JPEG batch arguments");
    int rc = device_setup();
    if (rc) return rc;
    const char* why = nullptr;
    const jpeg_stage::Api* a = jpeg_stage::api(&why);
    if (!a) return fail(BV_ERR_CUDA, "nvJPEG unavailable: %s", why);
    int dev = 0;
    BV_CUDA(cudaGetDevice(&dev));
    // backend: 3 = hardware JPEG engines, 2 = GPU-assisted Huffman decode; 0 = try 3, then 2
    const int order[2] = {backend == 0 ? 3 : backend, backend == 0 ? 2 : -1};
    for (int oi = 0; oi < 2; ++oi) {
        if (order[oi] < 0) break;
        jpeg_stage::BatchDecoder* d = jpeg_stage::batch_decoder(dev, order[oi], a);
        if (!d) continue;
        if (d->batch != n) {
            if (a->BatchedInitialize(d->handle, d->state, n, 1, NVJPEG_OUTPUT_Y) != NVJPEG_STATUS_SUCCESS) continue;
            d->batch = n;
        }
        std::vector<nvjpegImage_t> imgs(n);
        for (int i = 0; i < n; ++i) {
            memset(&imgs[i], 0, sizeof(nvjpegImage_t));
            imgs[i].channel[0] = outs[i];
            imgs[i].pitch[0] = static_cast<size_t>(pitches[i]);
        }
        const nvjpegStatus_t st = a->Batched(d->handle, d->state, host_datas, lengths, imgs.data(), reinterpret_cast<cudaStream_t>(stream));
        if (st == NVJPEG_STATUS_SUCCESS) return order[oi];
        d->batch = 0;   // the state may be unusable for this batch shape: re-initialise next time
        cudaGetLastError();
    }
    return fail(BV_ERR_INVALID, "no batched nvJPEG backend could decode this batch (use the per-image path)");
}

int32_t bv_quantize_frames_f32(const float* x, int32_t B, int32_t C, int32_t H, int32_t W, uint8_t* out, int32_t* bad,
                               bv_stream stream) {
    if (!x || !out || !bad || B <= 0 || (C != 1 && C != 3) || H <= 0 || W <= 0) return fail(BV_ERR_INVALID, "bad frame arguments");
    if (((long long)H * W) % 4 != 0) return fail(BV_ERR_INVALID, "H * W must be a multiple of 4");
    if ((reinterpret_cast<uintptr_t>(x) & 15u) || (reinterpret_cast<uintptr_t>(out) & 3u))
        return fail(BV_ERR_INVALID, "frames must be 16-byte aligned, the 8-bit output 4-byte aligned");
    int rc = device_setup();
    if (rc) return rc;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    BV_CUDA(cudaMemsetAsync(bad, 0, sizeof(int32_t), st));
    const long long hw4 = (long long)H * W / 4, total4 = hw4 * B;
    const int blocks = (int)std::min<long long>((total4 + 255) / 256, (long long)g_num_sms * 16);
    bv::quantize_frames_kernel<<<blocks, 256, 0, st>>>(x, out, C, hw4, total4, bad);
    BV_CUDA(cudaGetLastError());
    return BV_OK;
}

int32_t bv_conv2d_nhwc(const void* x, int32_t B, int32_t H, int32_t W, const bv_conv* c, const void* x2, int32_t H2,
                       int32_t W2, const bv_conv* c2, const void* residual, int32_t relu, void* out, int32_t out_fp32,
                       bv_stream stream) {
    if (!x || !c || !out) return fail(BV_ERR_INVALID, "null argument");
    int rc = device_setup();
    if (rc) return rc;
    ConvOperand ops[2];
    ops[0] = ConvOperand{x, H, W, *c};
    int nops = 1;
    if (x2 && c2) {
        ops[1] = ConvOperand{x2, H2, W2, *c2};
        nops = 2;
    }
    ConvLaunch L;
    rc = build_conv(&L, B, ops, nops, residual, relu, out, out_fp32);
    if (rc) return rc;
    return launch_conv(L, reinterpret_cast<cudaStream_t>(stream));
}

// ---------------------------------------------------------------------------------------------------------------
// Resize + CenterCrop (resize.cuh)
// ---------------------------------------------------------------------------------------------------------------
}  // extern "C"

namespace {

struct ResampleTable {
    int ksize = 0;
    std::vector<int> bounds;  // [out][2]
    std::vector<int> kk;      // [out][ksize]
    int* d_bounds = nullptr;
    int* d_kk = nullptr;
};

// Pillow Resample.c precompute_coeffs (BILINEAR, support 1.0, box = the whole axis) + normalize_coeffs_8bpc.
void build_resample_table(int in_size, int out_size, ResampleTable* t) {
    double scale, filterscale;
    filterscale = scale = (double)in_size / out_size;
    if (filterscale < 1.0) filterscale = 1.0;
    const double support = 1.0 * filterscale;
    const int ksize = (int)ceil(support) * 2 + 1;
    t->ksize = ksize;
    t->bounds.assign((size_t)out_size * 2, 0);
    t->kk.assign((size_t)out_size * ksize, 0);
    std::vector<double> k(ksize);
    const double ss = 1.0 / filterscale;
    for (int xx = 0; xx < out_size; ++xx) {
        const double center = 0.0 + (xx + 0.5) * scale;
        double ww = 0.0;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        for (int x = 0; x < ksize; ++x) k[x] = 0.0;
        for (int x = 0; x < xmax; ++x) {
            double a = (x + xmin - center + 0.5) * ss;
            if (a < 0.0) a = -a;
            const double w = (a < 1.0) ? 1.0 - a : 0.0;
            k[x] = w;
            ww += w;
        }
        for (int x = 0; x < xmax; ++x)
            if (ww != 0.0) k[x] /= ww;
        for (int x = 0; x < ksize; ++x) {
            const double v = k[x] * (double)(1 << bv::kResamplePrecisionBits);
            t->kk[(size_t)xx * ksize + x] = (k[x] < 0) ? (int)(-0.5 + v) : (int)(0.5 + v);
        }
        t->bounds[2 * xx] = xmin;
        t->bounds[2 * xx + 1] = xmax;
    }
}

std::mutex g_resample_mutex;
std::map<std::pair<int, int>, ResampleTable> g_resample_tables[16];  // per device

// Device copy of the table for (in_size -> out_size), built once per process and device.
int get_resample_table(int in_size, int out_size, const ResampleTable** out) {
    int dev = 0;
    BV_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 16) return fail(BV_ERR_INVALID, "device index %d out of range", dev);
    std::lock_guard<std::mutex> lock(g_resample_mutex);
    auto& cache = g_resample_tables[dev];
    auto it = cache.find({in_size, out_size});
    if (it == cache.end()) {
        ResampleTable t;
        build_resample_table(in_size, out_size, &t);
        BV_CUDA(cudaMalloc(&t.d_bounds, t.bounds.size() * sizeof(int)));
        BV_CUDA(cudaMalloc(&t.d_kk, t.kk.size() * sizeof(int)));
        BV_CUDA(cudaMemcpy(t.d_bounds, t.bounds.data(), t.bounds.size() * sizeof(int), cudaMemcpyHostToDevice));
        BV_CUDA(cudaMemcpy(t.d_kk, t.kk.data(), t.kk.size() * sizeof(int), cudaMemcpyHostToDevice));
        it = cache.emplace(std::make_pair(in_size, out_size), std::move(t)).first;
    }
    *out = &it->second;
    return BV_OK;
}

// torchvision _compute_resized_output_size (int size) and center_crop offsets
void resized_size(int h, int w, int size, int* nh, int* nw) {
    const int short_side = (w <= h) ? w : h, long_side = (w <= h) ? h : w;
    const int new_long = (int)((double)size * long_side / short_side);
    if (w <= h) { *nw = size; *nh = new_long; } else { *nh = size; *nw = new_long; }
}
int py_round_half(int num) {  // int(round(num / 2.0)) with Python's round-half-to-even
    if (num % 2 == 0) return num / 2;
    const int lo = (num - 1) / 2;  // num >= 0 here
    return (lo % 2 == 0) ? lo : lo + 1;
}

struct ResizePlan {
    int nh, nw, top, left;
    bool need_h, need_v;
    int row_first, row_count;  // source rows the horizontal pass must produce for the cropped vertical outputs
    const ResampleTable* th = nullptr;
    const ResampleTable* tv = nullptr;
};

int plan_resize(int h, int w, int size, int crop, bool tables, ResizePlan* pl) {
    if (h <= 0 || w <= 0 || size <= 0 || crop <= 0) return fail(BV_ERR_INVALID, "bad resize arguments");
    resized_size(h, w, size, &pl->nh, &pl->nw);
    if (pl->nh < crop || pl->nw < crop)
        return fail(BV_ERR_INVALID, "crop %d exceeds the resized frame %dx%d (the reference never pads)", crop, pl->nh, pl->nw);
    pl->top = py_round_half(pl->nh - crop);
    pl->left = py_round_half(pl->nw - crop);
    pl->need_h = pl->nw != w;
    pl->need_v = pl->nh != h;
    pl->row_first = pl->need_v ? 0 : pl->top;
    pl->row_count = pl->need_v ? h : crop;
    if (pl->need_v) {
        // rows touched by the cropped vertical outputs: [bounds[top].first, bounds[top+crop-1].first + count)
        ResampleTable tmp;
        const ResampleTable* tv = &tmp;
        if (tables) {
            int rc = get_resample_table(h, pl->nh, &pl->tv);
            if (rc) return rc;
            tv = pl->tv;
        } else {
            build_resample_table(h, pl->nh, &tmp);
        }
        pl->row_first = tv->bounds[2 * pl->top];
        const int last = pl->top + crop - 1;
        pl->row_count = tv->bounds[2 * last] + tv->bounds[2 * last + 1] - pl->row_first;
    }
    if (pl->need_h && tables) {
        int rc = get_resample_table(w, pl->nw, &pl->th);
        if (rc) return rc;
    }
    return BV_OK;
}

}  // namespace

extern "C" {

size_t bv_resize_workspace_bytes(int32_t n, int32_t h, int32_t w, int32_t size, int32_t crop) {
    ResizePlan pl;
    if (n <= 0 || plan_resize(h, w, size, crop, false, &pl) != BV_OK) return 0;
    if (!(pl.need_h && pl.need_v)) return 256;  // a single pass writes straight to the output
    return align_up((size_t)n * pl.row_count * crop, 256) + 256;
}

int32_t bv_resize_center_crop_u8(const uint8_t* src, int32_t n, int32_t h, int32_t w, int32_t size, int32_t crop,
                                 uint8_t* out, void* workspace, size_t workspace_bytes, bv_stream stream) {
    if (!src || !out || n <= 0) return fail(BV_ERR_INVALID, "bad resize arguments");
    int rc = device_setup();
    if (rc) return rc;
    ResizePlan pl;
    if ((rc = plan_resize(h, w, size, crop, true, &pl))) return rc;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    auto blocks_for = [](long long total) { return (int)std::min<long long>((total + 255) / 256, (long long)g_num_sms * 32); };
    if (!pl.need_h && !pl.need_v) {
        bv::crop_copy_kernel<<<blocks_for((long long)n * crop * crop), 256, 0, st>>>(src, out, n, h, w, pl.top, pl.left, crop);
        BV_CUDA(cudaGetLastError());
        return BV_OK;
    }
    const uint8_t* vsrc = src;     // what the vertical pass reads
    int v_rows = h, v_pitch = w, v_col0 = pl.left, v_origin = 0;
    if (pl.need_h) {
        bv::ResamplePass ph{};
        ph.src = src;
        ph.bounds = pl.th->d_bounds;
        ph.kk = pl.th->d_kk;
        ph.ksize = pl.th->ksize;
        ph.n = n;
        ph.src_rows = h;
        ph.src_pitch = w;
        ph.dst_rows = pl.row_count;
        ph.dst_cols = crop;
        ph.dst_pitch = crop;
        ph.out0 = pl.left;
        ph.other0 = pl.row_first;
        ph.tap_origin = 0;
        if (pl.need_v) {
            const size_t need = (size_t)n * pl.row_count * crop;
            if (!workspace || workspace_bytes < need)
                return fail(BV_ERR_WORKSPACE, "resize workspace too small: %zu < %zu", workspace_bytes, need);
            ph.dst = reinterpret_cast<uint8_t*>(workspace);
        } else {
            ph.dst = out;          // rows [top, top + crop) of the source, columns resampled
        }
        bv::resample_horizontal_kernel<<<blocks_for((long long)n * pl.row_count * crop), 256, 0, st>>>(ph);
        BV_CUDA(cudaGetLastError());
        vsrc = ph.dst;
        v_rows = pl.row_count;
        v_pitch = crop;
        v_col0 = 0;
        v_origin = pl.row_first;
    }
    if (pl.need_v) {
        bv::ResamplePass pv{};
        pv.src = vsrc;
        pv.dst = out;
        pv.bounds = pl.tv->d_bounds;
        pv.kk = pl.tv->d_kk;
        pv.ksize = pl.tv->ksize;
        pv.n = n;
        pv.src_rows = v_rows;
        pv.src_pitch = v_pitch;
        pv.dst_rows = crop;
        pv.dst_cols = crop;
        pv.dst_pitch = crop;
        pv.out0 = pl.top;
        pv.other0 = v_col0;
        pv.tap_origin = v_origin;
        bv::resample_vertical_kernel<<<blocks_for((long long)n * crop * crop), 256, 0, st>>>(pv);
        BV_CUDA(cudaGetLastError());
    }
    return BV_OK;
}

int32_t bv_smooth_heatmaps(const float* heat, int32_t B, int32_t gh, int32_t gw, int32_t L, float sigma, float* out,
                           bv_stream stream) {
    if (!heat || !out || B <= 0 || L <= 0 || gh <= 0 || gw <= 0) return fail(BV_ERR_INVALID, "bad heat-map arguments");
    if (gh * gw > bv::kSmoothMaxCells) return fail(BV_ERR_INVALID, "patch grid %dx%d too large (max %d cells)", gh, gw, bv::kSmoothMaxCells);
    if (!(sigma > 0.f)) return fail(BV_ERR_INVALID, "sigma must be positive");
    int rc = device_setup();
    if (rc) return rc;
    bv::SmoothParams p{};
    // scipy.ndimage._filters: lw = int(truncate * sd + 0.5); phi = exp(-0.5 / sd^2 * x^2); phi /= phi.sum()
    const double sd = (double)sigma;
    const int radius = (int)(4.0 * sd + 0.5);
    if (radius > bv::kSmoothMaxRadius) return fail(BV_ERR_INVALID, "sigma %.3f needs radius %d > %d", sigma, radius, bv::kSmoothMaxRadius);
    double w[2 * bv::kSmoothMaxRadius + 1], sum = 0;
    for (int k = -radius; k <= radius; ++k) sum += (w[k + radius] = exp(-0.5 / (sd * sd) * (double)k * k));
    for (int k = 0; k <= 2 * radius; ++k) p.w[k] = (float)(w[k] / sum);
    p.heat = heat;
    p.out = out;
    p.B = B; p.gh = gh; p.gw = gw; p.L = L;
    p.radius = radius;
    bv::heat_smooth_kernel<<<B * L, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    BV_CUDA(cudaGetLastError());
    return BV_OK;
}

int32_t bv_heatmaps_to_image_size(const float* heat, int32_t B, int32_t gh, int32_t gw, int32_t L, int32_t height,
                                  int32_t width, int32_t resize_size, int32_t crop_size, float* out, bv_stream stream) {
    if (!heat || !out || B <= 0 || L <= 0 || gh <= 0 || gw <= 0 || height <= 0 || width <= 0 || resize_size < 0 || crop_size < 0)
        return fail(BV_ERR_INVALID, "bad heat-map arguments");
    if ((long long)B * L > 65535) return fail(BV_ERR_INVALID, "more than 65535 maps in one call");
    int rc = device_setup();
    if (rc) return rc;
    // vlp/inference_engine.py:133-154: with a crop the grid covers a square of `side` original pixels, centred
    int side_h = height, side_w = width, top = 0, left = 0;
    if (crop_size > 0) {
        const int smallest = std::min(height, width);
        const int side = resize_size > 0 ? (int)((double)crop_size * smallest / resize_size) : crop_size;   // int(crop * min / resize)
        if (side <= 0) return fail(BV_ERR_INVALID, "the crop covers no pixel of a %dx%d image", height, width);
        side_h = side_w = side;
        left = (int)std::floor((width - side) / 2.0);     // F.pad margins: (floor(mw / 2), ceil(mw / 2), floor(mh / 2), ceil(mh / 2))
        top = (int)std::floor((height - side) / 2.0);
    }
    const int blocks_x = std::min((height * width + 255) / 256, 64);
    bv::heat_to_image_kernel<<<dim3(blocks_x, B * L), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        heat, out, gh, gw, L, height, width, side_h, side_w, top, left);
    BV_CUDA(cudaGetLastError());
    return BV_OK;
}

int32_t bv_l1_block_nhwc(const void* t1, int32_t B, int32_t H, int32_t W, const bv_conv* c2, const bv_conv* c3,
                         const void* residual, void* out1, const bv_conv* next, void* out2, bv_stream stream) {
    if (!t1 || !c2 || !c3 || !residual || !out1 || !next || !out2) return fail(BV_ERR_INVALID, "null argument");
    int rc = device_setup();
    if (rc) return rc;
    L1Launch L;
    if ((rc = build_l1_block(&L, B, H, W, t1, *c2, *c3, residual, out1, *next, out2))) return rc;
    return launch_l1_block(L, reinterpret_cast<cudaStream_t>(stream));
}

int32_t bv_pair_chain_nhwc(const void* x, int32_t B, int32_t H, int32_t W, const bv_conv* c, const void* residual, void* out1,
                           const bv_conv* next, void* out2, bv_stream stream) {
    if (!x || !c || !residual || !out1 || !next || !out2) return fail(BV_ERR_INVALID, "null argument");
    int rc = device_setup();
    if (rc) return rc;
    PairChainLaunch L;
    if ((rc = build_pair_chain(&L, B, H, W, x, *c, residual, out1, *next, out2))) return rc;
    return launch_pair_chain(L, reinterpret_cast<cudaStream_t>(stream));
}

int32_t bv_l1_block_ds_nhwc(const void* t1, int32_t B, int32_t H, int32_t W, const bv_conv* c2, const bv_conv* c3,
                            const void* x0, const bv_conv* ds, void* out1, const bv_conv* next, void* out2, bv_stream stream) {
    if (!t1 || !c2 || !c3 || !x0 || !ds || !out1 || !next || !out2) return fail(BV_ERR_INVALID, "null argument");
    int rc = device_setup();
    if (rc) return rc;
    L1Launch L;
    if ((rc = build_l1_block(&L, B, H, W, t1, *c2, *c3, nullptr, out1, *next, out2, x0, ds))) return rc;
    return launch_l1_block(L, reinterpret_cast<cudaStream_t>(stream));
}

int32_t bv_stem_u8_nhwc(const void* frames, int32_t B, int32_t H, int32_t W, const bv_conv* w8, void* out,
                        int32_t variant, bv_stream stream) {
    if (!frames || !w8 || !w8->w || !w8->bias || !out) return fail(BV_ERR_INVALID, "null argument");
    if (B <= 0 || H <= 0 || W <= 0 || H % 32 || W % 32) return fail(BV_ERR_INVALID, "H, W must be multiples of 32");
    int rc = device_setup();
    if (rc) return rc;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (variant == 0) {
        if (reinterpret_cast<uintptr_t>(frames) & 7u) return fail(BV_ERR_INVALID, "frames must be 8-byte aligned");
        bv::StemRowsParams p{};
        p.frames = reinterpret_cast<const uint8_t*>(frames);
        p.w = reinterpret_cast<const __nv_bfloat16*>(w8->w);
        p.out = reinterpret_cast<__nv_bfloat16*>(out);
        p.B = B;
        p.H = H;
        p.W = W;
        p.strips_x = (W / 4 + bv::kSrStripPx - 1) / bv::kSrStripPx;
        if (const char* d = getenv("BV_SR_DEBUG")) p.debug = atoi(d);
        const int grid = std::min(B * p.strips_x, g_num_sms);
        launch_stem_rows(p, grid, st);
    } else if (variant == 1) {
        bv::StemParams p{};
        if ((rc = make_tmap_2d(&p.tmW, w8->w, 64, 64, 64, 64))) return rc;
        p.frames = reinterpret_cast<const uint8_t*>(frames);
        p.bias = w8->bias;
        p.out = reinterpret_cast<__nv_bfloat16*>(out);
        p.B = B;
        p.H = H;
        p.W = W;
        p.tiles_x = (W / 4) / bv::kStemPool;
        p.tiles_y = (H / 4) / bv::kStemPool;
        const int grid = std::min(B * p.tiles_x * p.tiles_y, 2 * g_num_sms);
        bv::stem_fused_kernel<<<grid, bv::kStemThreads, bv::kStemSmemRequest, st>>>(p);
    } else {
        return fail(BV_ERR_INVALID, "unknown stem variant %d", variant);
    }
    BV_CUDA(cudaGetLastError());
    return BV_OK;
}

int32_t bv_stem_conv1_u8_nhwc(const void* frames, int32_t B, int32_t H, int32_t W, const bv_conv* w8, const bv_conv* c1,
                              void* out, void* out1, bv_stream stream) {
    if (!frames || !w8 || !w8->w || !c1 || !c1->w || !c1->bias || !out || !out1) return fail(BV_ERR_INVALID, "null argument");
    if (B <= 0 || H <= 0 || W <= 0 || H % 32 || W % 32) return fail(BV_ERR_INVALID, "H, W must be multiples of 32");
    if (c1->cin != 64 || c1->cout != 64 || c1->r != 1 || c1->s != 1 || c1->stride != 1 || c1->pad != 0)
        return fail(BV_ERR_INVALID, "the fused conv1 must be 1x1, 64 -> 64, stride 1");
    if (reinterpret_cast<uintptr_t>(frames) & 7u) return fail(BV_ERR_INVALID, "frames must be 8-byte aligned");
    int rc = device_setup();
    if (rc) return rc;
    bv::StemRowsParams p{};
    p.frames = reinterpret_cast<const uint8_t*>(frames);
    p.w = reinterpret_cast<const __nv_bfloat16*>(w8->w);
    p.out = reinterpret_cast<__nv_bfloat16*>(out);
    p.B = B;
    p.H = H;
    p.W = W;
    p.strips_x = (W / 4 + bv::kSrStripPx - 1) / bv::kSrStripPx;
    p.w1 = reinterpret_cast<const __nv_bfloat16*>(c1->w);
    p.b1 = c1->bias;
    p.out1 = reinterpret_cast<__nv_bfloat16*>(out1);
    if (const char* d = getenv("BV_SR_DEBUG")) p.debug = atoi(d);
    launch_stem_rows(p, std::min(B * p.strips_x, g_num_sms), reinterpret_cast<cudaStream_t>(stream));
    BV_CUDA(cudaGetLastError());
    return BV_OK;
}

int32_t bv_pair_gemm_test(const void* a, const void* w, int32_t M, int32_t N, int32_t K, float* out, bv_stream stream) {
    if (!a || !w || !out) return fail(BV_ERR_INVALID, "null argument");
    if (M <= 0 || N < 32 || N > 256 || N % 32 || K <= 0 || K % 64) return fail(BV_ERR_INVALID, "unsupported pair GEMM shape");
    int rc = device_setup();
    if (rc) return rc;
    bv::PairGemmParams p{};
    if ((rc = make_tmap_2d(&p.tmA, a, (uint64_t)K, (uint64_t)M, 64, 128))) return rc;
    if ((rc = make_tmap_2d(&p.tmB, w, (uint64_t)K, (uint64_t)N, 64, (uint32_t)(N / 2)))) return rc;
    p.out = out;
    p.M = M;
    p.N = N;
    p.K = K;
    p.num_pair_tiles = (M + 255) / 256;
    cudaLaunchConfig_t cfg{};
    const int pairs = std::min(p.num_pair_tiles, g_num_sms / 2);
    cfg.gridDim = dim3(2 * pairs, 1, 1);
    cfg.blockDim = dim3(bv::kPairThreads, 1, 1);
    cfg.dynamicSmemBytes = bv::kPairSmemBytes;
    cfg.stream = reinterpret_cast<cudaStream_t>(stream);
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    BV_CUDA(cudaLaunchKernelEx(&cfg, bv::pair_gemm_kernel, p));
    return BV_OK;
}

int32_t bv_conv_chain_nhwc(const void* x, int32_t B, int32_t H, int32_t W, const bv_conv* c, const void* x2, int32_t H2,
                           int32_t W2, const bv_conv* c2, const void* residual, void* out1, const bv_conv* next,
                           void* out2, bv_stream stream) {
    if (!x || !c || !out1 || !next || !out2) return fail(BV_ERR_INVALID, "null argument");
    int rc = device_setup();
    if (rc) return rc;
    ConvOperand ops[2];
    ops[0] = ConvOperand{x, H, W, *c};
    int nops = 1;
    if (x2 && c2) {
        ops[1] = ConvOperand{x2, H2, W2, *c2};
        nops = 2;
    }
    ChainLaunch L;
    rc = build_chain(&L, B, ops, nops, residual, out1, *next, out2);
    if (rc) return rc;
    return launch_chain(L, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
