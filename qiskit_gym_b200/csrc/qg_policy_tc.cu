// qg_policy_tc.cu — the policy network for LARGE batches (the rollout collector at 65 536 environments) on the 5th-generation tensor
// cores: tcgen05.mma with accumulators in tensor memory, operands streamed into shared memory by bulk copies (cp.async.bulk, the TMA
// unit's 1-D form) that complete on mbarriers, tcgen05.ld epilogues.  SURVEY.md §8f row 3 at collector scale; the small-batch search
// keeps the f32 FFMA kernel of qg_policy_kernels.cuh (8 rows per CTA is latency-, not throughput-bound).
//
// Arithmetic.  A twisterl BasicPolicy is Linear -> ReLU chains in f32.  To stay within the 1e-4 logit tolerance of the f32 module on
// f16 tensor-core inputs, every f32 value v is carried as two halves  v ~ hi + lo,  hi = half(v), lo = half(v - hi)  (22 significant bits) and
// a product is evaluated as  a_hi*w_hi + a_hi*w_lo + a_lo*w_hi  with f32 accumulation in tensor memory (the dropped lo*lo term is 2^-22
// relative).  The first layer's input is 0/1 — exact in one half — so it needs two products, the others three.
//
// One kernel per layer:  Y[128-row tile][NT columns] = act(X W^T + b),  persistent CTAs looping over (row tile, column tile) pairs,
//   warp 0  loader  : one lane issues the stage's bulk copies (X tile images written by the previous layer's epilogue, W tile images laid
//                     out once at creation), 3 stages in flight, full / empty mbarriers;
//   warp 1  MMA     : one lane issues  tcgen05.mma.cta_group::1.kind::f16  M = 128, N = NT, K = 16  over the stage's 64-wide K block and
//                     commits the stage back to the loader (tcgen05.commit -> empty barrier) and, after the last K block, the accumulator to
//                     the epilogue; two accumulator buffers in tensor memory, so tile i+1's products overlap tile i's epilogue;
//   warps 2-5 epilogue: thread = output row (tensor-memory lane): tcgen05.ld 16 columns at a time, + bias, ReLU, split into halves, written as
//                     the next layer's X tile image (coalesced 16-byte stores); the last layer writes logits / softmax / value instead.
// Operand images use the no-swizzle K-major canonical layout  [k / 8][row][k % 8]  (8-row x 16-byte core matrices contiguous; leading
// byte offset = rows * 16 between the two 8-wide K chunks of one MMA, stride byte offset = 128 between 8-row groups), which is also the
// layout in which one epilogue thread per row writes 16 bytes per 8 columns with consecutive threads at consecutive addresses.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "qg_host.hpp"

namespace qg {
namespace tc {

constexpr int kM = 128;              // rows per tile (UMMA M, tensor-memory lanes)
constexpr int kKB = 64;              // K block (halves) per pipeline stage
constexpr int kStages = 3;
constexpr int kMaxNT = 128;          // widest column tile
constexpr int kThreads = 192;        // 6 warps: loader, MMA, 4 epilogue
constexpr int kImgA = kM * kKB;      // halves per X tile image (16 KB)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n .reg .pred p;\n"
        "TC_WAIT_LOOP:\n"
        " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        " @p bra TC_WAIT_DONE;\n"
        " bra TC_WAIT_LOOP;\n"
        "TC_WAIT_DONE:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* sdst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(sdst)), "l"(gsrc), "r"(bytes),
                 "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
// D[tmem] (+)= A[smem] * B[smem]^T, f16 inputs, f32 accumulate
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n .reg .pred p;\n"
        " setp.ne.b32 p, %4, 0;\n"
        " tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// all MMAs issued so far by this thread arrive on the mbarrier when they have completed (implies tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                   "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}
// shared-memory matrix descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor: start address, leading / stride byte offsets in 16-byte
// units, version 1 = sm_100, layout type 0)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): f32 accumulator, f16 A and B, both K-major, M = 128, N = n
__host__ __device__ constexpr uint32_t instr_desc(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kM >> 4) << 24); }

struct LayerArgs {
    const __half* X;         // input tile images  [m_tiles][Kb][a_terms][8][128][8]
    const __half* W;         // weight tile images [2 terms][n_tiles][Kb][8][NT][8]
    const float* bias;       // [n_tiles * NT]
    __half* Y;               // next layer's X images [m_tiles][out_Kb][2][8][128][8], or null (last layer)
    float* logits; float* probs; float* values;      // last layer outputs (each may be null)
    int a_terms, Kb, NT, n_tiles, m_tiles, out_Kb, relu, num_actions, has_value;
    long long batch;
};

struct SmemPlan { uint32_t a_off[kStages][2], b_off[kStages][2], bar_off, total; };
__host__ __device__ inline SmemPlan plan_smem(int a_terms, int NT) {
    SmemPlan p{}; uint32_t o = 0;
    for (int s = 0; s < kStages; ++s) {
        for (int t = 0; t < 2; ++t) { p.a_off[s][t] = o; if (t < a_terms) o += kImgA * 2; }
        for (int t = 0; t < 2; ++t) { p.b_off[s][t] = o; o += (uint32_t)NT * kKB * 2; }
    }
    p.bar_off = o; o += 128;
    p.total = o;
    return p;
}

__global__ void __launch_bounds__(kThreads, 1) k_tc_layer(const __grid_constant__ LayerArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const SmemPlan sp = plan_smem(a.a_terms, a.NT);
    uint64_t* const full = reinterpret_cast<uint64_t*>(smem + sp.bar_off);      // [kStages]
    uint64_t* const empty = full + kStages;                                     // [kStages]
    uint64_t* const acc_full = empty + kStages;                                 // [2]
    uint64_t* const acc_empty = acc_full + 2;                                   // [2]
    uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_tiles = a.m_tiles * a.n_tiles;
    const uint32_t acc_cols = 128;                                              // columns per accumulator buffer (NT <= 128)

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(acc_full + s, 1); mbar_init(acc_empty + s, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 1) {            // one warp allocates the CTA's tensor memory: 2 accumulator buffers of 128 columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(2 * acc_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== loader =====
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            const uint32_t bytes = (uint32_t)a.a_terms * kImgA * 2 + 2u * (uint32_t)a.NT * kKB * 2;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                const int mt = t / a.n_tiles, nt = t - mt * a.n_tiles;
                for (int kb = 0; kb < a.Kb; ++kb) {
                    mbar_wait(empty + stage, phase ^ 1u);
                    mbar_expect_tx(full + stage, bytes);
                    for (int term = 0; term < a.a_terms; ++term)
                        bulk_g2s(smem + sp.a_off[stage][term], a.X + (((size_t)mt * a.Kb + kb) * a.a_terms + term) * kImgA, kImgA * 2, full + stage);
                    for (int term = 0; term < 2; ++term)
                        bulk_g2s(smem + sp.b_off[stage][term], a.W + (((size_t)term * a.n_tiles + nt) * a.Kb + kb) * ((size_t)a.NT * kKB), (uint32_t)a.NT * kKB * 2, full + stage);
                    if (++stage == kStages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            uint32_t stage = 0, phase = 0, as = 0, aphase = 0;
            const uint32_t idesc = instr_desc(a.NT);
            const uint32_t lbo_a = kM * 16, lbo_b = (uint32_t)a.NT * 16, sbo = 128;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                mbar_wait(acc_empty + as, aphase ^ 1u);             // the epilogue has drained this accumulator buffer
                tc_fence_after();
                const uint32_t d = tmem_base + as * acc_cols;
                for (int kb = 0; kb < a.Kb; ++kb) {
                    mbar_wait(full + stage, phase);
                    tc_fence_after();
                    const uint32_t sa0 = smem_u32(smem + sp.a_off[stage][0]), sa1 = smem_u32(smem + sp.a_off[stage][1]);
                    const uint32_t sb0 = smem_u32(smem + sp.b_off[stage][0]), sb1 = smem_u32(smem + sp.b_off[stage][1]);
#pragma unroll
                    for (int k = 0; k < kKB / 16; ++k) {            // one MMA covers K = 16: two 8-wide chunks, lbo apart
                        const uint64_t a_hi = smem_desc(sa0 + 2u * k * lbo_a, lbo_a, sbo), b_hi = smem_desc(sb0 + 2u * k * lbo_b, lbo_b, sbo);
                        const uint64_t b_lo = smem_desc(sb1 + 2u * k * lbo_b, lbo_b, sbo);
                        umma_f16(d, a_hi, b_hi, idesc, (kb | k) ? 1u : 0u);
                        umma_f16(d, a_hi, b_lo, idesc, 1u);
                        if (a.a_terms == 2) umma_f16(d, smem_desc(sa1 + 2u * k * lbo_a, lbo_a, sbo), b_hi, idesc, 1u);
                    }
                    umma_commit(empty + stage);                     // the stage is free once these MMAs have read it
                    if (++stage == kStages) { stage = 0; phase ^= 1u; }
                }
                umma_commit(acc_full + as);                         // accumulator complete -> epilogue
                if (++as == 2) { as = 0; aphase ^= 1u; }
            }
        }
    } else {
        // ===== epilogue: warps 2..5; warp w may touch tensor-memory lanes 32 * (w % 4) .. + 31 =====
        const int q = warp & 3, row = q * 32 + lane;
        uint32_t as = 0, aphase = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            const int mt = t / a.n_tiles, nt = t - mt * a.n_tiles;
            mbar_wait(acc_full + as, aphase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + as * acc_cols + ((uint32_t)(q * 32) << 16);
            const long long grow = (long long)mt * kM + row;
            if (a.Y) {
                for (int c0 = 0; c0 < a.NT; c0 += 16) {
                    uint32_t v[16];
                    tmem_ld16(taddr + (uint32_t)c0, v);
                    const int n0 = nt * a.NT + c0;
                    uint32_t hi[8], lo[8];
#pragma unroll
                    for (int i = 0; i < 16; i += 2) {
                        float x0 = __uint_as_float(v[i]) + __ldg(a.bias + n0 + i), x1 = __uint_as_float(v[i + 1]) + __ldg(a.bias + n0 + i + 1);
                        if (a.relu) { x0 = fmaxf(x0, 0.0f); x1 = fmaxf(x1, 0.0f); }
                        const __half h0 = __float2half_rn(x0), h1 = __float2half_rn(x1);
                        const __half l0 = __float2half_rn(x0 - __half2float(h0)), l1 = __float2half_rn(x1 - __half2float(h1));
                        hi[i >> 1] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
                        lo[i >> 1] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
                    }
                    // columns n0 .. n0+15 = K block n0 / 64 of the next layer, chunks (n0 % 64) / 8 and the one after
                    const int okb = n0 >> 6, ch = (n0 & 63) >> 3;
                    __half* img = a.Y + (((size_t)mt * a.out_Kb + okb) * 2) * kImgA;
                    uint4* p_hi = reinterpret_cast<uint4*>(img + ((size_t)ch * kM + row) * 8);
                    uint4* p_lo = reinterpret_cast<uint4*>(img + kImgA + ((size_t)ch * kM + row) * 8);
                    p_hi[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]); p_hi[kM] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
                    p_lo[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]); p_lo[kM] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
                }
            } else {
                // last layer: columns 0 .. A-1 are the action logits, column A the value head; softmax over the logits (three passes over the
                // accumulator: max, sum, write — tensor-memory reads are cheap, 80 live registers are not)
                const int A = a.num_actions;
                float mx = -INFINITY;
                for (int c0 = 0; c0 < a.NT; c0 += 16) {
                    uint32_t v[16];
                    tmem_ld16(taddr + (uint32_t)c0, v);
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int n = c0 + i;
                        const float x = __uint_as_float(v[i]) + __ldg(a.bias + n);
                        if (n < A) { mx = fmaxf(mx, x); if (a.logits && grow < a.batch) a.logits[(size_t)grow * A + n] = x; }
                        else if (n == A && a.has_value && a.values && grow < a.batch) a.values[grow] = x;
                    }
                }
                if (a.probs) {
                    float sum = 0.0f;
                    for (int c0 = 0; c0 < a.NT; c0 += 16) {
                        uint32_t v[16];
                        tmem_ld16(taddr + (uint32_t)c0, v);
#pragma unroll
                        for (int i = 0; i < 16; ++i) { const int n = c0 + i; if (n < A) sum += expf(__uint_as_float(v[i]) + __ldg(a.bias + n) - mx); }
                    }
                    const float inv = 1.0f / sum;
                    for (int c0 = 0; c0 < a.NT; c0 += 16) {
                        uint32_t v[16];
                        tmem_ld16(taddr + (uint32_t)c0, v);
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const int n = c0 + i;
                            if (n < A && grow < a.batch) a.probs[(size_t)grow * A + n] = expf(__uint_as_float(v[i]) + __ldg(a.bias + n) - mx) * inv;
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty + as);             // 4 epilogue warps -> buffer free for the MMA warp
            if (++as == 2) { as = 0; aphase ^= 1u; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(2 * acc_cols) : "memory");
}

// packed observation bits -> the first layer's X tile images (one half per entry: 0 / 1 are exact): thread = (row, 8-column chunk)
__global__ void k_tc_expand_bits(const uint32_t* __restrict__ bits, int obs_words, int obs_size, long long batch, int Kb, __half* __restrict__ X, long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int row = (int)(i % kM);
    long long r = i / kM;
    const int ch = (int)(r % 8); r /= 8;
    const int kb = (int)(r % Kb);
    const long long mt = r / Kb;
    const long long grow = mt * kM + row;
    const int col0 = kb * kKB + ch * 8;
    uint32_t b = 0;
    if (grow < batch && col0 < obs_size) {
        b = (bits[(size_t)grow * obs_words + (col0 >> 5)] >> (col0 & 31)) & 0xFFu;
        if (col0 + 8 > obs_size) b &= (1u << (obs_size - col0)) - 1u;
    }
    const uint32_t one = 0x3C00u;     // half(1.0)
    uint4 o;
    o.x = ((b & 1u) ? one : 0u) | ((b & 2u) ? one << 16 : 0u);
    o.y = ((b & 4u) ? one : 0u) | ((b & 8u) ? one << 16 : 0u);
    o.z = ((b & 16u) ? one : 0u) | ((b & 32u) ? one << 16 : 0u);
    o.w = ((b & 64u) ? one : 0u) | ((b & 128u) ? one << 16 : 0u);
    reinterpret_cast<uint4*>(X)[i] = o;       // image order [mt][kb][ch][row] x 16 bytes == i
}

}  // namespace tc
}  // namespace qg

using namespace qg;

struct TcLayer {
    int K = 0, N = 0, Kb = 0, NT = 0, n_tiles = 0, Npad = 0, a_terms = 2, relu = 1;
    __half* W = nullptr; float* bias = nullptr;
};
struct qg_policy_tc {
    int device = 0, obs_size = 0, obs_words = 0, num_actions = 0, has_value = 0, num_sms = 148;
    long long max_batch = 0; int m_tiles = 0;
    std::vector<TcLayer> layers;
    std::vector<__half*> acts;     // acts[l] = X images of layer l
};

namespace {
#define TC_CUDA_OK(expr)                                                                       \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                     \
            return QG_ERR_CUDA;                                                                \
        }                                                                                      \
    } while (0)
inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
}  // namespace

extern "C" {

void qg_policy_tc_destroy(qg_policy_tc* p) {
    if (!p) return;
    for (auto& l : p->layers) { if (l.W) cudaFree(l.W); if (l.bias) cudaFree(l.bias); }
    for (auto* a : p->acts) if (a) cudaFree(a);
    delete p;
}

int qg_policy_tc_create(int32_t device, int32_t obs_size, int32_t num_layers, const int32_t* out_features, const float* const* weights_host,
                        const float* const* biases_host, const float* value_weight_host, float value_bias, int64_t max_batch, qg_policy_tc** out) {
    if (!out || !out_features || !weights_host || !biases_host) { set_error("null argument"); return QG_ERR_INVALID; }
    *out = nullptr;
    if (obs_size < 1 || num_layers < 1 || num_layers > 8 || max_batch < 1) { set_error("qg_policy_tc_create: bad sizes"); return QG_ERR_INVALID; }
    int ndev = 0;
    TC_CUDA_OK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) { set_error("no such CUDA device (the engine has no CPU fallback)"); return QG_ERR_CUDA; }
    TC_CUDA_OK(cudaSetDevice(device));
    int major = 0;
    TC_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    if (major != 10) { set_error("the tensor-core policy needs an sm_100 device (tcgen05)"); return QG_ERR_UNSUPPORTED; }
    qg_policy_tc* p = new (std::nothrow) qg_policy_tc();
    if (!p) { set_error("out of memory"); return QG_ERR_INVALID; }
    p->device = device; p->obs_size = obs_size; p->obs_words = (obs_size + 31) / 32; p->max_batch = max_batch;
    p->m_tiles = (int)((max_batch + tc::kM - 1) / tc::kM);
    p->has_value = value_weight_host ? 1 : 0;
    p->num_actions = out_features[num_layers - 1];
    { int v = 0; if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && v > 0) p->num_sms = v; }
    auto fail = [&](int rc) { qg_policy_tc_destroy(p); return rc; };
    int K = obs_size;
    for (int l = 0; l < num_layers; ++l) {
        TcLayer L;
        const bool last = l == num_layers - 1;
        L.K = K; L.N = out_features[l] + (last ? p->has_value : 0);
        if (out_features[l] < 1 || out_features[l] > 4096) { set_error("qg_policy_tc_create: layer width outside 1..4096"); return fail(QG_ERR_UNSUPPORTED); }
        L.Kb = round_up(K, tc::kKB) / tc::kKB;
        L.a_terms = l == 0 ? 1 : 2; L.relu = last ? 0 : 1;
        if (last) {
            L.Npad = round_up(L.N, 16);
            if (L.Npad > tc::kMaxNT) { set_error("qg_policy_tc_create: more than 127 actions are not supported by the tensor-core head"); return fail(QG_ERR_UNSUPPORTED); }
            L.NT = L.Npad; L.n_tiles = 1;
        } else {
            L.Npad = round_up(L.N, tc::kKB);                      // = the next layer's padded K
            L.NT = (L.Npad % 128 == 0) ? 128 : 64; L.n_tiles = L.Npad / L.NT;
        }
        // weight tile images [term][n_tile][kb][k/8][n][k%8], halves; zero padding rows / columns
        const size_t img = (size_t)L.NT * tc::kKB, count = (size_t)2 * L.n_tiles * L.Kb * img;
        std::vector<__half> wimg(count, __float2half(0.0f));
        std::vector<float> bias((size_t)L.Npad, 0.0f);
        const float* W = weights_host[l];
        for (int n = 0; n < L.N; ++n) {
            const bool vrow = last && p->has_value && n == out_features[l];
            bias[n] = vrow ? value_bias : biases_host[l][n];
            const float* wr = vrow ? value_weight_host : W + (size_t)n * K;
            const int nt = n / L.NT, nl = n % L.NT;
            for (int k = 0; k < K; ++k) {
                const float w = wr[k];
                const __half hi = __float2half_rn(w), lo = __float2half_rn(w - __half2float(hi));
                const int kb = k / tc::kKB, kl = k % tc::kKB;
                const size_t at = (((size_t)nt * L.Kb + kb) * img) + ((size_t)(kl >> 3) * L.NT + nl) * 8 + (kl & 7);
                wimg[at] = hi;
                wimg[(size_t)L.n_tiles * L.Kb * img + at] = lo;
            }
        }
        cudaError_t ce = cudaMalloc(&L.W, count * sizeof(__half));
        if (ce == cudaSuccess) ce = cudaMemcpy(L.W, wimg.data(), count * sizeof(__half), cudaMemcpyHostToDevice);
        if (ce == cudaSuccess) ce = cudaMalloc(&L.bias, bias.size() * 4);
        if (ce == cudaSuccess) ce = cudaMemcpy(L.bias, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice);
        p->layers.push_back(L);
        if (ce != cudaSuccess) { set_error(std::string("qg_policy_tc_create: ") + cudaGetErrorString(ce)); return fail(QG_ERR_CUDA); }
        // this layer's input images
        __half* X = nullptr;
        const size_t xbytes = (size_t)p->m_tiles * L.Kb * L.a_terms * tc::kImgA * sizeof(__half);
        ce = cudaMalloc(&X, xbytes);
        if (ce == cudaSuccess) ce = cudaMemset(X, 0, xbytes);
        p->acts.push_back(X);
        if (ce != cudaSuccess) { set_error(std::string("qg_policy_tc_create: ") + cudaGetErrorString(ce)); return fail(QG_ERR_CUDA); }
        K = out_features[l];
    }
    // (layer l writes round_up(N_l, 64) columns = all K blocks of layer l+1: padded columns come out as relu(0 + 0) = 0)
    size_t max_smem = 0;
    for (auto& L : p->layers) max_smem = std::max<size_t>(max_smem, tc::plan_smem(L.a_terms, L.NT).total);
    cudaError_t ce = cudaFuncSetAttribute(tc::k_tc_layer, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem);
    if (ce != cudaSuccess) { set_error(std::string("qg_policy_tc_create: ") + cudaGetErrorString(ce)); return fail(QG_ERR_CUDA); }
    *out = p;
    return QG_OK;
}

int32_t qg_policy_tc_num_actions(const qg_policy_tc* p) { return p ? p->num_actions : 0; }

int qg_policy_tc_forward_bits(qg_policy_tc* p, const uint32_t* obs_bits_dev, int64_t batch, float* probs_dev, float* logits_dev, float* values_dev,
                              qg_stream stream) {
    if (!p || !obs_bits_dev) { set_error("null argument"); return QG_ERR_INVALID; }
    if (batch < 0 || batch > p->max_batch) { set_error("qg_policy_tc_forward_bits: batch larger than the handle was created for"); return QG_ERR_INVALID; }
    if (values_dev && !p->has_value) { set_error("qg_policy_tc_forward_bits: the policy has no value head"); return QG_ERR_INVALID; }
    if (batch == 0) return QG_OK;
    TC_CUDA_OK(cudaSetDevice(p->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int m_tiles = (int)((batch + tc::kM - 1) / tc::kM);
    {
        const TcLayer& L0 = p->layers[0];
        const long long total = (long long)m_tiles * L0.Kb * 8 * tc::kM;
        tc::k_tc_expand_bits<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(obs_bits_dev, p->obs_words, p->obs_size, batch, L0.Kb, p->acts[0], total);
        TC_CUDA_OK(cudaGetLastError());
    }
    for (size_t l = 0; l < p->layers.size(); ++l) {
        const TcLayer& L = p->layers[l];
        const bool last = l + 1 == p->layers.size();
        tc::LayerArgs a{};
        a.X = p->acts[l]; a.W = L.W; a.bias = L.bias; a.Y = last ? nullptr : p->acts[l + 1];
        a.logits = last ? logits_dev : nullptr; a.probs = last ? probs_dev : nullptr; a.values = last ? values_dev : nullptr;
        a.a_terms = L.a_terms; a.Kb = L.Kb; a.NT = L.NT; a.n_tiles = L.n_tiles; a.m_tiles = m_tiles;
        a.out_Kb = last ? 0 : p->layers[l + 1].Kb; a.relu = L.relu; a.num_actions = p->num_actions; a.has_value = p->has_value; a.batch = batch;
        const int tiles = m_tiles * L.n_tiles;
        const size_t smem = tc::plan_smem(L.a_terms, L.NT).total;
        tc::k_tc_layer<<<(unsigned)std::min(tiles, p->num_sms), tc::kThreads, smem, st>>>(a);
        TC_CUDA_OK(cudaGetLastError());
    }
    return QG_OK;
}

}  // extern "C"
