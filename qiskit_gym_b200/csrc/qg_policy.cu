// qg_policy.cu — the policy's action network evaluated straight from packed-bit observations.
//
// twisterl feeds Env::observe()'s sparse indices to an embedding-bag first layer (the reason observe() returns indices,
// e.g. clifford.rs:361-368); the dense f32 observation the engine materialises for a PyTorch policy is 32x larger than
// the information it carries.  Here the observation stays packed (one bit per entry, written by the step kernel) and the
// first layer is a gather-sum of weight columns over the set bits; the remaining layers of the BasicPolicy-shaped MLP
// (Linear -> ReLU chain, examples/models/*.pt) and the softmax run in the same kernel, so one synth-search decision is
// two launches (this one + the fused sample/step kernel) instead of a dozen library kernels.
//
// Decomposition: one CTA of 256 threads owns 8 batch rows.  Activations live in shared memory as [feature][row]
// (32 bytes per feature: every thread reads the same address, a broadcast); thread j owns output features j, j+256, ..
// and keeps one f32 accumulator per (feature, row) in registers, so each weight feeds 8 FMAs.  Weights are kept
// transposed ([in][out], rows padded to 16 bytes) and stream L2 -> shared memory in 16 KB tiles with cp.async.bulk
// completing on mbarriers, three tiles in flight, one flat tile schedule across all layers so the pipeline never
// drains at a layer boundary.  The first layer's tiles are the weight rows of the observation entries set in any of
// the CTA's rows (its input is then the rows' 0/1 values, so it runs through the same FMA loop as the dense layers).
// Sums run over the input features in ascending order, in f32.
#include <algorithm>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "qg_host.hpp"

namespace qg {

constexpr int kPolRows = 8;          // batch rows per CTA
constexpr int kPolConsumers = 256;   // 8 compute warps
constexpr int kPolThreads = kPolConsumers + 32;   // + 1 producer warp
constexpr int kPolOutPerThread = 4;  // layer widths up to 1024
constexpr int kPolMaxLayers = 8;
constexpr int kPolStages = 3;        // weight tiles in flight
constexpr int kPolTileFloats = 4096; // 16 KB per weight tile

struct PolicyDev {
    int32_t num_layers, obs_size, obs_words;
    int32_t width[kPolMaxLayers];        // output features of layer l (layer 0 consumes the observation)
    int32_t stride[kPolMaxLayers];       // width rounded up to a multiple of 4 floats: row stride of the transposed weights
    const float* wt[kPolMaxLayers];      // transposed weights [in][stride]
    const float* bias[kPolMaxLayers];
    int32_t act0_floats, act1_floats;    // activation buffers ([feature][row]); act1 also holds the first layer's 0/1 inputs
};

// ---- mbarrier / bulk-copy primitives (weights stream L2 -> shared memory with cp.async.bulk, no register staging) ------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n .reg .pred p;\n"
        "WAIT_LOOP:\n"
        " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        " @p bra DONE;\n"
        " bra WAIT_LOOP;\n"
        "DONE:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* sdst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(sdst)), "l"(gsrc), "r"(bytes),
                 "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void consumers_sync() { asm volatile("bar.sync 1, %0;\n" ::"n"(kPolConsumers) : "memory"); }

// Rows of layer l's transposed weights consumed per tile
__device__ __forceinline__ int tile_rows(const PolicyDev& p, int l) { return max(1, kPolTileFloats / p.stride[l]); }
__device__ __forceinline__ int layer_tiles(const PolicyDev& p, int l, int U) {
    const int K = l == 0 ? U : p.width[l - 1], kt = tile_rows(p, l);
    return (K + kt - 1) / kt;
}

// One layer for the consumer warps: NI output features per thread (j = tid + i * 256), 8 rows each.
template <int NI>
__device__ __forceinline__ void consume_layer(const PolicyDev& p, int l, int K, const float* __restrict__ src, float* __restrict__ dst, const float* tiles,
                                              uint64_t* full, uint64_t* empty, int& G, int tid, int lane) {
    const int out = p.width[l], ostr = p.stride[l], kt = tile_rows(p, l), nt = (K + kt - 1) / kt;
    float acc[NI][kPolRows];
    int jc[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) {
        const int j = tid + i * kPolConsumers;
        jc[i] = min(j, ostr - 1);                          // out-of-range features read a valid word and are never written
        const float b = j < out ? __ldg(p.bias[l] + j) : 0.0f;
#pragma unroll
        for (int r = 0; r < kPolRows; ++r) acc[i][r] = b;
    }
    const bool active = NI > 1 || (tid & ~31) < out;       // warp-uniform: this warp owns at least one real feature
    for (int t = 0; t < nt; ++t, ++G) {
        const int stage = G % kPolStages;
        mbar_wait(full + stage, (uint32_t)((G / kPolStages) & 1));
        if (active) {
            const float* tile = tiles + (size_t)stage * kPolTileFloats;
            const int k0 = t * kt, rows = min(kt, K - k0);
            const float4* h = reinterpret_cast<const float4*>(src + (size_t)k0 * kPolRows);
#pragma unroll 4
            for (int u = 0; u < rows; ++u) {
                const float4 h0 = h[2 * u], h1 = h[2 * u + 1];
#pragma unroll
                for (int i = 0; i < NI; ++i) {
                    const float w = tile[u * ostr + jc[i]];
                    acc[i][0] = fmaf(w, h0.x, acc[i][0]); acc[i][1] = fmaf(w, h0.y, acc[i][1]);
                    acc[i][2] = fmaf(w, h0.z, acc[i][2]); acc[i][3] = fmaf(w, h0.w, acc[i][3]);
                    acc[i][4] = fmaf(w, h1.x, acc[i][4]); acc[i][5] = fmaf(w, h1.y, acc[i][5]);
                    acc[i][6] = fmaf(w, h1.z, acc[i][6]); acc[i][7] = fmaf(w, h1.w, acc[i][7]);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + stage);         // this warp is done with the stage
    }
    const bool last = l == p.num_layers - 1;
#pragma unroll
    for (int i = 0; i < NI; ++i) {
        const int j = tid + i * kPolConsumers;
        if (j < out) {
            float4 lo, hi;
            lo.x = acc[i][0]; lo.y = acc[i][1]; lo.z = acc[i][2]; lo.w = acc[i][3];
            hi.x = acc[i][4]; hi.y = acc[i][5]; hi.z = acc[i][6]; hi.w = acc[i][7];
            if (!last) {
                lo.x = fmaxf(lo.x, 0.f); lo.y = fmaxf(lo.y, 0.f); lo.z = fmaxf(lo.z, 0.f); lo.w = fmaxf(lo.w, 0.f);
                hi.x = fmaxf(hi.x, 0.f); hi.y = fmaxf(hi.y, 0.f); hi.z = fmaxf(hi.z, 0.f); hi.w = fmaxf(hi.w, 0.f);
            }
            reinterpret_cast<float4*>(dst + (size_t)j * kPolRows)[0] = lo;
            reinterpret_cast<float4*>(dst + (size_t)j * kPolRows)[1] = hi;
        }
    }
}

// Warps 0..7 compute (consumers); warp 8 is the producer: one lane streams the weight tiles of all layers, in order, through
// a ring of kPolStages shared-memory stages (full / empty mbarriers), so no consumer ever waits for copy issue.
__global__ void __launch_bounds__(kPolThreads) k_policy_mlp(const __grid_constant__ PolicyDev p, const uint32_t* __restrict__ bits, int64_t B,
                                                             float* __restrict__ probs, float* __restrict__ logits_out) {
    extern __shared__ __align__(128) float sm[];
    float* tiles = sm;                                                     // [kPolStages][kPolTileFloats]
    float* act0 = tiles + kPolStages * kPolTileFloats;
    float* act1 = act0 + p.act0_floats;
    uint64_t* full = reinterpret_cast<uint64_t*>(act1 + p.act1_floats);    // [kPolStages]
    uint64_t* empty = full + kPolStages;                                   // [kPolStages]
    uint32_t* rowbits = reinterpret_cast<uint32_t*>(empty + kPolStages);   // [kPolRows][obs_words]
    int* wcnt = reinterpret_cast<int*>(rowbits + kPolRows * p.obs_words);  // [8] per-warp counts, [15] = U
    uint16_t* uidx = reinterpret_cast<uint16_t*>(wcnt + 16);               // [obs_size] observation entries set in any of the 8 rows
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool producer = tid >= kPolConsumers;
    const int64_t row0 = (int64_t)blockIdx.x * kPolRows;

    if (tid == 0) {
        for (int s = 0; s < kPolStages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, kPolConsumers / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    // programmatic dependent launch: everything above overlaps the tail of the previous kernel in the stream (the step kernel that
    // writes `bits`); the next kernel (the step kernel that reads `probs`) may be scheduled now, it waits for this grid itself
    asm volatile("griddepcontrol.wait;\n" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
    // ---- 1. the entries set in any of the CTA's rows, ascending, with the rows' 0/1 values as the first layer's input ----
    if (!producer) {
        for (int i = tid; i < kPolRows * p.obs_words; i += kPolConsumers) {
            const int r = i / p.obs_words, w = i - r * p.obs_words;
            uint32_t word = (row0 + r < B) ? bits[(size_t)(row0 + r) * p.obs_words + w] : 0u;
            if (w == p.obs_words - 1 && (p.obs_size & 31)) word &= (1u << (p.obs_size & 31)) - 1u;
            rowbits[i] = word;
        }
        consumers_sync();
        int U = 0;
        for (int base = 0; base < p.obs_size; base += kPolConsumers) {
            const int k = base + tid;
            uint32_t m = 0;
            if (k < p.obs_size) {
#pragma unroll
                for (int r = 0; r < kPolRows; ++r) m |= ((rowbits[r * p.obs_words + (k >> 5)] >> (k & 31)) & 1u) << r;
            }
            const uint32_t vote = __ballot_sync(0xFFFFFFFFu, m != 0);
            if (lane == 0) wcnt[warp] = __popc(vote);
            consumers_sync();
            int before = U, total = U;
#pragma unroll
            for (int w2 = 0; w2 < kPolConsumers / 32; ++w2) { const int c = wcnt[w2]; if (w2 < warp) before += c; total += c; }
            if (m) {
                const int at = before + __popc(vote & ((1u << lane) - 1u));
                uidx[at] = (uint16_t)k;
                float4 lo, hi;
                lo.x = (m & 1u) ? 1.f : 0.f; lo.y = (m & 2u) ? 1.f : 0.f; lo.z = (m & 4u) ? 1.f : 0.f; lo.w = (m & 8u) ? 1.f : 0.f;
                hi.x = (m & 16u) ? 1.f : 0.f; hi.y = (m & 32u) ? 1.f : 0.f; hi.z = (m & 64u) ? 1.f : 0.f; hi.w = (m & 128u) ? 1.f : 0.f;
                reinterpret_cast<float4*>(act1 + (size_t)at * kPolRows)[0] = lo;
                reinterpret_cast<float4*>(act1 + (size_t)at * kPolRows)[1] = hi;
            }
            U = total;
            consumers_sync();
        }
        if (tid == 0) wcnt[15] = U;
    }
    __syncthreads();                               // barriers initialised, U and uidx visible to the producer
    const int U = wcnt[15];

    // ---- 2. producer: tile G of the flat schedule -> (layer, first input row).  Layer 0 consumes the U listed observation
    // entries (gathered rows of the transposed first-layer weights), layer l > 0 its width[l-1] inputs (contiguous rows).
    if (producer) {
        if (lane == 0) {
            int G = 0;
            for (int l = 0; l < p.num_layers; ++l) {
                const int K = l == 0 ? U : p.width[l - 1], kt = tile_rows(p, l), nt = (K + kt - 1) / kt;
                const uint32_t row_bytes = (uint32_t)p.stride[l] * 4u;
                for (int t = 0; t < nt; ++t, ++G) {
                    const int stage = G % kPolStages, k0 = t * kt, rows = min(kt, K - k0);
                    if (G >= kPolStages) mbar_wait(empty + stage, (uint32_t)(((G / kPolStages) - 1) & 1));
                    float* dst = tiles + (size_t)stage * kPolTileFloats;
                    mbar_expect_tx(full + stage, row_bytes * (uint32_t)rows);
                    if (l == 0) {
                        for (int u = 0; u < rows; ++u) bulk_g2s(dst + (size_t)u * p.stride[0], p.wt[0] + (size_t)uidx[k0 + u] * p.stride[0], row_bytes, full + stage);
                    } else {
                        bulk_g2s(dst, p.wt[l] + (size_t)k0 * p.stride[l], row_bytes * (uint32_t)rows, full + stage);
                    }
                }
            }
        }
        return;
    }

    // ---- 3. consumers ------------------------------------------------------------------------------------------------------
    float* src = act1;                             // layer 0 reads the 0/1 inputs
    float* dst = act0;
    int G = 0;
    for (int l = 0; l < p.num_layers; ++l) {
        const int K = l == 0 ? U : p.width[l - 1];
        switch ((p.width[l] + kPolConsumers - 1) / kPolConsumers) {
            case 1: consume_layer<1>(p, l, K, src, dst, tiles, full, empty, G, tid, lane); break;
            case 2: consume_layer<2>(p, l, K, src, dst, tiles, full, empty, G, tid, lane); break;
            case 3: consume_layer<3>(p, l, K, src, dst, tiles, full, empty, G, tid, lane); break;
            default: consume_layer<4>(p, l, K, src, dst, tiles, full, empty, G, tid, lane); break;
        }
        consumers_sync();
        src = dst;
        dst = (dst == act0) ? act1 : act0;
    }

    // ---- 4. softmax over the action logits, warp r <-> row r -------------------------------------------------------------
    if (warp < kPolRows) {
        const int64_t row = row0 + warp;
        if (row < B) {
            const int A = p.width[p.num_layers - 1];
            float mx = -INFINITY;
            for (int a = lane; a < A; a += 32) mx = fmaxf(mx, src[(size_t)a * kPolRows + warp]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
            float sum = 0.0f;
            for (int a = lane; a < A; a += 32) sum += expf(src[(size_t)a * kPolRows + warp] - mx);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
            const float inv = 1.0f / sum;
            for (int a = lane; a < A; a += 32) {
                const float lg = src[(size_t)a * kPolRows + warp];
                if (probs) probs[(size_t)row * A + a] = expf(lg - mx) * inv;
                if (logits_out) logits_out[(size_t)row * A + a] = lg;
            }
        }
    }
}

}  // namespace qg

using namespace qg;

struct qg_policy {
    int device = 0;
    PolicyDev d{};
    std::vector<float*> bufs;
    size_t smem = 0;
};

#define POL_CUDA_OK(expr)                                                                      \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                     \
            return QG_ERR_CUDA;                                                                \
        }                                                                                      \
    } while (0)

extern "C" {

void qg_policy_destroy(qg_policy* p) {
    if (!p) return;
    for (float* b : p->bufs) cudaFree(b);
    delete p;
}

int qg_policy_create(int32_t device, int32_t obs_size, int32_t num_layers, const int32_t* out_features, const float* const* weights_host,
                     const float* const* biases_host, qg_policy** out) {
    if (!out) { set_error("null out"); return QG_ERR_INVALID; }
    *out = nullptr;
    if (obs_size < 1 || obs_size > 65535) { set_error("qg_policy_create: obs_size must be in [1, 65535]"); return QG_ERR_UNSUPPORTED; }
    if (num_layers < 1 || num_layers > kPolMaxLayers || !out_features || !weights_host || !biases_host) { set_error("qg_policy_create: 1..8 layers with weights and biases"); return QG_ERR_INVALID; }
    int maxw = 0;
    for (int l = 0; l < num_layers; ++l) {
        if (out_features[l] < 1 || out_features[l] > kPolConsumers * kPolOutPerThread) { set_error("qg_policy_create: layer widths must be in [1, 1024]"); return QG_ERR_UNSUPPORTED; }
        if (!weights_host[l] || !biases_host[l]) { set_error("qg_policy_create: null layer"); return QG_ERR_INVALID; }
        maxw = std::max(maxw, out_features[l]);
    }
    int ndev = 0;
    POL_CUDA_OK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) { set_error("no such CUDA device (the engine has no CPU fallback)"); return QG_ERR_CUDA; }
    POL_CUDA_OK(cudaSetDevice(device));
    qg_policy* p = new (std::nothrow) qg_policy();
    if (!p) { set_error("out of memory"); return QG_ERR_INVALID; }
    p->device = device;
    PolicyDev& d = p->d;
    d.num_layers = num_layers; d.obs_size = obs_size; d.obs_words = (obs_size + 31) / 32;
    d.act0_floats = maxw * kPolRows;
    d.act1_floats = std::max(maxw, obs_size) * kPolRows;
    for (int l = 0; l < num_layers; ++l) {
        const int in = l == 0 ? obs_size : out_features[l - 1], o = out_features[l], os = (o + 3) / 4 * 4;
        d.width[l] = o; d.stride[l] = os;
        std::vector<float> t((size_t)in * os, 0.0f);
        for (int j = 0; j < o; ++j) for (int k = 0; k < in; ++k) t[(size_t)k * os + j] = weights_host[l][(size_t)j * in + k];   // torch Linear.weight is [out][in]
        float *w = nullptr, *b = nullptr;
        cudaError_t ce = cudaMalloc(&w, t.size() * 4);
        if (ce == cudaSuccess) { p->bufs.push_back(w); ce = cudaMalloc(&b, (size_t)o * 4); }
        if (ce == cudaSuccess) { p->bufs.push_back(b); ce = cudaMemcpy(w, t.data(), t.size() * 4, cudaMemcpyHostToDevice); }
        if (ce == cudaSuccess) ce = cudaMemcpy(b, biases_host[l], (size_t)o * 4, cudaMemcpyHostToDevice);
        if (ce != cudaSuccess) { set_error(std::string("qg_policy_create: ") + cudaGetErrorString(ce)); qg_policy_destroy(p); return QG_ERR_CUDA; }
        d.wt[l] = w; d.bias[l] = b;
    }
    p->smem = ((size_t)kPolStages * kPolTileFloats + d.act0_floats + d.act1_floats) * 4 + 2 * kPolStages * 8 + (size_t)kPolRows * d.obs_words * 4 + 16 * 4 +
              (size_t)obs_size * 2 + 128;
    if (p->smem > 200 * 1024) { set_error("qg_policy_create: the network needs more shared memory than one SM has"); qg_policy_destroy(p); return QG_ERR_UNSUPPORTED; }
    {
        cudaError_t ce = cudaFuncSetAttribute(k_policy_mlp, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
        if (ce == cudaSuccess && p->smem > 48 * 1024) ce = cudaFuncSetAttribute(k_policy_mlp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem);
        if (ce != cudaSuccess) { set_error(std::string("qg_policy_create: ") + cudaGetErrorString(ce)); qg_policy_destroy(p); return QG_ERR_CUDA; }
    }
    *out = p;
    return QG_OK;
}

int32_t qg_policy_num_actions(const qg_policy* p) { return p ? p->d.width[p->d.num_layers - 1] : 0; }

int qg_policy_forward_bits(qg_policy* p, const uint32_t* obs_bits_dev, int64_t batch, float* probs_dev, float* logits_dev, qg_stream stream) {
    if (!p || !obs_bits_dev || (!probs_dev && !logits_dev)) { set_error("null argument"); return QG_ERR_INVALID; }
    if (batch < 0) { set_error("qg_policy_forward_bits: negative batch"); return QG_ERR_INVALID; }
    if (batch == 0) return QG_OK;
    int cur = -1;
    POL_CUDA_OK(cudaGetDevice(&cur));
    if (cur != p->device) POL_CUDA_OK(cudaSetDevice(p->device));
    const unsigned grid = (unsigned)((batch + kPolRows - 1) / kPolRows);
    cudaLaunchConfig_t lc{};
    lc.gridDim = dim3(grid); lc.blockDim = dim3(kPolThreads); lc.dynamicSmemBytes = p->smem; lc.stream = (cudaStream_t)stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = at; lc.numAttrs = 1;
    POL_CUDA_OK(cudaLaunchKernelEx(&lc, k_policy_mlp, p->d, obs_bits_dev, batch, probs_dev, logits_dev));
    return QG_OK;
}

}  // extern "C"
