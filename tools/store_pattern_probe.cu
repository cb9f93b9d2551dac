// store_pattern_probe.cu — write-bandwidth ceilings of the step kernel's OUTPUT PATTERN, without the env logic.
//
// The fused step kernel's warps each own 32 environments and write, per env-step, one contiguous slab of
// 32 * obs_size floats (C3: 32 KB) into that step's [B][obs] tensor; 2048 warps do this for 128 steps with
// a latency-bound phase between slabs.  This probe replays only that store pattern, to separate "the pattern
// cannot go faster" from "the kernel leaves bandwidth on the table":
//   fill     grid-stride st.global.cs.v4 over the whole buffer (the fill ceiling)
//   stg      one warp per 32-env tile, per step 64 x st.global.cs.v4 (512 B per warp instruction), optional
//            busy delay between steps (emulates phase 1)
//   tma      same tiles, but the warp fills a shared-memory chunk (st.shared.v4) and one lane issues
//            cp.async.bulk.global.shared::cta (chunks double buffered); the warp does not wait for the drain
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/store_pattern_probe tools/store_pattern_probe.cu
// Run  :  tools/store_pattern_probe [envs=65536] [obs=256] [steps=128] [delay_ns=0] [chunk_bytes=4096] [warps_per_cta=2]
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void busy_ns(int ns) {
    if (ns <= 0) return;
    long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); } while (t1 - t0 < ns);
}

__global__ void k_fill(float4* out, size_t n4) {
    const float4 v = make_float4(1.f, 0.f, 1.f, 0.f);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) __stcs(out + i, v);
}

// ring: [R][B][obs] floats
// skew_ns > 0: warp w starts hash(w) % skew_ns late, so the warps are spread over the step period (and over the ring slots) like
// the drifting warps of the real kernel instead of walking the ring in lock-step
__global__ void k_stg(float* ring, int R, long B, int obs, int steps, int delay_ns, int skew_ns = 0) {
    const int lane = threadIdx.x & 31;
    const long tile = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (tile * 32 >= B) return;
    const int n4 = obs * 8;   // float4 per tile
    if (skew_ns > 0) busy_ns((int)(((unsigned)tile * 2654435761u >> 8) % (unsigned)skew_ns));
    for (int t = 0; t < steps; ++t) {
        busy_ns(delay_ns);
        float4* out = reinterpret_cast<float4*>(ring + ((size_t)(t % R) * B + tile * 32) * obs);
        const float4 v = make_float4((float)(t & 1), 0.f, 1.f, (float)(lane & 1));
#pragma unroll 4
        for (int j = lane; j < n4; j += 32) __stcs(out + j, v);
    }
}

// K warps share one 32-env tile: warp w stores the 512-byte rows w, w+K, ... of the tile's slab (CTA = one tile).  With
// delay_ns > 0 every warp idles before each slab (a CTA that alternates between phase 1 and the expansion in lock-step); with 0
// the warps store back to back (a producer warp runs phase 1 ahead of the storing warps).
__global__ void k_stgk(float* ring, int R, long B, int obs, int steps, int delay_ns) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, K = blockDim.x >> 5;
    const long tile = blockIdx.x;
    if (tile * 32 >= B) return;
    const int n4 = obs * 8;   // float4 per tile
    for (int t = 0; t < steps; ++t) {
        busy_ns(delay_ns);
        float4* out = reinterpret_cast<float4*>(ring + ((size_t)(t % R) * B + tile * 32) * obs);
        const float4 v = make_float4((float)(t & 1), 0.f, 1.f, (float)(lane & 1));
#pragma unroll 4
        for (int j = warp * 32 + lane; j < n4; j += 32 * K) __stcs(out + j, v);
    }
}

// strided tiles: warp w owns the environments w, w + W, w + 2W, .. (W = number of warps), so the rows the W warps write at the same time
// are neighbours in memory (one contiguous W * obs * 4-byte region per k) instead of 32 KB apart
__global__ void k_stg_strided(float* ring, int R, long B, int obs, int steps, int delay_ns) {
    const int lane = threadIdx.x & 31;
    const long W = B / 32, w = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= W) return;
    const int per_env4 = obs / 4;             // float4 per env row
    for (int t = 0; t < steps; ++t) {
        busy_ns(delay_ns);
        const float4 v = make_float4((float)(t & 1), 0.f, 1.f, (float)(lane & 1));
        for (int k = 0; k < 32; ++k) {
            float4* out = reinterpret_cast<float4*>(ring + ((size_t)(t % R) * B + (w + W * k)) * obs);
            for (int j = lane; j < per_env4; j += 32) __stcs(out + j, v);
        }
    }
}

__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(ssrc);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gdst), "r"(s), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;\n" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

__global__ void k_tma(float* ring, int R, long B, int obs, int steps, int delay_ns, int chunk_bytes) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long tile = (long)blockIdx.x * (blockDim.x >> 5) + warp;
    if (tile * 32 >= B) return;
    uint8_t* buf = smem + (size_t)warp * 2 * chunk_bytes;
    const int slab = obs * 128;                 // bytes per tile per step
    int which = 0;
    for (int t = 0; t < steps; ++t) {
        busy_ns(delay_ns);
        uint8_t* out = reinterpret_cast<uint8_t*>(ring + ((size_t)(t % R) * B + tile * 32) * obs);
        const float4 v = make_float4((float)(t & 1), 0.f, 1.f, (float)(lane & 1));
        for (int o = 0; o < slab; o += chunk_bytes) {
            const int len = min(chunk_bytes, slab - o);
            uint8_t* b = buf + which * chunk_bytes;
            if (lane == 0) bulk_wait_read<1>();          // the bulk copy issued two chunks ago has read its buffer
            __syncwarp();
            for (int j = lane * 16; j < len; j += 512) *reinterpret_cast<float4*>(b + j) = v;
            fence_async_smem();
            __syncwarp();
            if (lane == 0) { bulk_store(out + o, b, (uint32_t)len); bulk_commit(); }
            which ^= 1;
        }
    }
    if (lane == 0) bulk_wait_read<0>();
    __syncwarp();
}

int main(int argc, char** argv) {
    const long B = argc > 1 ? atol(argv[1]) : 65536;
    const int obs = argc > 2 ? atoi(argv[2]) : 256;
    const int steps = argc > 3 ? atoi(argv[3]) : 128;
    const int delay = argc > 4 ? atoi(argv[4]) : 0;
    const int chunk = argc > 5 ? atoi(argv[5]) : 4096;
    const int wpc = argc > 6 ? atoi(argv[6]) : 2;
    const int R = 5;
    const size_t slot = (size_t)B * obs * 4, total = slot * R;
    float* ring; CK(cudaMalloc(&ring, total));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const long tiles = (B + 31) / 32;
    const unsigned grid = (unsigned)((tiles + wpc - 1) / wpc);
    const size_t sm = (size_t)wpc * 2 * chunk;
    CK(cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    auto timeit = [&](const char* name, auto launch, double bytes) {
        for (int i = 0; i < 3; ++i) launch();
        CK(cudaDeviceSynchronize());
        float best = 1e30f;
        for (int rep = 0; rep < 5; ++rep) {
            CK(cudaEventRecord(e0));
            launch();
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (ms < best) best = ms;
        }
        CK(cudaGetLastError());
        printf("{\"probe\": \"%s\", \"envs\": %ld, \"obs\": %d, \"steps\": %d, \"delay_ns\": %d, \"chunk\": %d, \"warps_per_cta\": %d, \"ms\": %.4f, \"gbs\": %.1f}\n",
               name, B, obs, steps, delay, chunk, wpc, best, bytes / best / 1e6);
    };
    const double bytes = (double)slot * steps;
    timeit("fill", [&] { for (int t = 0; t < steps; ++t) k_fill<<<148 * 8, 256>>>(reinterpret_cast<float4*>(ring + (size_t)(t % R) * B * obs), slot / 16); }, bytes);
    timeit("stg", [&] { k_stg<<<grid, wpc * 32>>>(ring, R, B, obs, steps, delay); }, bytes);
    timeit("stg_strided", [&] { k_stg_strided<<<grid, wpc * 32>>>(ring, R, B, obs, steps, delay); }, bytes);
    for (int skew = 5000; skew <= 80000; skew *= 4) {
        char nm[24]; snprintf(nm, sizeof nm, "stg_skew%d", skew);
        timeit(nm, [&] { k_stg<<<grid, wpc * 32>>>(ring, R, B, obs, steps, delay, skew); }, bytes);
    }
    for (int K = 2; K <= 4; K *= 2) {
        char nm[16]; snprintf(nm, sizeof nm, "stgk%d", K);
        timeit(nm, [&] { k_stgk<<<(unsigned)tiles, K * 32>>>(ring, R, B, obs, steps, delay); }, bytes);
    }
    timeit("tma", [&] { k_tma<<<grid, wpc * 32, sm>>>(ring, R, B, obs, steps, delay, chunk); }, bytes);
    CK(cudaFree(ring));
    return 0;
}
