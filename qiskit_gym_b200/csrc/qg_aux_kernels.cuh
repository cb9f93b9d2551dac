// qg_aux_kernels.cuh — small non-template kernels (record readers, best-rollout reduction); included by qg_engine.cu only.
#pragma once
#include "qg_kernels.cuh"

namespace qg {

// ---- small readers -----------------------------------------------------------------------------------
__global__ void k_read_status(const __grid_constant__ DevCfg c, float* reward, uint8_t* done, uint8_t* success, int32_t* depth) {
    const int64_t env = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= c.B) return;
    const uint32_t d = c.rec[(size_t)HD_DEPTH * c.Bpad + env], f = c.rec[(size_t)HD_FLAGS * c.Bpad + env];
    if (reward) reward[env] = __uint_as_float(c.rec[(size_t)HD_REWARD * c.Bpad + env]);
    if (done) done[env] = (d == 0 || (f & FL_SUCCESS)) ? 1 : 0;
    if (success) success[env] = (f & FL_SUCCESS) ? 1 : 0;
    if (depth) depth[env] = (int32_t)d;
}
__global__ void k_read_metrics(const __grid_constant__ DevCfg c, uint32_t* out) {
    const int64_t env = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= c.B) return;
    const uint32_t l = c.rec[(size_t)HD_LAYERS * c.Bpad + env];
    reinterpret_cast<uint4*>(out)[env] = make_uint4(c.rec[(size_t)HD_NCNOTS * c.Bpad + env], l >> 16, l & 0xFFFFu, c.rec[(size_t)HD_NGATES * c.Bpad + env]);
}
__global__ void k_read_errors(const __grid_constant__ DevCfg c, uint32_t* out) {
    const int64_t env = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= c.B) return;
    out[env] = (c.rec[(size_t)HD_FLAGS * c.Bpad + env] >> FL_ERR_SHIFT) & 0xFFu;
}
__global__ void k_fill_f32(float* p, int64_t n, float v) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
__global__ void k_copy_f32(const float* src, float* dst, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i];
}

// ---- best-rollout reduction (synth search): arg-max of a packed key ----------------------------------
// key = success<<62 | orderable_u32(return)<<30 | (2^30-1 - global rollout id)
__device__ __forceinline__ unsigned long long rollout_key(bool success, float ret, int64_t gid) {
    uint32_t u = __float_as_uint(ret);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);      // order-preserving map f32 -> u32
    return ((unsigned long long)(success ? 1 : 0) << 62) | ((unsigned long long)u << 30) | (unsigned long long)((0x3FFFFFFFll - gid) & 0x3FFFFFFFll);
}
__global__ void k_best(const __grid_constant__ DevCfg c, unsigned long long* best) {
    const int64_t env = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long key = 0;
    if (env < c.B) {
        const uint32_t f = c.rec[(size_t)HD_FLAGS * c.Bpad + env];
        key = rollout_key((f & FL_SUCCESS) != 0, c.ret[env], c.first_id + env);
    }
    for (int o = 16; o > 0; o >>= 1) { const unsigned long long other = __shfl_xor_sync(0xFFFFFFFFu, key, o); key = other > key ? other : key; }
    if ((threadIdx.x & 31) == 0 && key) atomicMax(best, key);
}

}  // namespace qg
