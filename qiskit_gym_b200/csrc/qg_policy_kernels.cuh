// qg_policy_kernels.cuh — device code of the fused policy network (see qg_policy.cu for the design notes); shared by the
// stand-alone kernel k_policy_mlp and the fused search kernel (qg_search_fused.cuh).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace qg {


constexpr int kPolRows = 8;          // batch rows per CTA
constexpr int kPolHalf = 256;         // threads that together own all output features of a layer (feature j = thread + i * 256)
constexpr int kPolHalves = 2;         // the inputs of a layer are split over this many such groups (partial sums combined in shared memory)
constexpr int kPolConsumers = kPolHalf * kPolHalves;   // 16 compute warps
constexpr int kPolMaxWidth = 1024;    // widest layer (4 features per thread)
constexpr int kPolThreads = kPolConsumers + 32;   // + 1 producer warp
constexpr int kPolMaxLayers = 8;
#ifndef QG_POL_STAGES
#define QG_POL_STAGES 3
#endif
constexpr int kPolStages = QG_POL_STAGES;   // weight tiles in flight (4 stages measured: 23.3 vs 22.8 us per decision, no gain)
constexpr int kPolTileFloats = 8192; // 32 KB per weight tile (16 KB tiles: +12 % time per decision; 64 KB x 2 stages: -4 % but the widest networks no longer fit)
constexpr int kPolPartFloats = kPolMaxWidth * 8;   // partial sums handed between the thread groups: 1024 features (or 16 warps x 64 features) x 8 rows

struct PolicyDev {
    int32_t num_layers, obs_size, obs_words;
    int32_t width[kPolMaxLayers];        // output features of layer l (layer 0 consumes the observation)
    int32_t stride[kPolMaxLayers];       // width rounded up to a multiple of 4 floats: row stride of the transposed weights
    const float* wt[kPolMaxLayers];      // transposed weights [in][stride] (layer 0: null, see w0q)
    const float* bias[kPolMaxLayers];
    const int32_t* w0q;                  // first layer's transposed weights in fixed point: w0q[k][j] = rint(W0[j][k] * 2^w0_shift), [obs_size][stride[0]]
    float w0_scale;                      // 2^-w0_shift
    int32_t act0_floats, act1_floats;    // activation buffers ([feature][row])
    int32_t num_actions;                 // outputs of the last layer that are action logits; one more (width - 1) is the value head's output when present
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

__device__ __forceinline__ void cp_async_16(void* sdst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(sdst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// d0 += w * h0, d1 += w * h1 as ONE instruction (fma.rn.f32x2 -> FFMA2 with the weight as a broadcast scalar operand): the same two
// IEEE fused multiply-adds as two fmaf, half the issue slots.  The mov.b64 packs / unpacks are register-pair naming only (ptxas
// emits no instruction for them when the accumulators stay in aligned pairs).
__device__ __forceinline__ void ffma2_bcast(float& d0, float& d1, float w, float h0, float h1) {
    unsigned long long a, b, c;
    asm("mov.b64 %0, {%1, %1};" : "=l"(a) : "f"(w));
    asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(h0), "f"(h1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(d0), "f"(d1));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c) : "l"(a), "l"(b));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(c));
}

// acc[i][r] += tile[u][j_i] * h[u][r] over the weight rows u = u0, u0 + ustep, .. of one tile (h: [u][8] activations, one 32-byte
// broadcast per u)
template <int NI>
__device__ __forceinline__ void fma_tile(float (&acc)[NI][kPolRows], const float* __restrict__ tile, const float4* __restrict__ h, int rows, int ostr, const int (&jc)[NI],
                                         int u0, int ustep) {
#pragma unroll 4
    for (int u = u0; u < rows; u += ustep) {
        const float4 h0 = h[2 * u], h1 = h[2 * u + 1];
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            const float w = tile[u * ostr + jc[i]];
#ifdef QG_POLICY_FFMA1
            acc[i][0] = fmaf(w, h0.x, acc[i][0]); acc[i][1] = fmaf(w, h0.y, acc[i][1]);
            acc[i][2] = fmaf(w, h0.z, acc[i][2]); acc[i][3] = fmaf(w, h0.w, acc[i][3]);
            acc[i][4] = fmaf(w, h1.x, acc[i][4]); acc[i][5] = fmaf(w, h1.y, acc[i][5]);
            acc[i][6] = fmaf(w, h1.z, acc[i][6]); acc[i][7] = fmaf(w, h1.w, acc[i][7]);
#else
            ffma2_bcast(acc[i][0], acc[i][1], w, h0.x, h0.y); ffma2_bcast(acc[i][2], acc[i][3], w, h0.z, h0.w);
            ffma2_bcast(acc[i][4], acc[i][5], w, h1.x, h1.y); ffma2_bcast(acc[i][6], acc[i][7], w, h1.z, h1.w);
#endif
        }
    }
}
__device__ __forceinline__ void store_features(float* __restrict__ dst, int j, const float (&v)[kPolRows], bool relu) {
    float4 lo, hi;
    lo.x = v[0]; lo.y = v[1]; lo.z = v[2]; lo.w = v[3];
    hi.x = v[4]; hi.y = v[5]; hi.z = v[6]; hi.w = v[7];
    if (relu) {
        lo.x = fmaxf(lo.x, 0.f); lo.y = fmaxf(lo.y, 0.f); lo.z = fmaxf(lo.z, 0.f); lo.w = fmaxf(lo.w, 0.f);
        hi.x = fmaxf(hi.x, 0.f); hi.y = fmaxf(hi.y, 0.f); hi.z = fmaxf(hi.z, 0.f); hi.w = fmaxf(hi.w, 0.f);
    }
    reinterpret_cast<float4*>(dst + (size_t)j * kPolRows)[0] = lo;
    reinterpret_cast<float4*>(dst + (size_t)j * kPolRows)[1] = hi;
}

// ---- shared-memory plan of one policy CTA ---------------------------------------------------------------------------------------
struct PolicySmem {
    float* tiles;        // [kPolStages][kPolTileFloats] weight tiles
    float* act0;         // [act0_floats]
    float* act1;         // [act1_floats]
    uint64_t* full;      // [kPolStages]
    uint64_t* empty;     // [kPolStages]
    float* part;         // [kPolPartFloats] partial sums [group][feature][row] of the current layer
    uint32_t* rowbits;   // [kPolRows][obs_words] the rows' packed observations
    uint32_t* oldbits;   // [kPolRows][obs_words] the observations the first layer's accumulators (acc0) currently stand for
    int* wcnt;           // [16] per-warp counts, [31] = U
    uint32_t* uent;      // [policy_ent_cap] per-row lists of the observation entries that changed: index | lost << 31
};
// entries the lists hold: every row's list fits on its own (<= obs_size entries), typical fresh sums of all 8 rows fit together
__host__ __device__ inline int policy_ent_cap(const PolicyDev& p) { return p.obs_size > 2048 ? p.obs_size : (8 * p.obs_size < 2048 ? 8 * p.obs_size : 2048); }
__host__ __device__ inline size_t policy_smem_bytes(const PolicyDev& p) {
    size_t b = ((size_t)kPolStages * kPolTileFloats + p.act0_floats + p.act1_floats + kPolPartFloats) * 4 + (2 * kPolStages + 2) * 8 +
               (size_t)2 * kPolRows * p.obs_words * 4 + 32 * 4 + (size_t)policy_ent_cap(p) * 4;
    return (b + 127) / 128 * 128;
}
__device__ __forceinline__ PolicySmem policy_smem_carve(const PolicyDev& p, float* sm) {
    PolicySmem s;
    s.tiles = sm;
    s.act0 = s.tiles + kPolStages * kPolTileFloats;
    s.act1 = s.act0 + p.act0_floats;
    s.part = s.act1 + p.act1_floats;
    s.full = reinterpret_cast<uint64_t*>(s.part + kPolPartFloats);
    s.empty = s.full + kPolStages;
    s.rowbits = reinterpret_cast<uint32_t*>(s.empty + kPolStages + 2);
    s.oldbits = s.rowbits + kPolRows * p.obs_words;
    s.wcnt = reinterpret_cast<int*>(s.oldbits + kPolRows * p.obs_words);
    s.uent = reinterpret_cast<uint32_t*>(s.wcnt + 32);
    return s;
}
// The first layer: h[r][j] = act(bias[j] + sum over the set observation entries k of row r of W0[j][k]), as an EXACT sum.  The weights
// are fixed-point integers (w0q, 31 significant bits below the largest |weight|: finer than the f32 weights' own 24 bits there) and the
// accumulators 64-bit integers, so the sum does not depend on the order of its terms.  That is what lets the one-launch search update
// it instead of recomputing it: between two decisions of a rollout only a few entries change (a SWAP moves two of a permutation's 27
// one-hot positions), so the CTA adds the weight rows of the entries that appeared and subtracts those of the entries that vanished
// (`acc0`: its accumulators [row][feature] in global memory, L2 resident) — the same integers a fresh sum over all set entries gives,
// bit for bit (the stand-alone kernel and the first decision do exactly that: acc0 == nullptr / fresh).  ~30 weight rows per decision
// instead of ~200, fetched with plain coalesced loads, no shared-memory staging.
// The changed entries are found word by word (new ^ old of the packed observations, one (row, word) pair per thread) and appended to
// per-row lists with shared-memory atomics — any order will do, the sum is exact.  ent[]: entry index | lost << 31, row r's list at
// ent[off[r] .. off[r] + cnt[r]).  Thread t owns features t, t + 512 and keeps row r's accumulator in a register of its own.  When the
// lists of all 8 rows do not fit `cap` entries (dense observations summed afresh), rows are processed in groups.
// stream != nullptr: the rows' packed observations are read from the step kernel's own [word][33] shared-memory bit stream (row r of the
// CTA = environment r of the tile, rows >= stream_rows are empty) instead of ps.rowbits.
__device__ __forceinline__ void layer0_fixed(const PolicyDev& p, const PolicySmem& ps, float* __restrict__ dst, long long* __restrict__ acc0, bool fresh, bool relu,
                                             int tid, const uint32_t* __restrict__ stream = nullptr, int stream_rows = 0) {
    const int out = p.width[0], ostr = p.stride[0], OW = p.obs_words, NW = kPolRows * OW, cap = policy_ent_cap(p);
    const int32_t* __restrict__ wq = p.w0q;
    const uint32_t* __restrict__ rowbits = ps.rowbits; uint32_t* __restrict__ oldbits = ps.oldbits; uint32_t* __restrict__ ent = ps.uent;
    int* const cnt = ps.wcnt; int* const fill = ps.wcnt + kPolRows;        // per-row list lengths / append cursors
    const uint32_t last_mask = (p.obs_size & 31) ? ((1u << (p.obs_size & 31)) - 1u) : 0xFFFFFFFFu;
    auto now_at = [&](int i, int r, int w) -> uint32_t {
        if (!stream) return rowbits[i];
        const uint32_t v = r < stream_rows ? stream[w * 33 + r] : 0u;
        return w == OW - 1 ? (v & last_mask) : v;
    };
    if (tid < 2 * kPolRows) cnt[tid] = 0;
    consumers_sync();
    for (int i = tid; i < NW; i += kPolConsumers) {
        const int r = i / OW;
        const uint32_t d = now_at(i, r, i - r * OW) ^ (fresh ? 0u : oldbits[i]);
        if (d) atomicAdd(&cnt[r], __popc(d));
    }
    consumers_sync();
    for (int jb = 0; jb < out; jb += kPolConsumers) {
        const int j = jb + tid;
        const bool mine = j < out;
        long long acc[kPolRows];
#pragma unroll
        for (int r = 0; r < kPolRows; ++r) acc[r] = (mine && acc0 && !fresh) ? acc0[(size_t)r * out + j] : 0ll;
        for (int r0 = 0; r0 < kPolRows;) {
            // rows [r0, r1): as many whole rows as fit the list (one row always does: cap >= obs_size)
            int r1 = r0, tot = 0, off[kPolRows + 1];
#pragma unroll
            for (int r = 0; r < kPolRows; ++r) {
                off[r] = tot;
                if (r >= r0 && r == r1 && (r == r0 || tot + cnt[r] <= cap)) { tot += cnt[r]; r1 = r + 1; }
            }
            off[kPolRows] = tot;
            if (tot > 0) {
                for (int i = tid; i < NW; i += kPolConsumers) {
                    const int r = i / OW, w = i - r * OW;
                    if (r < r0 || r >= r1) continue;
                    const uint32_t was = fresh ? 0u : oldbits[i];
                    uint32_t d = now_at(i, r, w) ^ was;
                    if (d) {
                        int at = off[r] + atomicAdd(&fill[r], __popc(d));
                        while (d) {
                            const int bit = __ffs(d) - 1;
                            d &= d - 1;
                            ent[at++] = (uint32_t)(w * 32 + bit) | (((was >> bit) & 1u) << 31);
                        }
                    }
                }
                consumers_sync();
                if (mine) {
                    // position t of every row's list in one batch: up to 16 independent loads in flight per thread, so the number of
                    // L2 round trips is half the longest list, not the sum of the lists
                    int c[kPolRows], tmax = 0;
#pragma unroll
                    for (int r = 0; r < kPolRows; ++r) { c[r] = (r >= r0 && r < r1) ? cnt[r] : 0; tmax = max(tmax, c[r]); }
                    for (int t = 0; t < tmax; t += 2) {
                        int w[2][kPolRows];
#pragma unroll
                        for (int i = 0; i < 2; ++i)
#pragma unroll
                            for (int r = 0; r < kPolRows; ++r) {
                                w[i][r] = 0;
                                if (t + i < c[r]) {
                                    const uint32_t e = ent[off[r] + t + i];
                                    const int v = __ldg(wq + (size_t)(e & 0x7FFFFFFFu) * ostr + j);
                                    w[i][r] = (e >> 31) ? -v : v;
                                }
                            }
#pragma unroll
                        for (int i = 0; i < 2; ++i)
#pragma unroll
                            for (int r = 0; r < kPolRows; ++r) acc[r] += (long long)w[i][r];
                    }
                }
            }
            r0 = r1;
            if (r0 < kPolRows || jb + kPolConsumers < out) {          // the lists are rebuilt: everybody is done reading them, cursors back to zero
                consumers_sync();
                if (tid < kPolRows) fill[tid] = 0;
                consumers_sync();
            }
        }
        if (mine) {
            float v[kPolRows];
            const float b = __ldg(p.bias[0] + j);
#pragma unroll
            for (int r = 0; r < kPolRows; ++r) {
                if (acc0) acc0[(size_t)r * out + j] = acc[r];
                v[r] = __fadd_rn(b, __fmul_rn((float)acc[r], p.w0_scale));       // (float)int64: one rounding; the scale is a power of two
            }
            store_features(dst, j, v, relu);
        }
    }
    if (acc0) {                                    // the accumulators now stand for these observations
        consumers_sync();
        for (int i = tid; i < NW; i += kPolConsumers) { const int r = i / OW; oldbits[i] = now_at(i, r, i - r * OW); }
    }
}

// One dense layer for the consumer warps: the 512 compute threads are two halves of 256; within a half, thread ht owns the output
// features ht + i * 256 (NI of them) for all 8 rows, and half h takes the inputs u = h, h + 2, .. of every tile; half 1 hands its
// partial sums to half 0 through shared memory.  Tiles arrive from the producer warp (full / empty mbarriers).
// (A mapping with 4 consecutive features per thread — one LDS.128 for the weights, 3 loads per 32 FMAs — was measured 40 % slower:
// its per-tile bookkeeping outweighs the saved shared-memory traffic at 2 inputs per group and tile.)
template <int NI>
__device__ __forceinline__ void consume_layer(const PolicyDev& p, int l, int K, const float* __restrict__ src, float* __restrict__ dst, float* tiles,
                                              uint64_t* full, uint64_t* empty, float* part, int& G, int tid, int lane) {
    const int out = p.width[l], ostr = p.stride[l], kt = tile_rows(p, l), nt = (K + kt - 1) / kt;
    const int half = tid / kPolHalf, ht = tid - half * kPolHalf;
    float acc[NI][kPolRows];
    int jc[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) {
        const int j = ht + i * kPolHalf;
        jc[i] = min(j, ostr - 1);                          // out-of-range features read a valid word and are never written
        const float b = (half == 0 && j < out) ? __ldg(p.bias[l] + j) : 0.0f;
#pragma unroll
        for (int r = 0; r < kPolRows; ++r) acc[i][r] = b;
    }
    const bool active = NI > 1 || (ht & ~31) < out;        // warp-uniform: this warp owns at least one real feature
    for (int t = 0; t < nt; ++t, ++G) {
        const int stage = G % kPolStages;
        mbar_wait(full + stage, (uint32_t)((G / kPolStages) & 1));
        if (active) {
            const int k0 = t * kt;
            fma_tile<NI>(acc, tiles + (size_t)stage * kPolTileFloats, reinterpret_cast<const float4*>(src + (size_t)k0 * kPolRows), min(kt, K - k0), ostr, jc, half,
                         kPolHalves);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + stage);         // this warp is done with the stage
    }
    // combine the halves' partial sums (half 1 -> shared memory -> half 0), activation, store
    if (half == 1) {
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            const int j = ht + i * kPolHalf;
            if (j < out) store_features(part, j, acc[i], false);
        }
    }
    consumers_sync();
    if (half == 0) {
        const bool last = l == p.num_layers - 1;
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            const int j = ht + i * kPolHalf;
            if (j < out) {
                const float4 lo = reinterpret_cast<const float4*>(part + (size_t)j * kPolRows)[0], hi = reinterpret_cast<const float4*>(part + (size_t)j * kPolRows)[1];
                acc[i][0] += lo.x; acc[i][1] += lo.y; acc[i][2] += lo.z; acc[i][3] += lo.w;
                acc[i][4] += hi.x; acc[i][5] += hi.y; acc[i][6] += hi.z; acc[i][7] += hi.w;
                store_features(dst, j, acc[i], !last);
            }
        }
    }
}

// A layer of 65..256 output features (the BasicPolicy's second layer): one feature per thread would leave the loop bound by its
// shared-memory loads (two 16-byte activation broadcasts + one weight per 4 FFMA2), so a thread owns TWO adjacent features (one 8-byte
// weight load, 3 loads per 8 FFMA2) and the inputs are split over four groups of 128 threads (group g takes u = g, g + 4, .. of every
// tile); groups 1..3 hand their partial sums to group 0 through shared memory, which adds them in group order.
__device__ __forceinline__ void consume_layer_pairs(const PolicyDev& p, int l, int K, const float* __restrict__ src, float* __restrict__ dst, const float* tiles,
                                                    uint64_t* full, uint64_t* empty, float* part, int& G, int tid, int lane) {
    constexpr int kGroups = 4, kGroupThreads = kPolConsumers / kGroups;       // 4 x 128
    const int out = p.width[l], ostr = p.stride[l], kt = tile_rows(p, l), nt = (K + kt - 1) / kt;
    const int grp = tid / kGroupThreads, gt = tid - grp * kGroupThreads;
    const int j0 = 2 * gt, jc = min(j0, ostr - 2);                             // out-of-range pairs read valid words and are never written
    float acc[2][kPolRows];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float b = (grp == 0 && j0 + i < out) ? __ldg(p.bias[l] + j0 + i) : 0.0f;
#pragma unroll
        for (int r = 0; r < kPolRows; ++r) acc[i][r] = b;
    }
    const bool active = (j0 & ~63) < out;                                      // warp-uniform: this warp owns at least one real feature
    for (int t = 0; t < nt; ++t, ++G) {
        const int stage = G % kPolStages;
        mbar_wait(full + stage, (uint32_t)((G / kPolStages) & 1));
        if (active) {
            const int k0 = t * kt, rows = min(kt, K - k0);
            const float* __restrict__ tile = tiles + (size_t)stage * kPolTileFloats;
            const float4* __restrict__ h = reinterpret_cast<const float4*>(src + (size_t)k0 * kPolRows);
#pragma unroll 4
            for (int u = grp; u < rows; u += kGroups) {
                const float4 h0 = h[2 * u], h1 = h[2 * u + 1];
                const float2 w = *reinterpret_cast<const float2*>(tile + u * ostr + jc);
                ffma2_bcast(acc[0][0], acc[0][1], w.x, h0.x, h0.y); ffma2_bcast(acc[0][2], acc[0][3], w.x, h0.z, h0.w);
                ffma2_bcast(acc[0][4], acc[0][5], w.x, h1.x, h1.y); ffma2_bcast(acc[0][6], acc[0][7], w.x, h1.z, h1.w);
                ffma2_bcast(acc[1][0], acc[1][1], w.y, h0.x, h0.y); ffma2_bcast(acc[1][2], acc[1][3], w.y, h0.z, h0.w);
                ffma2_bcast(acc[1][4], acc[1][5], w.y, h1.x, h1.y); ffma2_bcast(acc[1][6], acc[1][7], w.y, h1.z, h1.w);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + stage);
    }
    // part: [group - 1][256 features][8 rows]
    if (grp > 0) {
#pragma unroll
        for (int i = 0; i < 2; ++i) if (j0 + i < out) store_features(part + (size_t)(grp - 1) * 256 * kPolRows, j0 + i, acc[i], false);
    }
    consumers_sync();
    if (grp == 0) {
        const bool last = l == p.num_layers - 1;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int j = j0 + i;
            if (j < out) {
#pragma unroll
                for (int g = 0; g < kGroups - 1; ++g) {
                    const float4 lo = reinterpret_cast<const float4*>(part + ((size_t)g * 256 + j) * kPolRows)[0], hi = reinterpret_cast<const float4*>(part + ((size_t)g * 256 + j) * kPolRows)[1];
                    acc[i][0] += lo.x; acc[i][1] += lo.y; acc[i][2] += lo.z; acc[i][3] += lo.w;
                    acc[i][4] += hi.x; acc[i][5] += hi.y; acc[i][6] += hi.z; acc[i][7] += hi.w;
                }
                store_features(dst, j, acc[i], !last);
            }
        }
    }
}

// A narrow layer (at most 64 output features, e.g. the action head): with one feature per thread only two warps would work and the
// layer would be a latency-bound chain over its K inputs, so the inputs are split over the 16 warps instead (warp w takes the rows
// u = w, w+16, .. of every tile; lane j owns features j and j+32), the partial sums are combined through shared memory in warp
// order and the bias is added last.
__device__ __forceinline__ void consume_layer_narrow(const PolicyDev& p, int l, int K, const float* __restrict__ src, float* __restrict__ dst, const float* tiles,
                                                     uint64_t* full, uint64_t* empty, float* part, int& G, int tid, int lane) {
    const int out = p.width[l], ostr = p.stride[l], kt = tile_rows(p, l), nt = (K + kt - 1) / kt, warp = tid >> 5;
    float acc[2][kPolRows];
    const int jc[2] = {min(lane, ostr - 1), min(lane + 32, ostr - 1)};
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int r = 0; r < kPolRows; ++r) acc[i][r] = 0.0f;
    for (int t = 0; t < nt; ++t, ++G) {
        const int stage = G % kPolStages;
        mbar_wait(full + stage, (uint32_t)((G / kPolStages) & 1));
        const int k0 = t * kt;
        fma_tile<2>(acc, tiles + (size_t)stage * kPolTileFloats, reinterpret_cast<const float4*>(src + (size_t)k0 * kPolRows), min(kt, K - k0), ostr, jc, warp, kPolConsumers / 32);
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + stage);
    }
    // part: [warp][64 features][8 rows]
#pragma unroll
    for (int i = 0; i < 2; ++i) store_features(part + (size_t)warp * 64 * kPolRows, lane + 32 * i, acc[i], false);
    consumers_sync();
    if (tid < out) {
        float v[kPolRows];
        const float b = __ldg(p.bias[l] + tid);
#pragma unroll
        for (int r = 0; r < kPolRows; ++r) v[r] = 0.0f;
        for (int w = 0; w < kPolConsumers / 32; ++w) {
            const float4 lo = reinterpret_cast<const float4*>(part + ((size_t)w * 64 + tid) * kPolRows)[0];
            const float4 hi = reinterpret_cast<const float4*>(part + ((size_t)w * 64 + tid) * kPolRows)[1];
            v[0] += lo.x; v[1] += lo.y; v[2] += lo.z; v[3] += lo.w; v[4] += hi.x; v[5] += hi.y; v[6] += hi.z; v[7] += hi.w;
        }
#pragma unroll
        for (int r = 0; r < kPolRows; ++r) v[r] += b;
        store_features(dst, tid, v, l != p.num_layers - 1);
    }
}

#ifdef QG_SEARCH_PROBE       // tools build: cycles thread 0 of CTA 0 spends per part of a decision of the one-launch search (qg_search_debug_read)
static __device__ long long g_search_prof[16];      // (one copy per translation unit; printed by k_search_fused) [0] observation words, [1 + l] layer l (with its barrier), [9] soft-max, [10] env step, [11] decisions, [12] total
#define QG_SP_T0() const long long _sp0 = clock64()
#define QG_SP_ADD(i) do { if (threadIdx.x == 0 && blockIdx.x == 0) g_search_prof[i] += clock64() - _sp0; } while (0)
#else
#define QG_SP_T0() do { } while (0)
#define QG_SP_ADD(i) do { } while (0)
#endif
__device__ __forceinline__ void policy_init_barriers(const PolicySmem& ps, int tid) {
    if (tid == 0) {
        for (int s = 0; s < kPolStages; ++s) { mbar_init(ps.full + s, 1); mbar_init(ps.empty + s, kPolConsumers / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
}

// One forward pass for the CTA's kPolRows batch rows starting at row0; every one of the kPolThreads threads calls it.  Warps 0..15
// compute (consumers); warp 16 is the producer: one lane streams the weight tiles of the layers after the first, in order, through a
// ring of kPolStages shared-memory stages (full / empty mbarriers) — it starts right away, so the second layer's first tiles land
// while the consumers are still busy with the first layer (whose weights they read themselves, layer0_fixed).
// G is the running tile counter of the ring (same value in every thread; it carries over when the function is called again) and
// `pass` counts the calls (0, 1, 2, ..).
// `bits` may have been written earlier by this CTA in the same kernel: it is read with ld.global.cg, never through the
// non-coherent path.
// bits == nullptr: the rows' packed observations are in `stream` (the step's shared-memory bit stream of the fused search kernel), or,
// without a stream, already in ps.rowbits.
// acc0: the CTA's first-layer accumulators [kPolRows][width[0]] (int64, global memory) when the caller evaluates the same rows again
// and again (the one-launch search): pass 0 sums every set entry, later passes only apply what changed since the previous pass.
// nullptr: a fresh sum, nothing kept.
// probs_rows: row stride of `probs` is the action count and row r of the CTA goes to probs + (row_base + r) * A, with
// row_base = row0 for the global [B][A] tensor or 0 for a CTA-private (shared-memory) buffer.
__device__ __forceinline__ void policy_forward_rows(const PolicyDev& p, const PolicySmem& ps, const uint32_t* bits, int64_t row0, int64_t B,
                                                    float* probs, float* logits_out, int& G, int pass = 0, int64_t probs_row_base = -1,
                                                    long long* acc0 = nullptr, const uint32_t* stream = nullptr, int stream_rows = 0, float* values_out = nullptr) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool producer = tid >= kPolConsumers;
    float* const tiles = ps.tiles; float* const act0 = ps.act0; float* const act1 = ps.act1;
    uint64_t* const full = ps.full; uint64_t* const empty = ps.empty;
    uint32_t* const rowbits = ps.rowbits;

    // ---- producer: the contiguous weight rows of the layers after the first, tile after tile
    if (producer) {
        int total = 0;
        for (int l = 1; l < p.num_layers; ++l) total += layer_tiles(p, l, 0);
        if (lane == 0 && total > 0) {
            int g = G;
            for (int l = 1; l < p.num_layers; ++l) {
                const int K = p.width[l - 1], kt = tile_rows(p, l), nt = (K + kt - 1) / kt;
                const uint32_t row_bytes = (uint32_t)p.stride[l] * 4u;
                for (int t = 0; t < nt; ++t, ++g) {
                    const int stage = g % kPolStages, k0 = t * kt, rows = min(kt, K - k0);
                    if (g >= kPolStages) mbar_wait(empty + stage, (uint32_t)(((g / kPolStages) - 1) & 1));
                    mbar_expect_tx(full + stage, row_bytes * (uint32_t)rows);
                    bulk_g2s(tiles + (size_t)stage * kPolTileFloats, p.wt[l] + (size_t)k0 * p.stride[l], row_bytes * (uint32_t)rows, full + stage);
                }
            }
        }
        G += total;
        return;
    }

    // ---- 1. consumers: the rows' packed observations into shared memory
    const bool fresh = acc0 == nullptr || pass == 0;
    { QG_SP_T0();
    if (bits) {
        for (int i = tid; i < kPolRows * p.obs_words; i += kPolConsumers) {
            const int r = i / p.obs_words, w = i - r * p.obs_words;
            uint32_t word = (row0 + r < B) ? __ldcg(bits + (size_t)(row0 + r) * p.obs_words + w) : 0u;
            if (w == p.obs_words - 1 && (p.obs_size & 31)) word &= (1u << (p.obs_size & 31)) - 1u;
            rowbits[i] = word;
        }
    }
    QG_SP_ADD(0); }

    // ---- 2. layers
    float* src = act1;
    float* dst = act0;
    for (int l = 0; l < p.num_layers; ++l) {
        const int ni = (p.width[l] + kPolHalf - 1) / kPolHalf;
        QG_SP_T0();
        if (l == 0) {
            layer0_fixed(p, ps, dst, acc0, fresh, p.num_layers > 1, tid, bits ? nullptr : stream, stream_rows);
        } else if (p.width[l] <= 64) {
            consume_layer_narrow(p, l, p.width[l - 1], src, dst, tiles, full, empty, ps.part, G, tid, lane);
        } else if (p.width[l] <= 256) {
            consume_layer_pairs(p, l, p.width[l - 1], src, dst, tiles, full, empty, ps.part, G, tid, lane);
        } else {
            const int K = p.width[l - 1];
            switch (ni) {
                case 1: consume_layer<1>(p, l, K, src, dst, tiles, full, empty, ps.part, G, tid, lane); break;
                case 2: consume_layer<2>(p, l, K, src, dst, tiles, full, empty, ps.part, G, tid, lane); break;
                case 3: consume_layer<3>(p, l, K, src, dst, tiles, full, empty, ps.part, G, tid, lane); break;
                default: consume_layer<4>(p, l, K, src, dst, tiles, full, empty, ps.part, G, tid, lane); break;
            }
        }
        consumers_sync();
        QG_SP_ADD(1 + l);
        src = dst;
        dst = (dst == act0) ? act1 : act0;
    }

    // ---- 3. softmax over the action logits, warp r <-> row r -------------------------------------------------------------
    QG_SP_T0();
    if (warp < kPolRows) {
        const int64_t row = row0 + warp;
        if (row < B) {
            const int A = p.num_actions;           // (a value head, when present, is output A of the last layer: not part of the softmax)
            if (values_out && lane == 0) values_out[row] = p.width[p.num_layers - 1] > A ? src[(size_t)A * kPolRows + warp] : 0.0f;
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
                if (probs) probs[(size_t)((probs_row_base < 0 ? row0 : probs_row_base) + warp) * A + a] = expf(lg - mx) * inv;
                if (logits_out) logits_out[(size_t)row * A + a] = lg;
            }
        }
    }
    QG_SP_ADD(9);
}

}  // namespace qg
