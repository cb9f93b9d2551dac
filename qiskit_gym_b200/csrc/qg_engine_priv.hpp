// qg_engine_priv.hpp — the engine object behind the opaque qg_engine handle, shared by the translation units that implement the C ABI
// (qg_engine.cu: lifetime, state in, the fused step; qg_extras.cu: bulk read-out, packed host wire format, NCCL finish, DLPack).
#pragma once
#include <string>
#include <vector>

#include "qg_host.hpp"
#include "qg_launch.hpp"

struct qg_twists { qg::Twists t; };

struct qg_engine {
    qg_config cfg{};                 // gateset pointer re-targeted at `gates`
    std::vector<qg_gate> gates;
    qg::Layout L;
    qg::DevCfg dc{};
    int device = 0;
    int64_t B = 0, Bpad = 0;
    bool owns_ws = false;
    uint8_t* ws = nullptr;
    // workspace carve-outs
    uint32_t* staged = nullptr;      // [Bpad][PW]
    uint32_t* snap = nullptr;        // [W][Bpad] snapshot of the records
    bool has_snap = false;
    int32_t* io_actions = nullptr; uint8_t* io_coins = nullptr; float* io_reward = nullptr; uint8_t* io_done = nullptr; uint8_t* io_success = nullptr;
    unsigned long long* best = nullptr;
    // pinned host staging
    uint32_t* h_staged = nullptr; int64_t h_staged_words = 0;
    unsigned long long* h_best = nullptr;
    size_t smem_bytes = 0; int sm_warp_words = 0, sm_scr = 0, sm_obs = 0;      // 32-env tile layout without the concatenated stream (the one-launch search's)
    int cat_words = 0;               // words of a 32-env tile's concatenated observation stream (0: this config does not use expand_cat)
    int epw_forced = 0;              // (tools builds only: 16 / 32 forces the tile size)
    int replay_ctas = 0;             // (tools builds only, QG_REPLAY_CTAS: CTAs an SM may hold in replay launches, -1 = whatever fits; 0 = launch_step's rule)
    int step_ctas = 0;               // (tools builds only, QG_STEP_CTAS: the same bound for single-step launches)
    int pair_forced = 0;             // (tools builds only, QG_PAIR=0 -> -1: no warp pairs in replay launches)
    uint64_t magic_obs = 0, magic_A = 0; uint32_t magic_vpe = 0, magic_a4 = 0;
    int nperms = 0;
    int pdl_mode = 2;                // programmatic dependent launch variant (see StepArgs); 2 = dependents launch once this grid owns the records
    int stagger_ns = 0, num_sms = 148;
    size_t l2_persist_bytes = 0;     // (tools builds only, -DQG_TOOLS_KNOBS: persisting-L2 window over the replay's action stream)
    bool inv_bucket_enabled = true;  // (tools builds only: false forces the generic shared-memory Gauss-Jordan)
    bool all_symplectic = true;      // Clifford: every state loaded so far is symplectic (identity at construction, resets, checked set_state payloads)
    // qg_replay_host pipeline for pageable buffers (allocated on first use): two chunk buffers, copy-in / copy-out streams
    int rp_chunk = 0;
    int32_t* rp_actions[2] = {nullptr, nullptr}; uint8_t* rp_coins[2] = {nullptr, nullptr};
    float* rp_reward[2] = {nullptr, nullptr}; uint8_t* rp_done[2] = {nullptr, nullptr}; uint8_t* rp_success[2] = {nullptr, nullptr};
    cudaStream_t rp_in = nullptr, rp_out = nullptr;
    cudaEvent_t rp_ev_in[2] = {nullptr, nullptr}, rp_ev_run[2] = {nullptr, nullptr}, rp_ev_out[2] = {nullptr, nullptr}, rp_ev_start = nullptr;
    // qg_replay_host_packed: device staging of the host action / coin streams, filled by the copy engine while the kernel runs (qg_extras.cu)
    uint8_t* hp_act[2] = {nullptr, nullptr}; uint8_t* hp_coin[2] = {nullptr, nullptr}; size_t hp_cap = 0;      // two slots, used in turn
    uint32_t* hp_flags = nullptr;                    // device: [2 slots][4] chunk-arrived flags
    uint32_t* hp_ones = nullptr;                     // pinned host: four ones (the value a flag is raised to), four zeros (what clears a slot's flags)
    cudaStream_t hp_stream = nullptr;
    cudaEvent_t hp_ev_done[2] = {nullptr, nullptr}, hp_ev_zero[2] = {nullptr, nullptr};      // a slot's last launch is over / its flags are cleared
    bool hp_used[2] = {false, false}; int hp_next = 0;
    // qg_extras.cu (allocated on first use, freed by qg_destroy through qg_extras_release)
    uint32_t* fin_send = nullptr; uint32_t* fin_recv = nullptr; int fin_cap = 0, fin_world = 0;   // qg_search_finish exchange buffers
    uint32_t* h_fin = nullptr;                                                                     // pinned copy of the winning row
    uint32_t* bulk_sol = nullptr; int32_t* bulk_len = nullptr; int64_t bulk_count = 0; int bulk_cap = 0;   // qg_solutions_host staging
    float* dl_obs = nullptr; int dl_ring = 0;                                                      // engine-owned observation ring (qg_dlpack_obs)
};

namespace qg {

#define CUDA_OK(expr)                                                                          \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                     \
            return QG_ERR_CUDA;                                                                \
        }                                                                                      \
    } while (0)

// One launch of the fused step kernel over the engine's batch (qg_engine.cu).  logical_batch >= 0: the launch covers that many
// logical envs addressed through a.src_slot / a.dst_slot (qg_step_slots).
int launch_step(qg_engine* e, int mode, StepArgs a, cudaStream_t st, int64_t logical_batch = -1);
// The device-side address of a pinned (page-locked) host buffer, or nullptr for pageable memory.
void* mapped_host(const void* h);
// frees what qg_extras.cu allocated lazily
void extras_release(qg_engine* e);

}  // namespace qg
