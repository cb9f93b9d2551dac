// qg_oracle_c.cpp — C entry points of the CPU ORACLE (test infrastructure, NOT product code).
// See qg_oracle.hpp for provenance and parity status.  Used through ctypes by tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs only.
#include "qg_oracle.hpp"
#include "../include/qg_engine.h"

#include <atomic>
#include <chrono>
#include <memory>
#include <thread>

using namespace qgo;

namespace {
struct OneShotRng : Rng {   // feeds exactly the injected raw draw
    uint32_t v; bool used = false;
    explicit OneShotRng(uint32_t x) : v(x) {}
    uint32_t next_u32() override { if (used) throw std::runtime_error("injected draw consumed twice"); used = true; return v; }
};
thread_local std::string g_err;

Env* make_env(const qg_config* c) {
    std::vector<Gate> gs;
    for (int i = 0; i < c->num_gates; ++i) gs.push_back(Gate{c->gateset[i].kind, (size_t)c->gateset[i].q0, (size_t)c->gateset[i].q1});
    MetricsWeights w; w.n_cnots = c->w_n_cnots; w.n_layers_cnots = c->w_n_layers_cnots; w.n_layers = c->w_n_layers; w.n_gates = c->w_n_gates;
    const size_t n = (size_t)c->num_qubits;
    switch (c->env_kind) {
        case QG_ENV_PERMUTATION:
            return new Permutation(n, c->difficulty, gs, c->depth_slope, c->max_depth, w, c->add_inverts != 0, c->add_perms != 0, c->track_solution != 0);
        case QG_ENV_LINEAR_FUNCTION:
            return new MatrixEnv(false, n, c->difficulty, gs, c->depth_slope, c->max_depth, w, c->add_inverts != 0, c->add_perms != 0, c->track_solution != 0);
        case QG_ENV_CLIFFORD:
            return new MatrixEnv(true, n, c->difficulty, gs, c->depth_slope, c->max_depth, w, c->add_inverts != 0, c->add_perms != 0, c->track_solution != 0);
        case QG_ENV_PAULI_NETWORK: {
            const int fl = c->final_pauli_layers >= 0 ? c->final_pauli_layers : c->max_rotations + 2;   // pauli.rs:760
            return new PauliEnv(n, c->difficulty, gs, c->depth_slope, c->max_depth, c->max_rotations, c->pauli_diff_scale, c->num_qubits_decay,
                                (size_t)fl, w, c->add_perms != 0, c->pauli_layer_reward, c->track_solution != 0);
        }
    }
    throw std::runtime_error("unknown env kind");
}
}  // namespace

#define QGO_TRY(body) try { body } catch (const std::exception& ex) { g_err = ex.what(); return -1; }

extern "C" {

QG_API const char* qgo_last_error() { return g_err.c_str(); }

QG_API void* qgo_create(const qg_config* cfg) {
    try { return make_env(cfg); } catch (const std::exception& ex) { g_err = ex.what(); return nullptr; }
}
QG_API void* qgo_clone(void* h) { return ((Env*)h)->clone(); }
QG_API void qgo_destroy(void* h) { delete (Env*)h; }
QG_API int qgo_num_actions(void* h) { return (int)((Env*)h)->num_actions(); }
QG_API int qgo_obs_shape(void* h, int32_t out[2]) { auto s = ((Env*)h)->obs_shape(); out[0] = (int32_t)s[0]; out[1] = (int32_t)s[1]; return 0; }
QG_API void qgo_set_difficulty(void* h, int d) { ((Env*)h)->set_difficulty((size_t)d); }
QG_API int qgo_get_difficulty(void* h) { return (int)((Env*)h)->get_difficulty(); }
QG_API int qgo_set_state(void* h, const int64_t* s, int64_t len) {
    QGO_TRY(((Env*)h)->set_state(std::vector<int64_t>(s, s + len)); return 0;)
}
// reset() drawing from Philox(seed; env_id, draw#, STREAM_RESET) — the engine's reset stream.
QG_API int qgo_reset_philox(void* h, uint64_t seed, uint64_t env_id) {
    QGO_TRY(PhiloxRng r(seed, env_id, STREAM_RESET); ((Env*)h)->reset(r); return 0;)
}
// coin: -1 = no coin available (add_inverts=False), 0/1 = injected gen_bool(0.5) result.
QG_API int qgo_step(void* h, int64_t action, int coin) {
    QGO_TRY(
        if (coin < 0) { ((Env*)h)->step((size_t)action, nullptr); }
        else { OneShotRng r(coin ? 0x80000000u : 0u); ((Env*)h)->step((size_t)action, &r); }
        return 0;)
}
// Sparse observation (indices of ones).  has_raw: inject the raw 32-bit draw of the PauliEnv perm pick.
QG_API int qgo_observe(void* h, int64_t* out, int cap, int has_raw, uint32_t raw) {
    QGO_TRY(
        OneShotRng r(raw);
        auto o = ((Env*)h)->observe(has_raw ? &r : nullptr);
        int n = (int)o.size();
        for (int i = 0; i < n && i < cap; ++i) out[i] = (int64_t)o[i];
        return n;)
}
QG_API int qgo_masks(void* h, uint8_t* out) { auto m = ((Env*)h)->masks(); for (size_t i = 0; i < m.size(); ++i) out[i] = m[i] ? 1 : 0; return (int)m.size(); }
QG_API float qgo_reward(void* h) { return ((Env*)h)->reward(); }
QG_API int qgo_is_final(void* h) { return ((Env*)h)->is_final() ? 1 : 0; }
QG_API int qgo_success(void* h) { return ((Env*)h)->success() ? 1 : 0; }
QG_API int64_t qgo_depth(void* h) { return (int64_t)((Env*)h)->depth_left(); }
QG_API int qgo_solution(void* h, int64_t* out, int cap) {
    auto s = ((Env*)h)->solution(); int n = (int)s.size();
    for (int i = 0; i < n && i < cap; ++i) out[i] = (int64_t)s[i];
    return n;
}
QG_API int qgo_raw_state(void* h, uint8_t* out, int cap) {
    auto s = ((Env*)h)->raw_state(); int n = (int)s.size();
    for (int i = 0; i < n && i < cap; ++i) out[i] = s[i];
    return n;
}
QG_API void qgo_counts(void* h, int64_t out[4]) {
    auto c = ((Env*)h)->counts(); out[0] = (int64_t)c.n_cnots; out[1] = (int64_t)c.n_layers_cnots; out[2] = (int64_t)c.n_layers; out[3] = (int64_t)c.n_gates;
}
// twists(): for PauliEnv the public twists() is empty (pauli.rs:675-679); which=1 returns its internal
// qubit_perms/act_perms instead.
QG_API int64_t qgo_twists(void* h, int internal, int64_t* obs_out, int64_t obs_cap, int64_t* act_out, int64_t act_cap, int64_t* obs_len, int64_t* act_len) {
    Env* e = (Env*)h; std::pair<Perms, Perms> t;
    PauliEnv* pe = dynamic_cast<PauliEnv*>(e);
    if (internal && pe) t = {pe->qubit_perms, pe->act_perms}; else t = e->twists();
    *obs_len = t.first.empty() ? 0 : (int64_t)t.first[0].size();
    *act_len = t.second.empty() ? 0 : (int64_t)t.second[0].size();
    int64_t k = 0;
    for (auto& p : t.first) for (size_t v : p) { if (k < obs_cap) obs_out[k] = (int64_t)v; ++k; }
    k = 0;
    for (auto& p : t.second) for (size_t v : p) { if (k < act_cap) act_out[k] = (int64_t)v; ++k; }
    return (int64_t)t.first.size();
}

// ---------------------------------------------------------------------------------------
// Batch differential driver: B envs, per env set_state(target) then T steps with injected
// randomness; records everything the fused GPU step produces.
//   targets   : B payloads, payload b at targets + b*stride, length lens[b]
//   actions   : int32[T][B];  coins: uint8[T][B] or NULL;  perm_raw: uint32[T+1][B] or NULL
//               (perm_raw[0] feeds the observe() after set_state, perm_raw[t+1] the one after step t)
//   obs0      : uint8[B][OBS] dense observation right after set_state (may be NULL)
//   obs       : uint8[T][B][OBS] (may be NULL);  reward f32[T][B]; done/success u8[T][B]
//   counts    : int64[T][B][4] (may be NULL);  depth int64[T][B] (may be NULL)
//   final_state: uint8[B][state_cap] + final_state_len[B];  solutions int64[B][sol_cap] + sol_len[B]
// ---------------------------------------------------------------------------------------
QG_API int qgo_run_batch(const qg_config* cfg, int64_t B, const int64_t* targets, int64_t stride, const int64_t* lens,
                         int32_t T, const int32_t* actions, const uint8_t* coins, const uint32_t* perm_raw,
                         uint8_t* obs0, uint8_t* obs, float* reward, uint8_t* done, uint8_t* success,
                         int64_t* counts, int64_t* depth,
                         uint8_t* final_state, int64_t state_cap, int64_t* final_state_len,
                         int64_t* solutions, int64_t sol_cap, int64_t* sol_len) {
    QGO_TRY(
        std::unique_ptr<Env> proto(make_env(cfg));
        auto shp = proto->obs_shape(); const size_t OBS = shp[0] * shp[1];
        for (int64_t b = 0; b < B; ++b) {
            std::unique_ptr<Env> e(proto->clone());
            e->set_state(std::vector<int64_t>(targets + b * stride, targets + b * stride + lens[b]));
            auto observe = [&](int t) {
                if (perm_raw) { OneShotRng r(perm_raw[(size_t)t * B + b]); return e->observe(&r); }
                return e->observe(nullptr);
            };
            { auto o = observe(0); if (obs0) { uint8_t* dst = obs0 + (size_t)b * OBS; std::memset(dst, 0, OBS); for (size_t i : o) dst[i] = 1; } }
            for (int t = 0; t < T; ++t) {
                const size_t k = (size_t)t * B + b;
                const int32_t a = actions[k];
                if (coins) { OneShotRng r(coins[k] ? 0x80000000u : 0u); e->step((size_t)(int64_t)a, &r); }
                else e->step((size_t)(int64_t)a, nullptr);
                auto o = observe(t + 1);
                if (obs) { uint8_t* dst = obs + k * OBS; std::memset(dst, 0, OBS); for (size_t i : o) dst[i] = 1; }
                reward[k] = e->reward(); done[k] = e->is_final(); success[k] = e->success();
                if (counts) { auto c = e->counts(); counts[k * 4 + 0] = c.n_cnots; counts[k * 4 + 1] = c.n_layers_cnots; counts[k * 4 + 2] = c.n_layers; counts[k * 4 + 3] = c.n_gates; }
                if (depth) depth[k] = (int64_t)e->depth_left();
            }
            if (final_state) { auto s = e->raw_state(); final_state_len[b] = (int64_t)s.size(); for (size_t i = 0; i < s.size() && (int64_t)i < state_cap; ++i) final_state[b * state_cap + i] = s[i]; }
            if (solutions) { auto s = e->solution(); sol_len[b] = (int64_t)s.size(); for (size_t i = 0; i < s.size() && (int64_t)i < sol_cap; ++i) solutions[b * sol_cap + i] = (int64_t)s[i]; }
        }
        return 0;)
}

// ---------------------------------------------------------------------------------------
// CPU baseline driver: the way the reference is driven by twisterl's collectors — one cloned env
// per episode, per step the five trait calls that one fused GPU step replaces
// (step, observe, masks, reward, is_final), each returning freshly allocated vectors — with envs
// statically partitioned over `threads` std::threads (rayon pool over cloned envs in the reference).
// Returns elapsed seconds; *checksum folds the outputs so nothing is optimised away.
// ---------------------------------------------------------------------------------------
QG_API double qgo_bench(const qg_config* cfg, int64_t B, const int64_t* targets, int64_t stride, const int64_t* lens,
                        int32_t T, const int32_t* actions, const uint8_t* coins, int threads, uint64_t* checksum) {
    try {
        std::unique_ptr<Env> proto(make_env(cfg));
        std::vector<std::unique_ptr<Env>> envs;
        for (int64_t b = 0; b < B; ++b) {
            envs.emplace_back(proto->clone());
            envs.back()->set_state(std::vector<int64_t>(targets + b * stride, targets + b * stride + lens[b]));
        }
        if (threads < 1) threads = 1;
        std::vector<uint64_t> sums((size_t)threads, 0);
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> pool;
        for (int th = 0; th < threads; ++th) {
            pool.emplace_back([&, th]() {
                const int64_t lo = B * th / threads, hi = B * (th + 1) / threads; uint64_t acc = 0;
                for (int64_t b = lo; b < hi; ++b) {
                    Env* e = envs[(size_t)b].get();
                    for (int t = 0; t < T; ++t) {
                        const size_t k = (size_t)t * B + b;
                        if (coins) { OneShotRng r(coins[k] ? 0x80000000u : 0u); e->step((size_t)actions[k], &r); }
                        else e->step((size_t)actions[k], nullptr);
                        std::vector<size_t> o = e->observe(nullptr);
                        std::vector<bool> m = e->masks();
                        const float r = e->reward(); const bool f = e->is_final();
                        uint32_t rb; std::memcpy(&rb, &r, 4);
                        acc = acc * 1315423911ull + o.size() + (o.empty() ? 0 : o.back()) + (m.empty() ? 0 : (m[0] ? 1 : 0)) + rb + (f ? 7 : 0);
                    }
                }
                sums[(size_t)th] = acc;
            });
        }
        for (auto& t : pool) t.join();
        auto t1 = std::chrono::steady_clock::now();
        uint64_t tot = 0; for (uint64_t s : sums) tot ^= s;
        if (checksum) *checksum = tot;
        return std::chrono::duration<double>(t1 - t0).count();
    } catch (const std::exception& ex) { g_err = ex.what(); return -1.0; }
}

// ---------------------------------------------------------------------------------------
// Whole-batch digest for parity at BASELINE.json's full sizes (tests/test_full_size.py): one uint64 per environment that folds EVERY output of
// EVERY step — all set observation entries, all mask entries, the reward's bit pattern, is_final, success — so that a 65 536 x 128 run is
// compared env by env without storing 2 GB of observations.  Per step  h = sum(wobs[i] : entry i set) + sum(wmask[j] : action j allowed)
// + reward_bits * 0x9E3779B1 + is_final * 0x85EBCA6B + success * 0xC2B2AE35  and  digest = digest * 0x100000001B3 + h  (mod 2^64; the
// weights are < 2^31 so that the same sums can be formed in int64 tensor arithmetic on the device side).  Envs are statically
// partitioned over `threads` threads; the digest of an env does not depend on the partition.
// ---------------------------------------------------------------------------------------
QG_API int qgo_digest(const qg_config* cfg, int64_t B, const int64_t* targets, int64_t stride, const int64_t* lens,
                      int32_t T, const int32_t* actions, const uint8_t* coins, int threads,
                      const uint64_t* wobs, const uint64_t* wmask, uint64_t* digest) {
    try {
        std::unique_ptr<Env> proto(make_env(cfg));
        if (threads < 1) threads = 1;
        std::vector<std::string> errs((size_t)threads);
        std::vector<std::thread> pool;
        for (int th = 0; th < threads; ++th) {
            pool.emplace_back([&, th]() {
                try {
                    const int64_t lo = B * th / threads, hi = B * (th + 1) / threads;
                    for (int64_t b = lo; b < hi; ++b) {
                        std::unique_ptr<Env> e(proto->clone());
                        e->set_state(std::vector<int64_t>(targets + b * stride, targets + b * stride + lens[b]));
                        uint64_t d = 0;
                        for (int t = 0; t < T; ++t) {
                            const size_t k = (size_t)t * B + b;
                            if (coins) { OneShotRng r(coins[k] ? 0x80000000u : 0u); e->step((size_t)(int64_t)actions[k], &r); }
                            else e->step((size_t)(int64_t)actions[k], nullptr);
                            uint64_t h = 0;
                            for (size_t i : e->observe(nullptr)) h += wobs[i];
                            const std::vector<bool> m = e->masks();
                            for (size_t j = 0; j < m.size(); ++j) if (m[j]) h += wmask[j];
                            const float r = e->reward(); uint32_t rb; std::memcpy(&rb, &r, 4);
                            h += (uint64_t)rb * 0x9E3779B1ull + (e->is_final() ? 0x85EBCA6Bull : 0ull) + (e->success() ? 0xC2B2AE35ull : 0ull);
                            d = d * 0x100000001B3ull + h;
                        }
                        digest[b] = d;
                    }
                } catch (const std::exception& ex) { errs[(size_t)th] = ex.what(); }
            });
        }
        for (auto& t : pool) t.join();
        for (auto& er : errs) if (!er.empty()) { g_err = er; return -1; }
        return 0;
    } catch (const std::exception& ex) { g_err = ex.what(); return -1; }
}

QG_API int qgo_gate_kind_from_name(const char* name, int32_t n_idx) {
    const int k = parse_gate_name(name ? name : "", (size_t)n_idx);
    return k == -1 ? QG_ERR_INVALID : k == -2 ? QG_ERR_STATE : k;
}

QG_API uint32_t qgo_philox_draw(uint64_t seed, uint64_t env, uint32_t idx, uint32_t stream) { return Philox::draw(seed, env, idx, stream); }

}  // extern "C"
