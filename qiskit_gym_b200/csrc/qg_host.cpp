// qg_host.cpp — host-only logic of the engine (no CUDA calls): validation, layout, twists, packers.
#include "qg_host.hpp"

#include <algorithm>
#include <cctype>
#include <cstring>
#include <map>
#include <numeric>

#include "qg_common.cuh"

namespace qg {

static thread_local std::string g_error;
void set_error(const std::string& msg) { g_error = msg; }
const char* get_error() { return g_error.c_str(); }

static inline bool two_qubit(int kind) { return kind == QG_CX || kind == QG_CZ || kind == QG_SWAP; }

int validate_config(const qg_config* cfg) {
    if (!cfg) { set_error("null config"); return QG_ERR_INVALID; }
    if (cfg->env_kind < QG_ENV_PERMUTATION || cfg->env_kind > QG_ENV_PAULI_NETWORK) { set_error("unknown env kind"); return QG_ERR_INVALID; }
    const int n = cfg->num_qubits;
    if (n < 1) { set_error("num_qubits must be >= 1"); return QG_ERR_INVALID; }
    if (n > 255) { set_error("num_qubits > 255 is not supported by the engine"); return QG_ERR_UNSUPPORTED; }
    if (cfg->num_gates < 1 || !cfg->gateset) { set_error("gateset must not be empty"); return QG_ERR_INVALID; }
    if (cfg->num_gates > 65535) { set_error("more than 65535 actions are not supported"); return QG_ERR_UNSUPPORTED; }
    if (cfg->difficulty < 0 || cfg->depth_slope < 0 || cfg->max_depth < 0) { set_error("difficulty, depth_slope and max_depth must be non-negative"); return QG_ERR_INVALID; }
    for (int i = 0; i < cfg->num_gates; ++i) {
        const qg_gate& g = cfg->gateset[i];
        if (g.kind < QG_H || g.kind > QG_SWAP) { set_error("gate " + std::to_string(i) + ": unknown gate kind"); return QG_ERR_INVALID; }
        if (g.q0 < 0 || g.q0 >= n || (two_qubit(g.kind) && (g.q1 < 0 || g.q1 >= n))) {
            set_error("gate " + std::to_string(i) + ": qubit index out of range for " + std::to_string(n) + " qubits");
            return QG_ERR_INVALID;
        }
    }
    switch (cfg->env_kind) {
        case QG_ENV_LINEAR_FUNCTION: if (n > 64) { set_error("LinearFunction: num_qubits > 64 is not supported"); return QG_ERR_UNSUPPORTED; } break;
        case QG_ENV_CLIFFORD: if (n > 32) { set_error("Clifford: num_qubits > 32 is not supported"); return QG_ERR_UNSUPPORTED; } break;
        case QG_ENV_PAULI_NETWORK: {
            const int mr = std::max(cfg->max_rotations, 1);
            const int fl = cfg->final_pauli_layers >= 0 ? cfg->final_pauli_layers : cfg->max_rotations + 2;
            const int rt = std::max(mr, fl);
            if (rt > kMaxRot) { set_error("PauliNetwork: more than 16 rotations per env are not supported"); return QG_ERR_UNSUPPORTED; }
            if (2 * n + rt > 64) { set_error("PauliNetwork: 2*num_qubits + rotations > 64 is not supported"); return QG_ERR_UNSUPPORTED; }
            break;
        }
        default: break;
    }
    if (cfg->solution_capacity < 0 || cfg->solution_capacity > 65535) { set_error("solution_capacity must be in [0, 65535]"); return QG_ERR_INVALID; }
    if (cfg->tile_envs != 0 && cfg->tile_envs != 16 && cfg->tile_envs != 32) { set_error("tile_envs must be 0 (automatic), 16 or 32"); return QG_ERR_INVALID; }
    return QG_OK;
}

int make_layout(const qg_config* cfg, Layout& L) {
    const int rc = validate_config(cfg);
    if (rc != QG_OK) return rc;
    L = Layout();
    L.kind = cfg->env_kind; L.n = cfg->num_qubits; L.A = cfg->num_gates;
    const int n = L.n;
    L.MW = (n + 1) / 2;
    switch (L.kind) {
        case QG_ENV_PERMUTATION:
            L.D = n; L.obs_rows = n; L.obs_cols = n; L.SW = (n + 3) / 4; L.state_len = n;
            L.SCR = cfg->add_inverts ? L.SW : 0;
            break;
        case QG_ENV_LINEAR_FUNCTION:
        case QG_ENV_CLIFFORD:
            L.D = (L.kind == QG_ENV_CLIFFORD) ? 2 * n : n; L.obs_rows = L.obs_cols = L.D;
            L.SW = (L.D * L.D + 31) / 32; L.state_len = (int64_t)L.D * L.D;
            L.SCR = cfg->add_inverts ? 2 * L.SW : 0;
            break;
        case QG_ENV_PAULI_NETWORK: {
            L.max_rot = std::max(cfg->max_rotations, 1);                                     // pauli.rs:388
            const int fl = cfg->final_pauli_layers >= 0 ? cfg->final_pauli_layers : cfg->max_rotations + 2;
            L.Rtot = std::max(L.max_rot, fl);
            L.D = 2 * n; L.CW = (2 * n + L.Rtot + 31) / 32 * 32;     // row stride in bits: rows are padded to whole words
            L.obs_rows = 2 * n; L.obs_cols = 2 * n + L.max_rot;
            L.SW = (2 * n * L.CW + 31) / 32;
            L.XW = PX_ANTI + (L.Rtot + 1) / 2;
            L.SCR = L.Rtot;
            L.state_len = 0;
            break;
        }
    }
    L.obs_size = L.obs_rows * L.obs_cols;
    // observation bit stream built next to the state: always for PauliNetwork; for Permutation (one-hot rows) while it
    // fits comfortably in shared memory (n <= 64), else the expander tests the packed bytes directly
    L.OW = (L.kind == QG_ENV_PAULI_NETWORK || (L.kind == QG_ENV_PERMUTATION && n <= 64)) ? (L.obs_size + 31) / 32 + 1 : 0;
    L.off_lastg = HD_WORDS; L.off_lastcx = L.off_lastg + L.MW; L.off_state = L.off_lastcx + L.MW; L.off_extra = L.off_state + L.SW;
    L.W = L.off_extra + L.XW;
    L.PW = L.SW + L.XW + 1;
    int cap = cfg->solution_capacity;
    if (cap == 0) cap = cfg->max_depth + (L.kind == QG_ENV_PAULI_NETWORK ? L.Rtot : 0);
    if (!cfg->track_solution) cap = 0;
    L.sol_cap = std::min(std::max(cap, 1), 65535);
    if (L.SCR < 1) L.SCR = 1;
    return QG_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Gateset symmetries.  symmetry.rs enumerates coupling-graph automorphisms (VF2, then sort+dedup,
// 115-176) or, without any two-qubit gate, all permutations in Heap's order (84-113); an automorphism is
// kept when every permuted gate exists in the gateset (178-203).
// ---------------------------------------------------------------------------------------------------------
namespace {

typedef std::vector<int> IPerm;

void heap_order(int k, IPerm& p, std::vector<IPerm>& out) {
    if (k <= 1) { out.push_back(p); return; }
    heap_order(k - 1, p, out);
    for (int i = 0; i + 1 < k; ++i) {
        std::swap(p[(k & 1) ? 0 : i], p[k - 1]);
        heap_order(k - 1, p, out);
    }
}

struct AutoSearch {
    int n; const std::vector<std::vector<char>>& adj; std::vector<int> deg; IPerm img; std::vector<char> taken; std::vector<IPerm>& out;
    AutoSearch(const std::vector<std::vector<char>>& a, std::vector<IPerm>& o) : n((int)a.size()), adj(a), deg(a.size(), 0), img(a.size(), -1), taken(a.size(), 0), out(o) {
        for (int i = 0; i < n; ++i) deg[i] = (int)std::count(adj[i].begin(), adj[i].end(), (char)1);
    }
    void go(int v) {
        if (v == n) { out.push_back(img); return; }
        for (int cand = 0; cand < n; ++cand) {          // ascending candidates => lexicographic output order
            if (taken[cand] || deg[cand] != deg[v]) continue;
            bool fits = true;
            for (int u = 0; u < v; ++u) if (adj[v][u] != adj[cand][img[u]]) { fits = false; break; }
            if (!fits) continue;
            taken[cand] = 1; img[v] = cand;
            go(v + 1);
            taken[cand] = 0; img[v] = -1;
        }
    }
};

struct GKey {
    int kind, a, b;
    bool operator<(const GKey& o) const { return kind != o.kind ? kind < o.kind : (a != o.a ? a < o.a : b < o.b); }
};
GKey key_of(const qg_gate& g, const IPerm* p) {
    int a = g.q0, b = two_qubit(g.kind) ? g.q1 : -1;
    if (p) { a = (*p)[a]; if (b >= 0) b = (*p)[b]; }
    if (g.kind == QG_SWAP && b < a) std::swap(a, b);      // SWAP keys are order-insensitive (symmetry.rs:66-71)
    return GKey{g.kind, a, b};
}

}  // namespace

int compute_twists(const qg_config* cfg, bool internal_pauli, Twists& out) {
    out = Twists();
    const int rc = validate_config(cfg);
    if (rc != QG_OK) return rc;
    // PauliEnv::twists() is empty (pauli.rs:675-679); its qubit perms are internal.
    if (cfg->env_kind == QG_ENV_PAULI_NETWORK && !internal_pauli) return QG_OK;
    if (!cfg->add_perms) return QG_OK;
    const int n = cfg->num_qubits, A = cfg->num_gates;
    std::map<GKey, int> index;                            // later duplicates overwrite (symmetry.rs:217-223)
    for (int i = 0; i < A; ++i) index[key_of(cfg->gateset[i], nullptr)] = i;
    std::vector<std::vector<char>> adj(n, std::vector<char>(n, 0));
    bool any_edge = false;
    for (int i = 0; i < A; ++i) {
        const qg_gate& g = cfg->gateset[i];
        if (two_qubit(g.kind) && g.q0 != g.q1) { adj[g.q0][g.q1] = adj[g.q1][g.q0] = 1; any_edge = true; }
    }
    std::vector<IPerm> autos;
    if (!any_edge) { IPerm p(n); std::iota(p.begin(), p.end(), 0); heap_order(n, p, autos); }
    else {
        AutoSearch s(adj, autos); s.go(0);
        std::sort(autos.begin(), autos.end());
        autos.erase(std::unique(autos.begin(), autos.end()), autos.end());
    }
    auto act_perm_for = [&](const IPerm& p, std::vector<int64_t>& ap) {
        ap.clear();
        for (int i = 0; i < A; ++i) {
            auto it = index.find(key_of(cfg->gateset[i], &p));
            if (it == index.end()) return false;
            ap.push_back(it->second);
        }
        return true;
    };
    auto obs_perm_for = [&](const IPerm& p) {
        std::vector<int64_t> o;
        if (cfg->env_kind == QG_ENV_PAULI_NETWORK) { o.assign(p.begin(), p.end()); return o; }
        const int d = cfg->env_kind == QG_ENV_CLIFFORD ? 2 * n : n;
        auto lift = [&](int i) { return i < n ? p[i] : n + p[i - n]; };    // symmetry.rs:265-295
        o.resize((size_t)d * d);
        for (int r = 0; r < d; ++r) for (int c = 0; c < d; ++c) o[(size_t)r * d + c] = (int64_t)lift(r) * d + lift(c);
        return o;
    };
    std::vector<IPerm> seen;                              // Heap's order may not repeat, but mirror the reference's dedup
    for (const IPerm& p : autos) {
        std::vector<int64_t> ap;
        if (act_perm_for(p, ap)) { out.obs_perms.push_back(obs_perm_for(p)); out.act_perms.push_back(ap); }
    }
    if (out.obs_perms.empty()) {
        IPerm id(n); std::iota(id.begin(), id.end(), 0);
        std::vector<int64_t> ap;
        if (act_perm_for(id, ap)) { out.obs_perms.push_back(obs_perm_for(id)); out.act_perms.push_back(ap); }
    }
    return QG_OK;
}

// BFS all-pairs distances over the CX coupling graph, grouped by distance (pauli.rs:56-111).
void pauli_gen_tables(const qg_config* cfg, std::vector<uint32_t>& out) {
    const int n = cfg->num_qubits;
    std::vector<std::vector<int>> nb(n);
    std::vector<uint32_t> cx;
    for (int i = 0; i < cfg->num_gates; ++i) {
        const qg_gate& g = cfg->gateset[i];
        if (g.kind != QG_CX) continue;
        cx.push_back((uint32_t)g.q0 | ((uint32_t)g.q1 << 8));
        if (std::find(nb[g.q0].begin(), nb[g.q0].end(), g.q1) == nb[g.q0].end()) nb[g.q0].push_back(g.q1);
        if (std::find(nb[g.q1].begin(), nb[g.q1].end(), g.q0) == nb[g.q1].end()) nb[g.q1].push_back(g.q0);
    }
    std::map<int, std::vector<uint32_t>> by_dist;
    for (int src = 0; src < n; ++src) {
        std::vector<int> dist(n, -1), frontier{src};
        dist[src] = 0;
        for (size_t h = 0; h < frontier.size(); ++h) for (int v : nb[frontier[h]]) if (dist[v] < 0) { dist[v] = dist[frontier[h]] + 1; frontier.push_back(v); }
        for (int dst = src + 1; dst < n; ++dst) if (dist[dst] >= 0) by_dist[dist[dst]].push_back((uint32_t)src | ((uint32_t)dst << 8));
    }
    size_t np = 0; for (auto& kv : by_dist) np += kv.second.size();
    out.clear();
    out.push_back((uint32_t)by_dist.size()); out.push_back((uint32_t)np); out.push_back((uint32_t)cx.size());
    for (auto& kv : by_dist) out.push_back((uint32_t)kv.first);
    uint32_t off = 0;
    for (auto& kv : by_dist) { out.push_back(off); off += (uint32_t)kv.second.size(); }
    out.push_back(off);
    for (auto& kv : by_dist) out.insert(out.end(), kv.second.begin(), kv.second.end());
    out.insert(out.end(), cx.begin(), cx.end());
}

// ---------------------------------------------------------------------------------------------------------
// set_state packers
// ---------------------------------------------------------------------------------------------------------
static inline void put_bit(uint32_t* w, int64_t bit) { w[bit >> 5] |= 1u << (bit & 31); }

void pack_identity(const Layout& L, uint32_t* out) {
    std::memset(out, 0, sizeof(uint32_t) * (size_t)L.PW);
    if (L.kind == QG_ENV_PERMUTATION) { for (int q = 0; q < L.n; ++q) out[q >> 2] |= (uint32_t)q << ((q & 3) * 8); }
    else if (L.kind == QG_ENV_PAULI_NETWORK) { for (int r = 0; r < 2 * L.n; ++r) put_bit(out, (int64_t)r * L.CW + r); out[L.SW + PX_ORD0] = 0x76543210u; out[L.SW + PX_ORD1] = 0xFEDCBA98u; }
    else { for (int r = 0; r < L.D; ++r) put_bit(out, (int64_t)r * L.D + r); }
    out[L.PW - 1] = 1;   // solved
}

// Pauli label grammar ^[+-]?[ij1]?[IXYZ]*$ and symplectic encoding (pauli/pauli.rs:22-81).
static bool parse_label(const std::vector<int64_t>& chars, int n, uint64_t& x, uint64_t& z, int& base_phase) {
    size_t i = 0; int sign = 0; char unit = 0;
    if (i < chars.size() && (chars[i] == '+' || chars[i] == '-')) { sign = chars[i] == '-' ? -1 : 1; ++i; }
    if (i < chars.size() && (chars[i] == 'i' || chars[i] == 'j' || chars[i] == '1')) { unit = (char)chars[i]; ++i; }
    if ((int)(chars.size() - i) != n) return false;
    const bool imag = unit == 'i' || unit == 'j';
    const int phase = (sign < 0) ? (imag ? 1 : 2) : (imag ? 3 : 0);    // "", "-i", "-", "i" -> 0,1,2,3 (pauli.rs:29-37)
    x = z = 0; int ys = 0;
    for (int k = 0; k < n; ++k) {
        const int64_t ch = chars[chars.size() - 1 - (size_t)k];        // qubit k is the k-th char from the right (pauli.rs:62)
        if (ch == 'X') x |= 1ull << k; else if (ch == 'Z') z |= 1ull << k; else if (ch == 'Y') { x |= 1ull << k; z |= 1ull << k; ++ys; }
        else if (ch != 'I') return false;
    }
    base_phase = (phase + ys) % 4;
    return true;
}

int pack_state(const qg_config* cfg, const Layout& L, const int64_t* p, int64_t avail, uint32_t* out, int64_t* used) {
    std::memset(out, 0, sizeof(uint32_t) * (size_t)L.PW);
    const int n = L.n;
    if (L.kind == QG_ENV_PERMUTATION) {                   // permutation.rs:168-173
        if (avail < n) { set_error("set_state: permutation payload needs num_qubits entries"); return QG_ERR_STATE; }
        for (int q = 0; q < n; ++q) {
            if (p[q] < 0 || p[q] >= n) { set_error("set_state: permutation entry out of range (the reference would index out of bounds)"); return QG_ERR_STATE; }
            out[q >> 2] |= (uint32_t)p[q] << ((q & 3) * 8);
        }
        *used = n; return QG_OK;
    }
    if (L.kind != QG_ENV_PAULI_NETWORK) {                 // linear_function.rs:279-283, clifford.rs:299-304: x > 0 => 1
        const int64_t len = (int64_t)L.D * L.D;
        if (avail < len) { set_error("set_state: matrix payload needs dim*dim entries"); return QG_ERR_STATE; }
        for (int64_t i = 0; i < len; ++i) if (p[i] > 0) put_bit(out, i);
        *used = len; return QG_OK;
    }
    // PauliNetwork: [R, tableau(4n^2), len_0, chars_0..., len_1, ...] (pauli.rs:517-541)
    int64_t i = 0;
    auto next = [&](int64_t& v) { if (i < avail) { v = p[i++]; return true; } v = 0; return false; };
    int64_t v; next(v);
    const int64_t rot_count = std::max<int64_t>(v, 0);
    const int D = 2 * n;
    bool identity = true;
    for (int r = 0; r < D; ++r) for (int c = 0; c < D; ++c) {
        next(v);
        const bool bit = v > 0;
        if (bit) put_bit(out, (int64_t)r * L.CW + c);
        if (bit != (r == c)) identity = false;
    }
    std::vector<uint64_t> xs, zs; std::vector<int> ph;
    for (int64_t r = 0; r < rot_count; ++r) {
        next(v);
        const int64_t len = std::max<int64_t>(v, 0);
        std::vector<int64_t> chars;
        for (int64_t k = 0; k < len; ++k) { if (!next(v)) { set_error("malformed state: not enough characters for rotation string"); return QG_ERR_STATE; } chars.push_back(v); }
        if (r < L.max_rot) {                              // rotations beyond max_rotations are dropped (pauli.rs:537-539)
            uint64_t x, z; int bp;
            if (!parse_label(chars, n, x, z, bp)) { set_error("malformed state: invalid Pauli label (must match [+-]?[ij1]?[IXYZ]{num_qubits})"); return QG_ERR_STATE; }
            xs.push_back(x); zs.push_back(z); ph.push_back(bp);
        }
    }
    const int R = (int)xs.size();
    uint32_t plo = 0, phi = 0;
    uint32_t* X = out + L.SW;
    for (int r = 0; r < R; ++r) {
        for (int k = 0; k < n; ++k) {
            if ((xs[r] >> k) & 1) put_bit(out, (int64_t)k * L.CW + D + r);
            if ((zs[r] >> k) & 1) put_bit(out, (int64_t)(n + k) * L.CW + D + r);
        }
        plo |= (uint32_t)(ph[r] & 1) << r; phi |= (uint32_t)((ph[r] >> 1) & 1) << r;
        uint32_t anti = 0;                                // edges r -> j for j < r that anticommute (pauli_dag.rs:36-42)
        for (int j = 0; j < r; ++j) {
            const int par = __builtin_popcountll(xs[r] & zs[j]) + __builtin_popcountll(zs[r] & xs[j]);
            if (par & 1) anti |= 1u << j;
        }
        X[PX_ANTI + (r >> 1)] |= anti << ((r & 1) * 16);
    }
    X[PX_PLO] = plo; X[PX_PHI] = phi; X[PX_ALIVE] = R >= 32 ? 0xFFFFFFFFu : ((1u << R) - 1u);
    X[PX_ORD0] = 0x76543210u; X[PX_ORD1] = 0xFEDCBA98u;
    X[PX_MISC] = ((uint32_t)R << 16) | ((uint32_t)R << 24);
    out[L.PW - 1] = (R == 0 && identity) ? 1u : 0u;       // PauliNetwork::solved (pauli_network.rs:167-173)
    *used = i; return QG_OK;
}

bool is_symplectic(const Layout& L, const uint32_t* packed) {
    const int n = L.n, D = 2 * n;
    auto bit = [&](int r, int c) { const int64_t b = (int64_t)r * D + c; return (packed[b >> 5] >> (b & 31)) & 1u; };
    for (int i = 0; i < D; ++i)
        for (int j = i; j < D; ++j) {
            uint32_t s = 0;                      // sum_k M[i][k] M[j][k +- n]
            for (int k = 0; k < n; ++k) s ^= (bit(i, k) & bit(j, k + n)) ^ (bit(i, k + n) & bit(j, k));
            if (s != (uint32_t)(j == i + n ? 1 : 0)) return false;
        }
    return true;
}

int unpack_state(const Layout& L, const uint32_t* rec, uint8_t* out, int64_t cap, int64_t* len) {
    const uint32_t* S = rec + L.off_state;
    auto bit = [&](int64_t b) { return (uint8_t)((S[b >> 5] >> (b & 31)) & 1u); };
    int64_t need = 0;
    if (L.kind == QG_ENV_PERMUTATION) {
        need = L.n; if (cap < need) { set_error("get_state: buffer too small"); return QG_ERR_INVALID; }
        for (int q = 0; q < L.n; ++q) out[q] = (uint8_t)((S[q >> 2] >> ((q & 3) * 8)) & 0xFFu);
    } else if (L.kind != QG_ENV_PAULI_NETWORK) {
        need = (int64_t)L.D * L.D; if (cap < need) { set_error("get_state: buffer too small"); return QG_ERR_INVALID; }
        for (int64_t i = 0; i < need; ++i) out[i] = bit(i);
    } else {
        const uint32_t* X = rec + L.off_extra;
        const int R = (int)((X[PX_MISC] >> 16) & 0xFFu), D = 2 * L.n, cols = D + R;
        need = (int64_t)D * cols; if (cap < need) { set_error("get_state: buffer too small"); return QG_ERR_INVALID; }
        for (int r = 0; r < D; ++r) {
            for (int c = 0; c < D; ++c) out[(int64_t)r * cols + c] = bit((int64_t)r * L.CW + c);
            for (int k = 0; k < R; ++k) out[(int64_t)r * cols + D + k] = ((X[PX_ALIVE] >> k) & 1u) ? bit((int64_t)r * L.CW + D + k) : 0;
        }
    }
    *len = need;
    return QG_OK;
}

}  // namespace qg
