// qg_host.hpp — host-side pieces of the engine that need no GPU: config validation, layout sizing,
// gateset symmetries (twists) and the set_state payload packers.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/qg_engine.h"

namespace qg {

void set_error(const std::string& msg);
const char* get_error();

// Derived sizes of one configuration (all word counts are 32-bit words per environment).
struct Layout {
    int kind = 0, n = 0, D = 0, A = 0;
    int obs_rows = 0, obs_cols = 0, obs_size = 0;
    int SW = 0, MW = 0, XW = 0, W = 0;       // state / metrics (each array) / PauliNetwork extras / record
    int off_lastg = 0, off_lastcx = 0, off_state = 0, off_extra = 0;
    int OW = 0;                               // observation bit-stream words (PauliNetwork)
    int SCR = 0;                              // scratch words per env in shared memory
    int Rtot = 0, CW = 0, max_rot = 0;
    int sol_cap = 0;
    int PW = 0;                               // staged payload words (set_state)
    int64_t state_len = 0;                    // set_state payload length in i64 (0 = variable)
};

int validate_config(const qg_config* cfg);              // QG_OK or error (message set)
int make_layout(const qg_config* cfg, Layout& L);       // validates, then sizes

// twists (symmetry.rs:205-361)
struct Twists {
    std::vector<std::vector<int64_t>> obs_perms, act_perms;
};
int compute_twists(const qg_config* cfg, bool internal_pauli, Twists& out);

// PauliNetwork reset generator tables: coupling-graph distance classes (pauli.rs:56-111) and CX pairs (pauli.rs:357-364)
void pauli_gen_tables(const qg_config* cfg, std::vector<uint32_t>& out);

// set_state payload -> staged words.  Returns QG_OK / QG_ERR_STATE.  `used` = i64 entries consumed.
int pack_state(const qg_config* cfg, const Layout& L, const int64_t* payload, int64_t avail, uint32_t* out, int64_t* used);
// identity payload words (constructor state)
void pack_identity(const Layout& L, uint32_t* out);
// Clifford: is the packed 2n x 2n matrix symplectic (M J M^T = J, J = [[0,I],[I,0]])?  Every state the gates can reach from the
// identity is; a set_state payload need not be, and then the coin's inverse must not use the transpose shortcut (qg_gf2.cuh).
bool is_symplectic(const Layout& L, const uint32_t* packed);
// record column (W words) -> reference byte-per-entry layout
int unpack_state(const Layout& L, const uint32_t* rec, uint8_t* out, int64_t cap, int64_t* len);

}  // namespace qg
