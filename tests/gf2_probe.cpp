// Host build of qiskit_gym_b200/csrc/qg_gf2.cuh for tests/test_host.py (g++, no CUDA): the same template code the step kernel
// runs for the add_inverts coin, applied to one dense row-major bit stream.
#include <stdint.h>

#include "../qiskit_gym_b200/csrc/qg_gf2.cuh"

namespace {
struct Words {
    uint32_t* p;
    uint32_t& operator[](int i) const { return p[i]; }
};
}  // namespace

extern "C" {
// returns 1 on success, 0 if singular, -1 for an unsupported bucket
int probe_gauss_jordan(int dmax, int D, uint32_t* words) {
    Words S{words};
    switch (dmax) {
        case 8: return qg::gf2_invert_rows<8>(S, D) ? 1 : 0;
        case 16: return qg::gf2_invert_rows<16>(S, D) ? 1 : 0;
        case 32: return qg::gf2_invert_rows<32>(S, D) ? 1 : 0;
    }
    return -1;
}
int probe_symplectic(int dmax, int n, uint32_t* words) {
    Words S{words};
    switch (dmax) {
        case 8: qg::symplectic_invert_rows<8>(S, n); return 1;
        case 16: qg::symplectic_invert_rows<16>(S, n); return 1;
        case 32: qg::symplectic_invert_rows<32>(S, n); return 1;
    }
    return -1;
}
}
