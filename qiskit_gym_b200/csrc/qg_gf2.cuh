// qg_gf2.cuh — register-resident GF(2) matrix inversion for the `add_inverts` coin of LinearFunction / Clifford
// (reference: LFState::inverse linear_function.rs:124-146, CFState::inverse clifford.rs:147-170 — Gauss-Jordan on a
// byte-per-bit matrix pair).  The inverse of a matrix is unique, so any exact method gives the reference's result:
//   * gf2_invert_rows<DMAX>      Gauss-Jordan with every row held in one register (D <= DMAX <= 32), branch free:
//                                the pivot is fixed by adding the first later row that has the bit, elimination is a
//                                masked XOR.  All row indices are compile-time constants, so nothing spills.
//   * symplectic_invert_rows<DMAX>  for symplectic matrices (every Clifford state reached from the identity by gates, and
//                                every target produced by CliffordGym.get_state): M^-1 = J M^T J with J = [[0,I],[I,0]]
//                                — a bit-matrix transpose (log2(DMAX) butterfly stages) instead of an elimination.
// The functions are written against an accessor `Wd` (operator[](int) -> uint32_t&) over the dense row-major bit stream and
// use only portable integer code, so the same header is compiled for the host by tests/test_host.py.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define QG_HD __host__ __device__ __forceinline__
#else
#define QG_HD inline
#endif

namespace qg {

// bits [o, o+len) of the stream, len in 1..32
template <class Wd>
QG_HD uint32_t stream_get(const Wd& W, int o, int len) {
    const int w = o >> 5, s = o & 31;
    const uint32_t lo = W[w];
    const uint32_t hi = (s + len > 32) ? W[w + 1] : 0u;
    const uint32_t v = (uint32_t)(((((uint64_t)hi) << 32) | lo) >> s);
    return len >= 32 ? v : (v & ((1u << len) - 1u));
}
// stream[o, o+len) = v   (v has no bits above len)
template <class Wd>
QG_HD void stream_set(const Wd& W, int o, int len, uint32_t v) {
    const int w = o >> 5, s = o & 31;
    const uint32_t m = len >= 32 ? 0xFFFFFFFFu : ((1u << len) - 1u);
    W[w] = (W[w] & ~(m << s)) | (v << s);
    if (s + len > 32) W[w + 1] = (W[w + 1] & ~(m >> (32 - s))) | (v >> (32 - s));
}

// In-place transpose of a DMAX x DMAX bit matrix, row i in r[i], element (i, j) = bit j of r[i].
template <int DMAX>
QG_HD void transpose_bits(uint32_t (&r)[DMAX]) {
#pragma unroll
    for (int j = DMAX / 2; j >= 1; j >>= 1) {
        // mask: bit b set iff (b & j) == 0, within the low DMAX bits
        uint32_t m = 0;
#pragma unroll
        for (int b = 0; b < DMAX; ++b) if ((b & j) == 0) m |= 1u << b;
#pragma unroll
        for (int k = 0; k < DMAX; ++k) {
            if ((k & j) == 0) {
                // swap the (rows k.., cols with bit j) block with the (rows k+j.., cols without bit j) block
                const uint32_t t = ((r[k] >> j) ^ r[k + j]) & m;
                r[k + j] ^= t;
                r[k] ^= t << j;
            }
        }
    }
}

// S <- (J S J)^T = S^-1 for a symplectic D x D matrix, D = 2n <= DMAX.
template <int DMAX, class Wd>
QG_HD void symplectic_invert_rows(const Wd& S, int n) {
    const int D = 2 * n;
    const uint32_t half = (1u << n) - 1u;
    uint32_t r[DMAX];
#pragma unroll
    for (int i = 0; i < DMAX; ++i) {
        uint32_t v = 0;
        if (i < D) {
            const int src = i < n ? i + n : i - n;            // J on the left: row blocks swapped
            const uint32_t w = stream_get(S, src * D, D);
            v = (w >> n) | ((w & half) << n);                  // J on the right: column blocks swapped
        }
        r[i] = v;
    }
    transpose_bits<DMAX>(r);
#pragma unroll
    for (int i = 0; i < DMAX; ++i) if (i < D) stream_set(S, i * D, D, r[i]);
}

// Gauss-Jordan, D <= DMAX <= 16: row i of [M | I] in one word (M in bits [0,16), the accumulating inverse in [16,32)).
// Returns false (S untouched) if the matrix is singular.
template <int DMAX, class Wd>
QG_HD bool gf2_invert_rows16(const Wd& S, int D) {
    static_assert(DMAX <= 16, "two halves share a 32-bit word");
    uint32_t a[DMAX];
#pragma unroll
    for (int i = 0; i < DMAX; ++i) a[i] = i < D ? (stream_get(S, i * D, D) | (1u << (16 + i))) : 0u;
    uint32_t bad = 0;
#pragma unroll
    for (int col = 0; col < DMAX; ++col) {
        if (col < D) {
            uint32_t need = ~(a[col] >> col) & 1u;
#pragma unroll
            for (int i = col + 1; i < DMAX; ++i) {
                const uint32_t take = need & (a[i] >> col);   // bit 0 says: row i is the first later row with the bit
                a[col] ^= a[i] & (0u - (take & 1u));
                need &= ~take;
            }
            bad |= need & 1u;
            const uint32_t piv = a[col];
#pragma unroll
            for (int i = 0; i < DMAX; ++i)
                if (i != col) a[i] ^= piv & (0u - ((a[i] >> col) & 1u));
        }
    }
    if (bad) return false;
#pragma unroll
    for (int i = 0; i < DMAX; ++i) if (i < D) stream_set(S, i * D, D, a[i] >> 16);
    return true;
}

// Gauss-Jordan, D <= 32: M rows and inverse rows in separate registers.
template <class Wd>
QG_HD bool gf2_invert_rows32(const Wd& S, int D) {
    uint32_t m[32], v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) { m[i] = i < D ? stream_get(S, i * D, D) : 0u; v[i] = 1u << i; }
    uint32_t bad = 0;
#pragma unroll
    for (int col = 0; col < 32; ++col) {
        if (col < D) {
            uint32_t need = ~(m[col] >> col) & 1u;
#pragma unroll
            for (int i = col + 1; i < 32; ++i) {
                const uint32_t take = need & (m[i] >> col);
                const uint32_t msk = 0u - (take & 1u);
                m[col] ^= m[i] & msk; v[col] ^= v[i] & msk;
                need &= ~take;
            }
            bad |= need & 1u;
            const uint32_t pm = m[col], pv = v[col];
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (i != col) { const uint32_t msk = 0u - ((m[i] >> col) & 1u); m[i] ^= pm & msk; v[i] ^= pv & msk; }
        }
    }
    if (bad) return false;
#pragma unroll
    for (int i = 0; i < 32; ++i) if (i < D) stream_set(S, i * D, D, v[i]);
    return true;
}

template <int DMAX, class Wd>
QG_HD bool gf2_invert_rows(const Wd& S, int D) {
    if constexpr (DMAX <= 16) return gf2_invert_rows16<DMAX>(S, D);
    else return gf2_invert_rows32(S, D);
}

}  // namespace qg
