// 4-D table-free queries on the component-interleaved grid [nt][nz][ny][nx][4] (Bx, By, Bz, |B| or 0): the per-lane
// math of query_gridil4_kernel (arb_gridil4.cu), __host__ __device__ so that tests/host_emul/gridil4_host_emul.cu
// runs it on the CPU against alpha = A f (A.py:726-878).
//
// Four lanes own a query, lane k = z-plane k of the 4^4 neighbourhood, and the four t-planes l = 0..3 are passes:
// in pass l a lane holds the 4 x 4 (y, x) points of plane (k, l) with all components together (one 512-byte slot,
// four 128-byte rows of the grid) and contracts them with the Catmull-Rom weights (A = M^(x)4 apart from the
// A.py:860 term).  The quirk term needs fxyzt at the cell's 16 corners: every pass adds the plane's
// signed xy parity sums into T[ct] with the t sign of that plane (l = ct: -, l = ct + 2: +); after the last pass
// lane k and lane k ^ 2 hold the two z-planes of corner layer cz = k & 1, so one exchange gives
// fxyzt(cx, cy, cz, ct) = (T_{z = cz + 2} - T_{z = cz}) / 16, and e[c] = fxyzt(c - 1) - fxyzt(c) in the reference's corner
// order c = cx + 2 cy + 4 cz + 8 ct takes one more value from lane k ^ 1 (nodes::quirk4_lane does the rest).
#pragma once
#include "arb_gridfree.cuh"
#include "arb_nodes.cuh"

namespace arb {
namespace gridil4 {

using gridfree::fma_;
using gridfree::sel4;

// Gather addressing (shared with the host emulation).  A slot is the 4 x 4 (y, x) points of one (z, t) plane with their
// four components: 4 rows of 128 contiguous bytes, nx grid points (32 B each) apart.  One warp-wide 16-byte copy fills
// it: lane i brings bytes [16 i, 16 i + 16) of the slot, i.e. row i / 8, piece i % 8 of that row.
ARB_HD int64_t lane_piece_bytes(int lane, int64_t nx) { return (int64_t)(lane >> 3) * nx * 32 + (lane & 7) * 16; }
// first grid point (count of 32-byte points from the grid's first byte) of the plane lane k (z-plane) needs in pass l
ARB_HD int64_t plane_first_point(const int* idx, int k, int l, int64_t nx, int64_t ny, int64_t nz) {
    return ((((int64_t)idx[3] + l) * nz + idx[2] + k) * ny + idx[1]) * nx + idx[0];
}

struct Weights {          // Catmull-Rom weights of the query's cell fractions, computed once per query
    double wx[4], dwx[4], wy[4], dwy[4], wz[4], dwz[4], wt[4], dwt[4];
};

ARB_HD void make_weights(const double* f, Weights& W) {
    gridfree::catmull_rom(f[0], W.wx, W.dwx);
    gridfree::catmull_rom(f[1], W.wy, W.dwy);
    gridfree::catmull_rom(f[2], W.wz, W.dwz);
    gridfree::catmull_rom(f[3], W.wt, W.dwt);
}

struct Acc {
    double v[8];          // this lane's share of: components 0..2, |B| 3, d|B|/du, dv, dw 4..6, d|B|/ds 7
    double T[2][4][4];    // quirk: [ct][cx + 2 cy][component], parity sums over x, y and t of this lane's z-plane
};

template <bool QUIRK>
ARB_HD void clear(Acc& A) {
    ARB_UNROLL
    for (int i = 0; i < 8; ++i) A.v[i] = 0.0;
    if (QUIRK) {
        ARB_UNROLL
        for (int ct = 0; ct < 2; ++ct)
            ARB_UNROLL
            for (int q = 0; q < 4; ++q)
                ARB_UNROLL
                for (int c = 0; c < 4; ++c) A.T[ct][q][c] = 0.0;
    }
}

// pass l of lane k: slot = [j][i][c], plane (z = k, t = l) of the neighbourhood
template <bool BOTH, bool QUIRK>
ARB_HD void pass(Acc& A, const double* slot, int k, int l, const Weights& W) {
    constexpr int NC = BOTH ? 4 : 3;
    double out[7], Pq[4][4];
    gridfree::plane_il4<BOTH, QUIRK>(slot, sel4(W.wz, k), sel4(W.dwz, k), W.wx, W.dwx, W.wy, W.dwy, out, Pq);
    const double w = sel4(W.wt, l), dw = sel4(W.dwt, l);
    ARB_UNROLL
    for (int c = 0; c < 3; ++c) A.v[c] = fma_(w, out[c], A.v[c]);
    if (BOTH) {
        ARB_UNROLL
        for (int c = 3; c < 7; ++c) A.v[c] = fma_(w, out[c], A.v[c]);
        A.v[7] = fma_(dw, out[3], A.v[7]);
    }
    if (QUIRK) {
        const double s0 = (l == 0) ? -1.0 : ((l == 2) ? 1.0 : 0.0), s1 = (l == 1) ? -1.0 : ((l == 3) ? 1.0 : 0.0);
        ARB_UNROLL
        for (int q = 0; q < 4; ++q)
            ARB_UNROLL
            for (int c = 0; c < NC; ++c) {
                A.T[0][q][c] = fma_(s0, Pq[q][c], A.T[0][q][c]);
                A.T[1][q][c] = fma_(s1, Pq[q][c], A.T[1][q][c]);
            }
    }
}

// fxyzt at the corners (cx, cy, cz = k & 1, ct) from this lane's T and the T of lane k ^ 2
ARB_HD double corner(double own, double other, int k) {
    return 0.0625 * ((k >= 2) ? (own - other) : (other - own));
}

// A.py:860 term of lane k < 2 (corner layer cz = k): F[ct][cx + 2 cy][c] = fxyzt at its corners, F11p[ct][c] = fxyzt of
// corner (1, 1) of lane k ^ 1 (the corner before this lane's (0, 0) in the reference's order, see the header)
template <bool BOTH>
ARB_HD void quirk(Acc& A, const double (&F)[2][4][4], const double (&F11p)[2][4], int k, const double* frac) {
    constexpr int NC = BOTH ? 4 : 3;
    const int cz = k & 1;
    ARB_UNROLL
    for (int ct = 0; ct < 2; ++ct)
        ARB_UNROLL
        for (int c = 0; c < NC; ++c) {
            const double f15[4] = {F[ct][0][c], F[ct][1][c], F[ct][2][c], F[ct][3][c]};
            const double prev = (ct == 0) ? (cz ? F11p[0][c] : 0.0) : (cz ? F11p[1][c] : F11p[0][c]);
            if (BOTH && c == 3) {
                double g[5] = {A.v[3], A.v[4], A.v[5], A.v[6], A.v[7]};
                nodes::quirk4_lane<true>(f15, prev, cz, ct, frac, g);
                A.v[3] = g[0]; A.v[4] = g[1]; A.v[5] = g[2]; A.v[6] = g[3]; A.v[7] = g[4];
            } else {
                double g[1] = {A.v[c]};
                nodes::quirk4_lane<false>(f15, prev, cz, ct, frac, g);
                A.v[c] = g[0];
            }
        }
}

}  // namespace gridil4
}  // namespace arb
