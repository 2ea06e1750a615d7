// 4-D table-free queries on the component-interleaved grid [nt][nz][ny][nx][4] (Bx, By, Bz, |B| or 0): the per-lane
// math of query_gridil4_kernel (arb_gridil4.cu), __host__ __device__ so that tests/host_emul/gridil4_host_emul.cu
// runs it on the CPU against alpha = A f (A.py:726-878).
//
// Four lanes own a query, lane k = z-plane k of the 4^4 neighbourhood, and the four t-planes l = 0..3 are passes:
// in pass l a lane holds the 4 x 4 (y, x) points of plane (k, l) with all components together (one 512-byte slot,
// four 128-byte rows of the grid) and contracts them with the Catmull-Rom weights (A = M^(x)4 apart from the
// A.py:860 term).  The A.py:860 term is folded into a second, rank-one-per-corner set of weights on
// the same points (quirk_weights below), so a pass adds it from the plane's four signed xy parity sums and nothing is
// carried between passes or exchanged between lanes; the four lanes' shares simply add up.
#pragma once
#include "arb_gridfree.cuh"

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

struct Weights {          // Catmull-Rom weights of the query's x and y cell fractions, computed once per query
    double wx[4], dwx[4], wy[4], dwy[4];
};

ARB_HD void make_weights(const double* f, Weights& W) {
    gridfree::catmull_rom(f[0], W.wx, W.dwx);
    gridfree::catmull_rom(f[1], W.wy, W.dwy);
}

// The A.py:860 term as point weights.  The reference's matrix adds  sum_c Phi_c (F(c-1) - F(c))  to the M^(x)4 value,
// c = cx + 2 cy + 4 cz + 8 ct the reference's corner order, F = fxyzt at the corner, F(-1) = 0, Phi_c = the product of the
// four Hermite slope basis functions of corner c (arb_gridfree.cuh).  Summed by parts that is  sum_c F(c) G(c),
// G(c) = Phi_(c+1) - Phi_c (Phi_16 = 0), and F(c) is the 1/16 (+-) sum over the 16 neighbourhood points whose index parity
// per axis equals c's bits -- every point belongs to exactly one corner -- so the term is a second set of weights on the
// same points:  sum_(i,j,k,l) f[l][k][j][i] s_i s_j s_k s_l G(i&1, j&1, k&1, l&1) / 16,  s = +1 for index >= 2, else -1.
// For plane (k, l) that is the plane's four signed xy parity sums (gridfree::plane_il4) times the four numbers below;
// nothing is carried between passes or lanes.  Derivatives: G is multilinear in the four slope vectors, so d/du replaces
// hx by its derivative, and so on.
// g[cx + 2 cy] for cz = k & 1, ct = l & 1, including the factor s_k s_l / 16.  X0, X1 / Y0, Y1 / Z, ZN / T, TN are the slope
// basis values (or derivatives) of x, y at the two corners and of z, t at this corner layer and at the next one in the order
// (cz, ct) = (0,0), (1,0), (0,1), (1,1) (ZN * TN = 0 after the last).
ARB_HD void quirk_weights(double X0, double X1, double Y0, double Y1, double ZT, double ZTN, double sgn, double (&g)[4]) {
    const double dx = X1 - X0;
    g[0] = sgn * (dx * Y0 * ZT);                                  // Phi(1,0) - Phi(0,0)
    g[1] = sgn * ((X0 * Y1 - X1 * Y0) * ZT);                      // Phi(0,1) - Phi(1,0)
    g[2] = sgn * (dx * Y1 * ZT);                                  // Phi(1,1) - Phi(0,1)
    g[3] = sgn * fma_(X0 * Y0, ZTN, -(X1 * Y1) * ZT);             // Phi(0,0 of the next layer) - Phi(1,1)
}

ARB_HD double quirk_dot_a(double X0, double X1, double Y0, double Y1, double P0, double P1, double P2, double P3) {
    return fma_(X1 - X0, fma_(Y1, P2, Y0 * P0), fma_(fma_(X0, Y1, -(X1 * Y0)), P1, -(X1 * Y1) * P3));
}

// pass l of lane k: slot = [j][i][c], plane (z = k, t = l) of the neighbourhood; v = this lane's share of: components
// 0..2, |B| 3, d|B|/du, dv, dw 4..6, d|B|/ds 7
template <bool BOTH, bool QUIRK>
ARB_HD void pass(double (&v)[8], const double* slot, int k, int l, const double* frac, const Weights& W) {
    constexpr int NC = BOTH ? 4 : 3;
    double wz[4], dwz[4], wt[4], dwt[4];
    gridfree::catmull_rom(frac[2], wz, dwz);
    gridfree::catmull_rom(frac[3], wt, dwt);
    double out[7], Pq[4][4];
    gridfree::plane_il4<BOTH, QUIRK>(slot, sel4(wz, k), sel4(dwz, k), W.wx, W.dwx, W.wy, W.dwy, out, Pq);
    const double w = sel4(wt, l), dw = sel4(dwt, l);
    ARB_UNROLL
    for (int c = 0; c < 3; ++c) v[c] = fma_(w, out[c], v[c]);
    if (BOTH) {
        ARB_UNROLL
        for (int c = 3; c < 7; ++c) v[c] = fma_(w, out[c], v[c]);
        v[7] = fma_(dw, out[3], v[7]);
    }
    if (QUIRK) {
        double hx[2], dhx[2], hy[2], dhy[2], hz[2], dhz[2], ht[2], dht[2];
        gridfree::hermite_slope(frac[0], hx, dhx);
        gridfree::hermite_slope(frac[1], hy, dhy);
        gridfree::hermite_slope(frac[2], hz, dhz);
        gridfree::hermite_slope(frac[3], ht, dht);
        const int cz = k & 1, ct = l & 1;
        const bool last = cz && ct;                    // corner layer (1, 1): nothing follows
        const int czn = cz ^ 1, ctn = cz ? 1 : ct;     // (0,0) -> (1,0) -> (0,1) -> (1,1)
        const double sgn = (((k >> 1) == (l >> 1)) ? 0.0625 : -0.0625);
        const double z = cz ? hz[1] : hz[0], t = ct ? ht[1] : ht[0];
        const double zn = last ? 0.0 : (czn ? hz[1] : hz[0]), tn = ctn ? ht[1] : ht[0];
        double g[4];
        quirk_weights(hx[0], hx[1], hy[0], hy[1], z * t, zn * tn, sgn, g);
        ARB_UNROLL
        for (int c = 0; c < NC; ++c)
            v[c] = fma_(Pq[3][c], g[3], fma_(Pq[2][c], g[2], fma_(Pq[1][c], g[1], fma_(Pq[0][c], g[0], v[c]))));
        if (BOTH) {
            // |B| gradient: the dot of the parity sums with quirk_weights(X, Y, ZT, ZTN) is ZT * A(X, Y) + ZTN * B(X, Y),
            // A = (X1 - X0)(Y0 P0 + Y1 P2) + (X0 Y1 - X1 Y0) P1 - X1 Y1 P3,  B = X0 Y0 P3; a derivative replaces one vector
            const double dz = cz ? dhz[1] : dhz[0], dt = ct ? dht[1] : dht[0];
            const double dzn = last ? 0.0 : (czn ? dhz[1] : dhz[0]), dtn = ctn ? dht[1] : dht[0];
            const double P0 = Pq[0][3], P1 = Pq[1][3], P2 = Pq[2][3], P3 = Pq[3][3];
            const double A = quirk_dot_a(hx[0], hx[1], hy[0], hy[1], P0, P1, P2, P3), B = hx[0] * hy[0] * P3;
            const double Au = quirk_dot_a(dhx[0], dhx[1], hy[0], hy[1], P0, P1, P2, P3), Bu = dhx[0] * hy[0] * P3;
            const double Av = quirk_dot_a(hx[0], hx[1], dhy[0], dhy[1], P0, P1, P2, P3), Bv = hx[0] * dhy[0] * P3;
            const double zt = sgn * (z * t), ztn = sgn * (zn * tn);
            v[4] = fma_(ztn, Bu, fma_(zt, Au, v[4]));
            v[5] = fma_(ztn, Bv, fma_(zt, Av, v[5]));
            v[6] = fma_(sgn * (dzn * tn), B, fma_(sgn * (dz * t), A, v[6]));
            v[7] = fma_(sgn * (zn * dtn), B, fma_(sgn * (z * dt), A, v[7]));
        }
    }
}

}  // namespace gridil4
}  // namespace arb
