// Node (Hermite) table: instead of the 4^d monomial coefficients of every CELL (512 B / 2 KB per cell and
// component) store the 2^d derivative values of every grid NODE -- f, fx, fy, fxy, ... as central differences in
// unit-cell coordinates, i.e. exactly the rows of the reference's D matrix (A.py:129-173, 762-876) -- 64 B / 128 B
// per node and component: 4x (3-D, nodes stored as aligned x-pairs) / 16x (4-D) less memory, and neighbouring cells
// share their nodes in L2.
//
// The Lekien-Marsden polynomial of a cell is the tensor-product cubic Hermite interpolant of the b-vector
// (that is what alpha = inv(B) b says, A.py:112-125,175), so a query can be evaluated straight from the 2^d corner
// nodes with the Hermite basis
//     h00 = 2t^3 - 3t^2 + 1, h01 = -2t^3 + 3t^2 (values at 0 / 1), h10 = t^3 - 2t^2 + t, h11 = t^3 - t^2 (slopes):
//     p(u, v, w[, s]) = sum over corners c and derivative types tau of  prod_a h_{tau_a, c_a}(x_a) * N[c][tau].
// Same bytes per query as a cell block (2^d corners x 2^d values = 4^d doubles), gathered as 2^(d-1) x-pairs of nodes.
//
// 4-D quirk (A.py:860): the reference's b-vector holds, in the fxyzt slot of corner c, fxyzt of corner c-1 (zero for
// corner 0).  The nodes store the true fxyzt; the evaluation adds  sum_c Phi_c (fxyzt(c-1) - fxyzt(c)),
// Phi_c = h1cx(u) h1cy(v) h1cz(w) h1ct(s) -- the same rank-16 term as arb_build_sep.cuh / arb_gridfree.cuh.
//
// Layout: n' = n - 2 nodes per axis (grid points 1..n-2, node i = grid point i+1, so cell (ix, iy, ..) has corners at
// nodes (ix + cx, iy + cy, ..)), T = 2^d values per node, tau = tx + 2 ty + 4 tz (+ 8 tt).
//   4-D: [C][nt'][nz'][ny'][nx'][16]          (128 B per node: an x-pair is two whole 128-byte lines wherever it starts)
//   3-D: [C][nz'][ny'][nx' - 1 x-pairs][2][8] (64 B per node: pair ix = nodes ix, ix + 1 stored together, 128 B aligned --
//        every node twice, so that no x-pair straddles two lines; 4x smaller than the cell table instead of 8x)
// A query's slot in shared memory is  [cz][cy][cx][tau]  (3-D, 64 doubles) or, per lane (cz, ct) of a 4-D query,
// [cy][cx][tau] (64 doubles).  Everything here is __host__ __device__ so tests/host_emul/nodes_host_emul.cu runs it
// on the CPU against alpha = A f.
#pragma once
#include <stdint.h>
#include "arb_gridfree.cuh"   // ARB_HD, ARB_UNROLL, fma_, Pair2

namespace arb {
namespace nodes {

using gridfree::fma_;
using gridfree::Pair2;

struct Basis {            // index [corner]
    double v[2], s[2];    // value basis h00, h01; slope basis h10, h11
    double dv[2], ds[2];  // their derivatives
};

ARB_HD Basis hermite(double t) {
    Basis b;
    const double t2 = t * t, omt = 1.0 - t;
    b.v[1] = t2 * fma_(-2.0, t, 3.0);
    b.v[0] = 1.0 - b.v[1];
    b.s[0] = t * omt * omt;
    b.s[1] = t2 * (t - 1.0);
    b.dv[1] = 6.0 * t * omt;
    b.dv[0] = -b.dv[1];
    b.ds[0] = fma_(t, fma_(3.0, t, -4.0), 1.0);
    b.ds[1] = t * fma_(3.0, t, -2.0);
    return b;
}

// ---------------------------------------------------------------------------------------------------
// build: the 2^d central-difference values of one node from its 3^d neighbourhood.
// get(dx, dy, dz, dt) returns the grid value at offset (-1..1 per axis) from the node.
// ---------------------------------------------------------------------------------------------------
template <int D, typename Get>
ARB_HD void node_stencil(Get get, double* out) {
    // successive differencing: after axis a the array holds, for every offset of the remaining axes, the 2^(a+1)
    // (centre | half difference) combinations of the axes done so far
    constexpr int NT = (D == 4) ? 3 : 1;
    double a0[NT * 9][2];               // after x: [dt][dz][dy][tx]
    ARB_UNROLL
    for (int t = 0; t < NT; ++t)
        ARB_UNROLL
        for (int z = 0; z < 3; ++z)
            ARB_UNROLL
            for (int y = 0; y < 3; ++y) {
                const int tt = (D == 4) ? t - 1 : 0;
                const double m = get(-1, y - 1, z - 1, tt), c = get(0, y - 1, z - 1, tt), p = get(1, y - 1, z - 1, tt);
                a0[(t * 3 + z) * 3 + y][0] = c;
                a0[(t * 3 + z) * 3 + y][1] = 0.5 * (p - m);
            }
    double a1[NT * 3][4];               // after y: [dt][dz][tx + 2 ty]
    ARB_UNROLL
    for (int tz = 0; tz < NT * 3; ++tz)
        ARB_UNROLL
        for (int k = 0; k < 2; ++k) {
            a1[tz][k] = a0[tz * 3 + 1][k];
            a1[tz][2 + k] = 0.5 * (a0[tz * 3 + 2][k] - a0[tz * 3 + 0][k]);
        }
    double a2[NT][8];                   // after z: [dt][tx + 2 ty + 4 tz]
    ARB_UNROLL
    for (int t = 0; t < NT; ++t)
        ARB_UNROLL
        for (int k = 0; k < 4; ++k) {
            a2[t][k] = a1[t * 3 + 1][k];
            a2[t][4 + k] = 0.5 * (a1[t * 3 + 2][k] - a1[t * 3 + 0][k]);
        }
    if (D == 3) {
        ARB_UNROLL
        for (int k = 0; k < 8; ++k) out[k] = a2[0][k];
    } else {
        ARB_UNROLL
        for (int k = 0; k < 8; ++k) {
            out[k] = a2[NT / 2][k];
            out[8 + k] = 0.5 * (a2[NT - 1][k] - a2[0][k]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// x contraction of one row (two x-adjacent nodes, NP = 2^(d-1) (tx=0, tx=1) pairs each):
//   X[m] = sum_cx  N[cx][2m] hv[cx] + N[cx][2m+1] hs[cx]
// ---------------------------------------------------------------------------------------------------
template <int NP, bool GRAD>
ARB_HD void row_x(const double* row, const Basis& bx, double* X, double* Xu) {
    ARB_UNROLL
    for (int m = 0; m < NP; ++m) {
        const Pair2 a = *reinterpret_cast<const Pair2*>(row + 2 * m);
        const Pair2 b = *reinterpret_cast<const Pair2*>(row + 2 * NP + 2 * m);
        X[m] = fma_(b.y, bx.s[1], fma_(b.x, bx.v[1], fma_(a.y, bx.s[0], a.x * bx.v[0])));
        if (GRAD) Xu[m] = fma_(b.y, bx.ds[1], fma_(b.x, bx.dv[1], fma_(a.y, bx.ds[0], a.x * bx.dv[0])));
    }
}

// 3-D: the whole slot [cz][cy][cx][8] -> g[0] = value, g[1..3] = d/du, d/dv, d/dw (unit-cell coordinates)
template <bool GRAD>
ARB_HD void eval3(const double* slot, const double* f, double* g) {
    const Basis bx = hermite(f[0]), by = hermite(f[1]), bz = hermite(f[2]);
    double val = 0.0, gu = 0.0, gv = 0.0, gw = 0.0;
    ARB_UNROLL
    for (int cz = 0; cz < 2; ++cz) {
        double Y[2] = {0.0, 0.0}, Yu[2] = {0.0, 0.0}, Yv[2] = {0.0, 0.0};      // index tz
        ARB_UNROLL
        for (int cy = 0; cy < 2; ++cy) {
            double X[4], Xu[4];                                                 // index ty + 2 tz
            row_x<4, GRAD>(slot + (cz * 2 + cy) * 16, bx, X, Xu);
            ARB_UNROLL
            for (int tz = 0; tz < 2; ++tz) {
                Y[tz] = fma_(X[2 * tz + 1], by.s[cy], fma_(X[2 * tz], by.v[cy], Y[tz]));
                if (GRAD) {
                    Yu[tz] = fma_(Xu[2 * tz + 1], by.s[cy], fma_(Xu[2 * tz], by.v[cy], Yu[tz]));
                    Yv[tz] = fma_(X[2 * tz + 1], by.ds[cy], fma_(X[2 * tz], by.dv[cy], Yv[tz]));
                }
            }
        }
        val = fma_(Y[1], bz.s[cz], fma_(Y[0], bz.v[cz], val));
        if (GRAD) {
            gu = fma_(Yu[1], bz.s[cz], fma_(Yu[0], bz.v[cz], gu));
            gv = fma_(Yv[1], bz.s[cz], fma_(Yv[0], bz.v[cz], gv));
            gw = fma_(Y[1], bz.ds[cz], fma_(Y[0], bz.dv[cz], gw));
        }
    }
    g[0] = val;
    if (GRAD) { g[1] = gu; g[2] = gv; g[3] = gw; }
}

// 3-D, interleaved components, one lane = one row (cy, cz): slot [cx][c][8] (c = 0..3: Bx, By, Bz, |B|) -> this lane's
// share of out[0..2] = the three components, and (BOTH) out[3] = |B|, out[4..6] = its partials; four lanes add up.
template <bool BOTH>
ARB_HD void eval3_row(const double* slot, int cy, int cz, const double* f, double* out) {
    const Basis bx = hermite(f[0]), by = hermite(f[1]), bz = hermite(f[2]);
    const double yv = cy ? by.v[1] : by.v[0], ys = cy ? by.s[1] : by.s[0];
    const double zv = cz ? bz.v[1] : bz.v[0], zs = cz ? bz.s[1] : bz.s[0];
    const double w[4] = {yv * zv, ys * zv, yv * zs, ys * zs};                    // index ty + 2 tz
    ARB_UNROLL
    for (int c = 0; c < (BOTH ? 4 : 3); ++c) {
        double acc = 0.0, accu = 0.0, accv = 0.0, accw = 0.0;
        ARB_UNROLL
        for (int m = 0; m < 4; ++m) {
            const Pair2 a = *reinterpret_cast<const Pair2*>(slot + c * 8 + 2 * m);
            const Pair2 b = *reinterpret_cast<const Pair2*>(slot + 32 + c * 8 + 2 * m);
            const double X = fma_(b.y, bx.s[1], fma_(b.x, bx.v[1], fma_(a.y, bx.s[0], a.x * bx.v[0])));
            acc = fma_(X, w[m], acc);
            if (BOTH && c == 3) {
                const double dyv = cy ? by.dv[1] : by.dv[0], dys = cy ? by.ds[1] : by.ds[0];
                const double dzv = cz ? bz.dv[1] : bz.dv[0], dzs = cz ? bz.ds[1] : bz.ds[0];
                const double Xu = fma_(b.y, bx.ds[1], fma_(b.x, bx.dv[1], fma_(a.y, bx.ds[0], a.x * bx.dv[0])));
                accu = fma_(Xu, w[m], accu);
                accv = fma_(X, ((m & 1) ? dys : dyv) * ((m >> 1) ? zs : zv), accv);
                accw = fma_(X, ((m & 1) ? ys : yv) * ((m >> 1) ? dzs : dzv), accw);
            }
        }
        out[c] = acc;
        if (BOTH && c == 3) { out[4] = accu; out[5] = accv; out[6] = accw; }
    }
}

// 4-D quirk term of one lane: f15 = fxyzt of the lane's corners (cx, cy) = (0,0), (1,0), (0,1), (1,1) in reference
// order c = cx + 2 cy + 4 cz + 8 ct; e[c] = fxyzt(c - 1) - fxyzt(c); prev15 = fxyzt of corner (1, 1) of the lane before
// (cz + 2 ct - 1), 0 for the first lane.  Adds to g.
template <bool GRAD>
ARB_HD void quirk4_lane(const double (&f15)[4], double prev15, int cz, int ct, const double* f, double* g) {
    const Basis bx = hermite(f[0]), by = hermite(f[1]), bz = hermite(f[2]), bt = hermite(f[3]);
    const double zs = cz ? bz.s[1] : bz.s[0], ts = ct ? bt.s[1] : bt.s[0];
    const double e00 = prev15 - f15[0], e10 = f15[0] - f15[1], e01 = f15[1] - f15[2], e11 = f15[2] - f15[3];
    const double r0 = fma_(e10, bx.s[1], e00 * bx.s[0]), r1 = fma_(e11, bx.s[1], e01 * bx.s[0]);
    const double q = fma_(r1, by.s[1], r0 * by.s[0]);
    g[0] = fma_(q, zs * ts, g[0]);
    if (GRAD) {
        const double dzs = cz ? bz.ds[1] : bz.ds[0], dts = ct ? bt.ds[1] : bt.ds[0];
        const double ru0 = fma_(e10, bx.ds[1], e00 * bx.ds[0]), ru1 = fma_(e11, bx.ds[1], e01 * bx.ds[0]);
        const double qu = fma_(ru1, by.s[1], ru0 * by.s[0]);
        const double qv = fma_(r1, by.ds[1], r0 * by.ds[0]);
        g[1] = fma_(qu, zs * ts, g[1]);
        g[2] = fma_(qv, zs * ts, g[2]);
        g[3] = fma_(q, dzs * ts, g[3]);
        g[4] = fma_(q, zs * dts, g[4]);
    }
}

// 4-D, one lane = one (cz, ct): slot [cy][cx][16] -> this lane's share of value and the four partials; the four
// lanes' shares add up.  QUIRK: the A.py:860 term is added here (the kernel adds it after its shuffle instead).
template <bool GRAD, bool QUIRK>
ARB_HD void eval4_lane(const double* slot, int cz, int ct, const double* f, double prev15, double* g) {
    const Basis bx = hermite(f[0]), by = hermite(f[1]), bz = hermite(f[2]), bt = hermite(f[3]);
    double Y[4] = {0.0, 0.0, 0.0, 0.0}, Yu[4] = {0.0, 0.0, 0.0, 0.0}, Yv[4] = {0.0, 0.0, 0.0, 0.0};   // index tz + 2 tt
    ARB_UNROLL
    for (int cy = 0; cy < 2; ++cy) {
        double X[8], Xu[8];                                                     // index ty + 2 tz + 4 tt
        row_x<8, GRAD>(slot + cy * 32, bx, X, Xu);
        ARB_UNROLL
        for (int m = 0; m < 4; ++m) {
            Y[m] = fma_(X[2 * m + 1], by.s[cy], fma_(X[2 * m], by.v[cy], Y[m]));
            if (GRAD) {
                Yu[m] = fma_(Xu[2 * m + 1], by.s[cy], fma_(Xu[2 * m], by.v[cy], Yu[m]));
                Yv[m] = fma_(X[2 * m + 1], by.ds[cy], fma_(X[2 * m], by.dv[cy], Yv[m]));
            }
        }
    }
    const double zv = cz ? bz.v[1] : bz.v[0], zs = cz ? bz.s[1] : bz.s[0];
    const double tv = ct ? bt.v[1] : bt.v[0], ts = ct ? bt.s[1] : bt.s[0];
    // Z[tt] = sum_tz Y[tz + 2 tt] z-weight
    const double Z0 = fma_(Y[1], zs, Y[0] * zv), Z1 = fma_(Y[3], zs, Y[2] * zv);
    g[0] = fma_(Z1, ts, Z0 * tv);
    if (GRAD) {
        const double dzv = cz ? bz.dv[1] : bz.dv[0], dzs = cz ? bz.ds[1] : bz.ds[0];
        const double dtv = ct ? bt.dv[1] : bt.dv[0], dts = ct ? bt.ds[1] : bt.ds[0];
        const double U0 = fma_(Yu[1], zs, Yu[0] * zv), U1 = fma_(Yu[3], zs, Yu[2] * zv);
        const double V0 = fma_(Yv[1], zs, Yv[0] * zv), V1 = fma_(Yv[3], zs, Yv[2] * zv);
        const double W0 = fma_(Y[1], dzs, Y[0] * dzv), W1 = fma_(Y[3], dzs, Y[2] * dzv);
        g[1] = fma_(U1, ts, U0 * tv);
        g[2] = fma_(V1, ts, V0 * tv);
        g[3] = fma_(W1, ts, W0 * tv);
        g[4] = fma_(Z1, dts, Z0 * dtv);
    }
    if (QUIRK) {
        const double f15[4] = {slot[15], slot[16 + 15], slot[32 + 15], slot[48 + 15]};
        quirk4_lane<GRAD>(f15, prev15, cz, ct, f, g);
    }
}

}  // namespace nodes
}  // namespace arb
