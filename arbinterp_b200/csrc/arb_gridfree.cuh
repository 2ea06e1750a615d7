// Math of the table-free query kernels (arb_query.cu: query_grid_kernel, query_grid4_kernel), written
// __host__ __device__ so that tests/host_emul/gridfree4_host_emul.cu can run it on the CPU.
//
// 3-D: A == M (x) M (x) M, so value = sum f[k][j][i] wz_k wy_j wx_i with the Catmull-Rom weights
// w(t) = M^T [1, t, t^2, t^3].
// 4-D: the reference matrix is M^(x)4 plus the rank-16 term of A.py:860 (see arb_build_sep.cuh):
//     value = sum_l wt_l * (tricubic contraction of grid plane l)
//           + sum_c hx[cx] hy[cy] hz[cz] ht[ct] e[c],     e[c] = fxyzt(corner c-1) - fxyzt(corner c), fxyzt(-1) = 0,
// h[0](t) = t - 2t^2 + t^3, h[1](t) = t^3 - t^2 the cubic Hermite slope basis, and fxyzt(corner c) the
// 1/16 (+-) central difference over the 16 neighbourhood points whose index parity per axis equals c's bits.
// One lane owns one grid plane l of the 4^4 neighbourhood: it contracts the plane in (x, y, z) and, for the
// quirk, accumulates the plane's 8 signed parity sums S[cz][cy][cx]; fxyzt(c3, ct) = (S_{l=ct+2} - S_{l=ct})/16.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ARB_HD __host__ __device__ __forceinline__
#else
#define ARB_HD inline
#endif
#if defined(__CUDA_ARCH__)
#define ARB_UNROLL _Pragma("unroll")
#else
#define ARB_UNROLL
#endif

namespace arb {
namespace gridfree {

struct alignas(16) Pair2 { double x, y; };

ARB_HD double fma_(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return fma(a, b, c);
#else
    return __builtin_fma(a, b, c);
#endif
}

ARB_HD void catmull_rom(double t, double (&w)[4], double (&dw)[4]) {
    const double t2 = t * t, t3 = t2 * t;
    w[0] = fma_(-0.5, t3, fma_(1.0, t2, -0.5 * t));
    w[1] = fma_(1.5, t3, fma_(-2.5, t2, 1.0));
    w[2] = fma_(-1.5, t3, fma_(2.0, t2, 0.5 * t));
    w[3] = fma_(0.5, t3, -0.5 * t2);
    dw[0] = fma_(-1.5, t2, fma_(2.0, t, -0.5));
    dw[1] = fma_(4.5, t2, -5.0 * t);
    dw[2] = fma_(-4.5, t2, fma_(4.0, t, 0.5));
    dw[3] = fma_(1.5, t2, -t);
}

// cubic Hermite slope basis (the columns Hq of the Hermite inverse applied to [1, t, t^2, t^3]) and derivative
ARB_HD void hermite_slope(double t, double (&h)[2], double (&dh)[2]) {
    const double t2 = t * t;
    h[0] = fma_(t2, t - 2.0, t);
    h[1] = t2 * (t - 1.0);
    dh[0] = fma_(t, fma_(3.0, t, -4.0), 1.0);
    dh[1] = t * fma_(3.0, t, -2.0);
}

ARB_HD double sel4(const double (&v)[4], int i) { return i == 0 ? v[0] : (i == 1 ? v[1] : (i == 2 ? v[2] : v[3])); }

struct PlanePartial {
    double val, gx, gy, gz;   // tricubic contraction of the plane and its partials in (u, v, w)
    double S[8];              // signed parity sums, index cx + 2 cy + 4 cz (quirk only)
};

// One grid plane of the neighbourhood from its TMA box: 16 rows (row = 4k + j) of 6 doubles, 48 bytes apart,
// the neighbourhood's x index i sits at box position i + off (off = ix & 1, fp64 boxes start on even x).
// Rows are visited in the lane's rotated order (j by rj, k by rk: keeps the 16-byte shared-memory reads of a
// quarter-warp on distinct banks although the slots are 768 bytes apart).
template <bool GRAD, bool QUIRK>
ARB_HD void plane_partial(const unsigned char* box, int off, int rj, int rk, const double (&wx)[4],
                          const double (&dwx)[4], const double (&wy)[4], const double (&dwy)[4],
                          const double (&wz)[4], const double (&dwz)[4], PlanePartial& out) {
    double wyr[4], dwyr[4], wzr[4], dwzr[4], sj[4], sk[4];
    ARB_UNROLL
    for (int j = 0; j < 4; ++j) {
        const int jj = (j + rj) & 3, kk = (j + rk) & 3;
        wyr[j] = sel4(wy, jj); dwyr[j] = sel4(dwy, jj);
        wzr[j] = sel4(wz, kk); dwzr[j] = sel4(dwz, kk);
        sj[j] = (jj >= 2) ? 1.0 : -1.0;
        sk[j] = (kk >= 2) ? 1.0 : -1.0;
    }
    double T[2][2][2] = {{{0.0, 0.0}, {0.0, 0.0}}, {{0.0, 0.0}, {0.0, 0.0}}};   // [k & 1][j & 1][cx] by loop position
    double val = 0.0, gx = 0.0, gy = 0.0, gz = 0.0;
    ARB_UNROLL
    for (int k = 0; k < 4; ++k) {
        double P = 0.0, Px = 0.0, Py = 0.0;
        ARB_UNROLL
        for (int j = 0; j < 4; ++j) {
            const int row = 4 * ((k + rk) & 3) + ((j + rj) & 3);
            const Pair2* r = reinterpret_cast<const Pair2*>(box + row * 48);
            const Pair2 a = r[0], b = r[1], c = r[2];
            const double e0 = off ? a.y : a.x, e1 = off ? b.x : a.y, e2 = off ? b.y : b.x, e3 = off ? c.x : b.y;
            const double pp = fma_(e3, wx[3], fma_(e2, wx[2], fma_(e1, wx[1], e0 * wx[0])));
            P = fma_(wyr[j], pp, P);
            if (GRAD) {
                const double dp = fma_(e3, dwx[3], fma_(e2, dwx[2], fma_(e1, dwx[1], e0 * dwx[0])));
                Px = fma_(wyr[j], dp, Px);
                Py = fma_(dwyr[j], pp, Py);
            }
            if (QUIRK) {
                const double s = sj[j] * sk[k];
                T[k & 1][j & 1][0] = fma_(s, e2 - e0, T[k & 1][j & 1][0]);
                T[k & 1][j & 1][1] = fma_(s, e3 - e1, T[k & 1][j & 1][1]);
            }
        }
        val = fma_(wzr[k], P, val);
        if (GRAD) {
            gx = fma_(wzr[k], Px, gx);
            gy = fma_(wzr[k], Py, gy);
            gz = fma_(dwzr[k], P, gz);
        }
    }
    out.val = val; out.gx = gx; out.gy = gy; out.gz = gz;
    if (QUIRK) {
        // loop position parity -> index parity: (k + rk) & 1 = (k & 1) ^ (rk & 1), likewise j
        const int fk = rk & 1, fj = rj & 1;
        ARB_UNROLL
        for (int ck = 0; ck < 2; ++ck)
            ARB_UNROLL
            for (int cj = 0; cj < 2; ++cj)
                ARB_UNROLL
                for (int cx = 0; cx < 2; ++cx) {
                    const double same_k = fj ? T[ck][cj ^ 1][cx] : T[ck][cj][cx];
                    const double flip_k = fj ? T[ck ^ 1][cj ^ 1][cx] : T[ck ^ 1][cj][cx];
                    out.S[cx + 2 * cj + 4 * ck] = fk ? flip_k : same_k;
                }
    }
}

// 3-D table-free, interleaved components: one lane = z-plane k of the 4x4x4 neighbourhood, slot [j][i][c]
// (c = 0..3: Bx, By, Bz, |B|) -> this lane's share of out[0..2] = the components and (BOTH) out[3] = |B|,
// out[4..6] = its partials; the four lanes (k = 0..3) add up.  A == M (x) M (x) M.
template <bool BOTH>
ARB_HD void plane_il(const double* slot, int k, const double* f, double* out) {
    double wx[4], dwx[4], wy[4], dwy[4], wz[4], dwz[4];
    catmull_rom(f[0], wx, dwx);
    catmull_rom(f[1], wy, dwy);
    catmull_rom(f[2], wz, dwz);
    double P[4] = {0.0, 0.0, 0.0, 0.0}, Px = 0.0, Py = 0.0;
    ARB_UNROLL
    for (int j = 0; j < 4; ++j) {
        double pp[4] = {0.0, 0.0, 0.0, 0.0}, dp = 0.0;
        ARB_UNROLL
        for (int i = 0; i < 4; ++i) {
            const Pair2 a = *reinterpret_cast<const Pair2*>(slot + (j * 4 + i) * 4);
            const Pair2 b = *reinterpret_cast<const Pair2*>(slot + (j * 4 + i) * 4 + 2);
            pp[0] = fma_(a.x, wx[i], pp[0]);
            pp[1] = fma_(a.y, wx[i], pp[1]);
            pp[2] = fma_(b.x, wx[i], pp[2]);
            if (BOTH) {
                pp[3] = fma_(b.y, wx[i], pp[3]);
                dp = fma_(b.y, dwx[i], dp);
            }
        }
        ARB_UNROLL
        for (int c = 0; c < (BOTH ? 4 : 3); ++c) P[c] = fma_(wy[j], pp[c], P[c]);
        if (BOTH) {
            Px = fma_(wy[j], dp, Px);
            Py = fma_(dwy[j], pp[3], Py);
        }
    }
    const double wk = sel4(wz, k);
    out[0] = wk * P[0]; out[1] = wk * P[1]; out[2] = wk * P[2];
    if (BOTH) {
        out[3] = wk * P[3];
        out[4] = wk * Px;
        out[5] = wk * Py;
        out[6] = sel4(dwz, k) * P[3];
    }
}

// 4-D table-free, interleaved components (grid [nt][nz][ny][nx][4]): one lane = z-plane k of the 4^4 neighbourhood, one
// pass = t-plane l, slot [j][i][c] as in plane_il.  out[0..6] = this (k, l) plane's share of the (u, v, w) contraction
// with the z weight applied (the caller applies the t weight): components 0..2 and (BOTH) |B|, its d/du, d/dv, d/dw.
// QUIRK: Pq[cx + 2 cy][c] = the plane's signed xy parity sums  sum_{i, j} (+-) f[j][i][c]  over i in {cx, cx + 2},
// j in {cy, cy + 2} (sign + when i and j are both the upper or both the lower point) -- the xy part of the 1/16 (+-)
// central difference fxyzt at the cell's corners (A.py:860 term, see the header of this file).
template <bool BOTH, bool QUIRK>
ARB_HD void plane_il4(const double* slot, double wzk, double dwzk, const double (&wx)[4], const double (&dwx)[4],
                      const double (&wy)[4], const double (&dwy)[4], double (&out)[7], double (&Pq)[4][4]) {
    constexpr int NC = BOTH ? 4 : 3;
    double P[4] = {0.0, 0.0, 0.0, 0.0}, Px = 0.0, Py = 0.0;
    if (QUIRK) {
        ARB_UNROLL
        for (int q = 0; q < 4; ++q)
            ARB_UNROLL
            for (int c = 0; c < 4; ++c) Pq[q][c] = 0.0;
    }
    ARB_UNROLL
    for (int j = 0; j < 4; ++j) {
        double pp[4] = {0.0, 0.0, 0.0, 0.0}, dp = 0.0;
        ARB_UNROLL
        for (int i = 0; i < 4; ++i) {
            const Pair2 a = *reinterpret_cast<const Pair2*>(slot + (j * 4 + i) * 4);
            const Pair2 b = *reinterpret_cast<const Pair2*>(slot + (j * 4 + i) * 4 + 2);
            const double v[4] = {a.x, a.y, b.x, b.y};
            ARB_UNROLL
            for (int c = 0; c < NC; ++c) {
                pp[c] = fma_(v[c], wx[i], pp[c]);
                if (QUIRK) Pq[(j & 1) * 2 + (i & 1)][c] = ((i >> 1) == (j >> 1)) ? Pq[(j & 1) * 2 + (i & 1)][c] + v[c]
                                                                                 : Pq[(j & 1) * 2 + (i & 1)][c] - v[c];
            }
            if (BOTH) dp = fma_(v[3], dwx[i], dp);
        }
        ARB_UNROLL
        for (int c = 0; c < NC; ++c) P[c] = fma_(wy[j], pp[c], P[c]);
        if (BOTH) {
            Px = fma_(wy[j], dp, Px);
            Py = fma_(dwy[j], pp[3], Py);
        }
    }
    out[0] = wzk * P[0]; out[1] = wzk * P[1]; out[2] = wzk * P[2];
    if (BOTH) {
        out[3] = wzk * P[3];
        out[4] = wzk * Px;
        out[5] = wzk * Py;
        out[6] = dwzk * P[3];
    }
}

// Quirk term of one ct before the t factor: c[0] = sum_c3 hx hy hz e[c3], c[1..3] = its partials in u, v, w.
// g = fxyzt at the 8 corners of this ct, g7prev = fxyzt(corner 7 of ct-1) for ct = 1, 0 for ct = 0.
template <bool GRAD>
ARB_HD void corner_term(const double (&g)[8], double g7prev, const double (&hx)[2], const double (&dhx)[2],
                        const double (&hy)[2], const double (&dhy)[2], const double (&hz)[2], const double (&dhz)[2],
                        double (&c)[4]) {
    double e[8];
    e[0] = g7prev - g[0];
    ARB_UNROLL
    for (int i = 1; i < 8; ++i) e[i] = g[i - 1] - g[i];
    double v = 0.0, du = 0.0, dv = 0.0, dw = 0.0;
    ARB_UNROLL
    for (int cz = 0; cz < 2; ++cz) {
        double q = 0.0, qu = 0.0, qv = 0.0;
        ARB_UNROLL
        for (int cy = 0; cy < 2; ++cy) {
            const double e0 = e[4 * cz + 2 * cy], e1 = e[4 * cz + 2 * cy + 1];
            const double px = fma_(hx[1], e1, hx[0] * e0);
            q = fma_(hy[cy], px, q);
            if (GRAD) {
                const double pdx = fma_(dhx[1], e1, dhx[0] * e0);
                qu = fma_(hy[cy], pdx, qu);
                qv = fma_(dhy[cy], px, qv);
            }
        }
        v = fma_(hz[cz], q, v);
        if (GRAD) {
            du = fma_(hz[cz], qu, du);
            dv = fma_(hz[cz], qv, dv);
            dw = fma_(dhz[cz], q, dw);
        }
    }
    c[0] = v; c[1] = du; c[2] = dv; c[3] = dw;
}

}  // namespace gridfree
}  // namespace arb
