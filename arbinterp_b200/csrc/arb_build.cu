// Coefficient build: for every interpolatable cell, alpha = inv(B) * (D * f[4^d neighbourhood]).
// Replaces calcCoefficients1/2/3 + neighbourInd + allCoeffs of the reference
// (A.py:570-621, 523-525 tricubic; A.py:1322-1381, 1260-1262 quadcubic), which does the same with
// one Python-level dgemv of the fused matrix A = inv(B) D per cell and component.
//
// Three formulations of the same solve live here (arb_set_build_variant picks one; all write the same
// cell-major table and agree to round-off):
//   * separable, FP64 pipe (default; build_sep3_kernel / build_sep4_kernel, phases in arb_build_sep.cuh):
//     A factorises into 1-D line transforms, one pass per axis over a TMA-staged grid tile; ~200 FP64
//     instructions per 3-D cell, so the kernel runs at the HBM write roofline.
//   * dense, FP64 tensor pipe (build_kernel): one CTA builds one tile of cells for one component,
//       1. TMA (cp.async.bulk.tensor, SASS UTMALDG) stages the grid tile plus its 3-point halo into
//          shared memory; out-of-range points are zero-filled by the TMA unit and only ever feed
//          cells that are not stored.
//       2. stencil stage: the 2^d derivative fields (f, fx, fy, ..., fxyz[t]; central differences in
//          unit-cell coordinates, the rows of D, A.py:129-173 / 762-876) are evaluated at every
//          cell-corner point of the tile and kept in shared memory.  A cell's b-vector is a gather
//          of 2^d corners x 2^d types from these fields, so it is never materialised.
//       3. solve stage: alpha^T[cells x 4^d] = b^T * inv(B)^T on the FP64 tensor cores (mma.sync m8n8k4 f64,
//          SASS DMMA); inv(B) is integer, |entries| <= 27 / 81, read as pre-swizzled fragments from an
//          L1/L2-resident buffer.  Accumulators go straight to the table with 16-byte stores.
//   * Kronecker-factored, FP64 tensor pipe (build_kron_kernel): same stencil stage, then TWO small DMMA
//     contractions that exploit inv(B) = G_hi (x) G_lo.
// The 4-D reference matrix carries the A.py:860 off-by-one (row 240 of D is zero and rows
// 241..255 use the stencil centre of the previous corner); the DMMA kernels reproduce it in the b-vector
// gather, where the quadruple-mixed type reads corner c-1 (and 0 for c = 0), the separable kernel adds the
// equivalent rank-16 term in its t pass.
#include <cstdlib>
#include <vector>
#include <mutex>
#include <atomic>
#include "arb_common.cuh"
#include "arb_build_sep.cuh"

namespace arb {

// Tile configurations.  MT = TY*TZ*TT x-rows of 8 cells per CTA; every warp owns NT_W n-tiles (8
// coefficients each) for all MT m-tiles, so a thread holds MT*NT_W*2 FP64 accumulators.  Smaller
// tiles trade halo re-reads (served by L2) for more resident warps per SM, which is what keeps the
// DMMA pipe busy while other CTAs are in their TMA / stencil / store phases.
struct Cfg3A { static constexpr int D = 3, TX = 8, TY = 4, TZ = 4, TT = 1, WARPS = 4, NT_W = 2, MINB = 3, MINB_KRON = 6; };
struct Cfg3B { static constexpr int D = 3, TX = 8, TY = 4, TZ = 2, TT = 1, WARPS = 4, NT_W = 2, MINB = 5, MINB_KRON = 6; };
struct Cfg3C { static constexpr int D = 3, TX = 8, TY = 2, TZ = 2, TT = 1, WARPS = 4, NT_W = 2, MINB = 6, MINB_KRON = 8; };
struct Cfg4A { static constexpr int D = 4, TX = 8, TY = 2, TZ = 2, TT = 2, WARPS = 8, NT_W = 4, MINB = 1, MINB_KRON = 3; };
struct Cfg4B { static constexpr int D = 4, TX = 8, TY = 2, TZ = 2, TT = 1, WARPS = 8, NT_W = 4, MINB = 2, MINB_KRON = 3; };
struct Cfg4C { static constexpr int D = 4, TX = 8, TY = 2, TZ = 1, TT = 1, WARPS = 8, NT_W = 4, MINB = 3, MINB_KRON = 4; };

template <typename Cfg>
struct BuildShape {
    static constexpr int D = Cfg::D;
    static constexpr int NM = (D == 3) ? 64 : 256;
    static constexpr int NTYPE = 1 << D;
    static constexpr int MT = Cfg::TY * Cfg::TZ * Cfg::TT;          // m-tiles (x-rows of 8 cells)
    static constexpr int GX = 12;                                   // TX + 3 rounded up to 16 bytes
    static constexpr int GY = Cfg::TY + 3, GZ = Cfg::TZ + 3, GT = (D == 4) ? Cfg::TT + 3 : 1;
    static constexpr int GRID_ELEMS = GX * GY * GZ * GT;
    static constexpr int PX = 9, PY = Cfg::TY + 1, PZ = Cfg::TZ + 1, PT = (D == 4) ? Cfg::TT + 1 : 1;
    static constexpr int PXS = 10;                                  // padded x pitch of the derivative fields
    static constexpr int NPOINT = PX * PY * PZ * PT;
    static constexpr int TYPE_STRIDE = PXS * PY * PZ * PT;
    static constexpr int DERIV_ELEMS = TYPE_STRIDE * NTYPE;
    static constexpr int THREADS = Cfg::WARPS * 32;
    static constexpr size_t SMEM = (size_t)(GRID_ELEMS + DERIV_ELEMS) * 8 + 128;
};

struct BuildParams {
    double* table;
    const double* bfrag;       // inv(B)^T fragments: [kstep][ntile][lane]
    int64_t nc[4];             // cells per axis
    int64_t ntile[4];          // tiles per axis
    int ncomp;
    int quirk;
    unsigned char type_of_mask[16];   // derivative-axis bitmask -> b-vector type index (A.py:118-125 order)
    unsigned long long type_nibbles;  // the same map packed 4 bits per mask (register-only lookup)
};

// Central differences of one grid point over its 3x3x3 neighbourhood, all 8 axis subsets at once:
// o[mask], bit a of mask set = differentiated along axis a; unit spacing (the rows of D, A.py:132-173).
// Differences are taken axis by axis (0.5*(g+ - g-)), so a mixed derivative is the reference's
// +-0.25 / +-0.125 stencil up to the association order of the additions (scaling by 0.5 is exact).
__device__ __forceinline__ void stencil_cube(const double* g, int sx, int sy, int sz, double (&o)[8]) {
    double x0[3][3], x1[3][3];
#pragma unroll
    for (int dz = 0; dz < 3; ++dz)
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
            const double* r = g + (dz - 1) * sz + (dy - 1) * sy;
            x0[dz][dy] = r[0];
            x1[dz][dy] = 0.5 * (r[sx] - r[-sx]);
        }
    double y00[3], y01[3], y10[3], y11[3];     // [ydiff][xdiff][dz]
#pragma unroll
    for (int dz = 0; dz < 3; ++dz) {
        y00[dz] = x0[dz][1];
        y01[dz] = x1[dz][1];
        y10[dz] = 0.5 * (x0[dz][2] - x0[dz][0]);
        y11[dz] = 0.5 * (x1[dz][2] - x1[dz][0]);
    }
    o[0] = y00[1];                         // f
    o[1] = y01[1];                         // fx
    o[2] = y10[1];                         // fy
    o[3] = y11[1];                         // fxy
    o[4] = 0.5 * (y00[2] - y00[0]);        // fz
    o[5] = 0.5 * (y01[2] - y01[0]);        // fxz
    o[6] = 0.5 * (y10[2] - y10[0]);        // fyz
    o[7] = 0.5 * (y11[2] - y11[0]);        // fxyz
}

template <typename Cfg>
__global__ void __launch_bounds__(BuildShape<Cfg>::THREADS, Cfg::MINB)
build_kernel(const __grid_constant__ CUtensorMap tmap, const BuildParams p) {
    using S = BuildShape<Cfg>;
    constexpr int D = Cfg::D;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* gtile = reinterpret_cast<double*>(smem_raw);            // [GT][GZ][GY][GX]
    double* deriv = gtile + S::GRID_ELEMS;                          // [type][PT][PZ][PY][PXS]
    __shared__ uint64_t bar;

    // tile coordinates: blockIdx.x = tile (x fastest), blockIdx.y = component
    int64_t tl = blockIdx.x;
    const int64_t tx = tl % p.ntile[0]; tl /= p.ntile[0];
    const int64_t ty = tl % p.ntile[1]; tl /= p.ntile[1];
    const int64_t tz = (D == 4) ? tl % p.ntile[2] : tl;
    const int64_t tt = (D == 4) ? tl / p.ntile[2] : 0;
    const int comp = blockIdx.y;
    const int x0 = (int)(tx * Cfg::TX), y0 = (int)(ty * Cfg::TY), z0 = (int)(tz * Cfg::TZ), t0 = (int)(tt * Cfg::TT);

    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
        mbar_expect_tx(&bar, S::GRID_ELEMS * 8);
        if (D == 3) tma_load_4d(gtile, &tmap, &bar, x0, y0, z0, comp);
        else tma_load_5d(gtile, &tmap, &bar, x0, y0, z0, t0, comp);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NKS = S::NM / 4;
    // first B fragments travel while the tile is in flight
    double bf[Cfg::NT_W];
#pragma unroll
    for (int j = 0; j < Cfg::NT_W; ++j) bf[j] = __ldg(p.bfrag + ((size_t)(warp * Cfg::NT_W + j)) * 32 + lane);
    __syncthreads();
    mbar_wait(&bar, 0);

    // ---- stencil stage: one thread per corner point, all 2^D derivative types in registers ----
    // corner point (px,py,pz,pt) of the tile sits at grid-tile coordinate (+1,+1,+1,+1)
    for (int e = threadIdx.x; e < S::NPOINT; e += S::THREADS) {
        int r = e;
        const int px = r % S::PX; r /= S::PX;
        const int py = r % S::PY; r /= S::PY;
        const int pz = r % S::PZ;
        const int pt = r / S::PZ;
        constexpr int SX = 1, SY = S::GX, SZ = S::GX * S::GY, ST = S::GX * S::GY * S::GZ;
        const int centre = (((D == 4 ? (pt + 1) : 0) * S::GZ + (pz + 1)) * S::GY + (py + 1)) * S::GX + (px + 1);
        double* dst = deriv + ((pt * S::PZ + pz) * S::PY + py) * S::PXS + px;
        if (D == 3) {
            double o[8];
            stencil_cube(gtile + centre, SX, SY, SZ, o);
#pragma unroll
            for (int m = 0; m < 8; ++m) dst[p.type_of_mask[m] * S::TYPE_STRIDE] = o[m];
        } else {
            double lo[8], mid[8], hi[8];
            stencil_cube(gtile + centre - ST, SX, SY, SZ, lo);
            stencil_cube(gtile + centre, SX, SY, SZ, mid);
            stencil_cube(gtile + centre + ST, SX, SY, SZ, hi);
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                dst[p.type_of_mask[m] * S::TYPE_STRIDE] = mid[m];
                dst[p.type_of_mask[m | 8] * S::TYPE_STRIDE] = 0.5 * (hi[m] - lo[m]);
            }
        }
    }
    __syncthreads();

    // ---- solve stage (DMMA) ----------------------------------------------------------
    const int cellx = lane >> 2;              // A-fragment row = cell along x
    const int kq = lane & 3;                  // A-fragment column = corner (cx, cy) within a k-step
    double acc[S::MT][Cfg::NT_W][2];
#pragma unroll
    for (int m = 0; m < S::MT; ++m)
#pragma unroll
        for (int j = 0; j < Cfg::NT_W; ++j) acc[m][j][0] = acc[m][j][1] = 0.0;

    constexpr int KS_PER_TYPE = S::NTYPE / 4;     // k-steps per derivative type: 2 (3-D), 4 (4-D)
    const int lane_off = (kq >> 1) * S::PXS + cellx + (kq & 1);   // (cy, cx) part of the address
#pragma unroll 1
    for (int ks = 0; ks < NKS; ++ks) {
        double bn[Cfg::NT_W];                     // next k-step's B fragments (L1/L2) overlap this step's DMMAs
        const int ksn = (ks + 1 < NKS) ? ks + 1 : ks;
#pragma unroll
        for (int j = 0; j < Cfg::NT_W; ++j)
            bn[j] = __ldg(p.bfrag + ((size_t)ksn * (S::NM / 8) + warp * Cfg::NT_W + j) * 32 + lane);
        const int type = ks / KS_PER_TYPE;
        const int chi = ks % KS_PER_TYPE;         // high corner bits: cz (+ 2 ct)
        const double* dbase = deriv + type * S::TYPE_STRIDE;
        const bool shifted = (D == 4) && p.quirk && (type == S::NTYPE - 1);
#pragma unroll
        for (int m = 0; m < S::MT; ++m) {
            const int my = m % Cfg::TY, mz = (m / Cfg::TY) % Cfg::TZ, mt = m / (Cfg::TY * Cfg::TZ);
            double a;
            if (!shifted) {
                const int cz = chi & 1, ct = chi >> 1;
                a = dbase[(((mt + ct) * S::PZ + (mz + cz)) * S::PY + my) * S::PXS + lane_off];
            } else {
                // A.py:860: b[240 + c] = stencil at corner c-1, b[240] = 0
                const int c = chi * 4 + kq - 1;
                const int cx = c & 1, cy = (c >> 1) & 1, cz = (c >> 2) & 1, ct = (c >> 3) & 1;
                a = (c < 0) ? 0.0
                            : dbase[(((mt + ct) * S::PZ + (mz + cz)) * S::PY + (my + cy)) * S::PXS + cellx + cx];
            }
#pragma unroll
            for (int j = 0; j < Cfg::NT_W; ++j) dmma_884(acc[m][j][0], acc[m][j][1], a, bf[j]);
        }
#pragma unroll
        for (int j = 0; j < Cfg::NT_W; ++j) bf[j] = bn[j];
    }

    // ---- store: C fragment row = cell (lane>>2), columns = coefficients 2*(lane&3)+{0,1} --
    const int64_t cx = x0 + cellx;
#pragma unroll
    for (int m = 0; m < S::MT; ++m) {
        const int my = m % Cfg::TY, mz = (m / Cfg::TY) % Cfg::TZ, mt = m / (Cfg::TY * Cfg::TZ);
        const int64_t cy = y0 + my, cz = z0 + mz, ct = t0 + mt;
        bool ok = (cx < p.nc[0]) && (cy < p.nc[1]) && (cz < p.nc[2]);
        int64_t cell = cx + p.nc[0] * (cy + p.nc[1] * cz);
        if (D == 4) {
            ok = ok && (ct < p.nc[3]);
            cell += p.nc[0] * p.nc[1] * p.nc[2] * ct;
        }
        if (!ok) continue;
        double* dst = p.table + (cell * p.ncomp + comp) * S::NM + (warp * Cfg::NT_W) * 8 + 2 * kq;
#pragma unroll
        for (int j = 0; j < Cfg::NT_W; ++j) stg_stream_d2(dst + j * 8, acc[m][j][0], acc[m][j][1]);
    }
}

// ======================================================================================
// Kronecker-factored solve (build variants 9 and 4)
// ======================================================================================
// inv(B) is a Kronecker product of the 1-D cubic Hermite inverse H (4x4, integer):
//     alpha[kl][ji] = sum_{s_hi, s_lo} G_hi[kl][s_hi] * G_lo[ji][s_lo] * b[s_hi][s_lo]
// with "lo" = the (x, y) pair, "hi" = z (3-D) or the (z, t) pair (4-D), s_a = 2*(differentiated along a)
// + (corner bit along a), G_lo[i+4j][sx+4sy] = H[i][sx] H[j][sy] and G_hi likewise.  The solve then is two
// batched dense contractions with 4x4 / 16x16 blocks instead of one with a 64x64 / 256x256 block:
//   step A  U[cell][kl][s_lo]  = sum_{s_hi} b[cell][s_hi][s_lo] * G_hi^T      rows (cell, s_lo), K = s_hi
//   step B  alpha[cell][kl][ji] = sum_{s_lo} U[cell][kl][s_lo]  * G_lo^T      rows (cell, kl),   K = s_lo
// 6 DMMAs per 3-D cell instead of 16, 32 per 4-D cell instead of 256, no inv(B) traffic at all (the
// fragments of G_hi and G_lo are built from H in registers), ~40 registers per thread.  After the stencil
// stage the warps are independent: each owns a share of the tile's cells and a private U buffer, so the
// solve needs no block-level barrier.  Same b-vector gather as the dense kernel (including the A.py:860
// quirk), same table layout, results equal to round-off (different summation order).
__constant__ double c_H[16] = {1, 0, 0, 0, 0, 0, 1, 0, -3, 3, -2, -1, 2, -2, 1, 1};   // H[i][s], s = (f0, f1, f0', f1')

template <typename Cfg>
struct KronShape {
    using S = BuildShape<Cfg>;
    static constexpr int D = Cfg::D;
    static constexpr int NHI = (D == 3) ? 4 : 16;          // size of the "hi" index (k or (k,l))
    static constexpr int GC = (D == 3) ? 4 : 1;            // cells per warp pass
    static constexpr int KS_A = NHI / 4, NT_A = (NHI + 7) / 8;
    static constexpr int MT_A = GC * 2;                    // 16 s_lo rows per cell
    static constexpr int MT_B = GC * NHI / 8;
    static constexpr int US = 16;                          // s_lo pitch of U; columns are XOR-swizzled by the row
    static constexpr int U_PER_WARP = GC * NHI * US;
    static constexpr int NCELL = Cfg::TX * Cfg::TY * Cfg::TZ * Cfg::TT;
    // Derivative fields are stored by dm = fx | fz<<1 | fy<<2 | ft<<3 with a pitch == 4 (mod 16) doubles: the 16
    // lanes of a half-warp of a step-A gather differ in (fx, bit_x, fz, bit_z) only, and dm*KTS + bit_x +
    // bit_z*(PY*PXS) then hits 16 distinct 8-byte bank pairs (PY*PXS == 2 or 14 mod 16 for the tiles used).
    static constexpr int KTS = ((S::TYPE_STRIDE + 11) / 16) * 16 + 4;
    static constexpr int DERIV_K = KTS * S::NTYPE;
    static_assert(KTS >= S::TYPE_STRIDE, "padded pitch too small");
    static constexpr size_t SMEM = (size_t)(S::GRID_ELEMS + DERIV_K + Cfg::WARPS * U_PER_WARP) * 8 + 128;
};

template <typename Cfg>
__global__ void __launch_bounds__(BuildShape<Cfg>::THREADS, Cfg::MINB_KRON)
build_kron_kernel(const __grid_constant__ CUtensorMap tmap, const BuildParams p) {
    using S = BuildShape<Cfg>;
    using K = KronShape<Cfg>;
    constexpr int D = Cfg::D;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* gtile = reinterpret_cast<double*>(smem_raw);
    double* deriv = gtile + S::GRID_ELEMS;
    double* ubase = deriv + K::DERIV_K;
    __shared__ uint64_t bar;

    int64_t tl = blockIdx.x;
    const int64_t tx = tl % p.ntile[0]; tl /= p.ntile[0];
    const int64_t ty = tl % p.ntile[1]; tl /= p.ntile[1];
    const int64_t tz = (D == 4) ? tl % p.ntile[2] : tl;
    const int64_t tt = (D == 4) ? tl / p.ntile[2] : 0;
    const int comp = blockIdx.y;
    const int x0 = (int)(tx * Cfg::TX), y0 = (int)(ty * Cfg::TY), z0 = (int)(tz * Cfg::TZ), t0 = (int)(tt * Cfg::TT);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
        mbar_expect_tx(&bar, S::GRID_ELEMS * 8);
        if (D == 3) tma_load_4d(gtile, &tmap, &bar, x0, y0, z0, comp);
        else tma_load_5d(gtile, &tmap, &bar, x0, y0, z0, t0, comp);
    }
    __syncthreads();
    mbar_wait(&bar, 0);

    // ---- stencil stage: one thread per corner point, all 2^D derivative types in registers ----
    // corner point (px,py,pz,pt) of the tile sits at grid-tile coordinate (+1,+1,+1,+1)
    for (int e = threadIdx.x; e < S::NPOINT; e += S::THREADS) {
        int r = e;
        const int px = r % S::PX; r /= S::PX;
        const int py = r % S::PY; r /= S::PY;
        const int pz = r % S::PZ;
        const int pt = r / S::PZ;
        constexpr int SX = 1, SY = S::GX, SZ = S::GX * S::GY, ST = S::GX * S::GY * S::GZ;
        const int centre = (((D == 4 ? (pt + 1) : 0) * S::GZ + (pz + 1)) * S::GY + (py + 1)) * S::GX + (px + 1);
        double* dst = deriv + ((pt * S::PZ + pz) * S::PY + py) * S::PXS + px;
        if (D == 3) {
            double o[8];
            stencil_cube(gtile + centre, SX, SY, SZ, o);
#pragma unroll
            for (int m = 0; m < 8; ++m) dst[((m & 1) | (((m >> 2) & 1) << 1) | (((m >> 1) & 1) << 2)) * K::KTS] = o[m];
        } else {
            double lo[8], mid[8], hi[8];
            stencil_cube(gtile + centre - ST, SX, SY, SZ, lo);
            stencil_cube(gtile + centre, SX, SY, SZ, mid);
            stencil_cube(gtile + centre + ST, SX, SY, SZ, hi);
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const int dm = (m & 1) | (((m >> 2) & 1) << 1) | (((m >> 1) & 1) << 2);
                dst[dm * K::KTS] = mid[m];
                dst[(dm | 8) * K::KTS] = 0.5 * (hi[m] - lo[m]);
            }
        }
    }
    __syncthreads();


    // ---- constant fragments (B operands): element [k = 4*ks + q][n = 8*nt + r], q = lane&3, r = lane>>2 ----
    const int q = lane & 3, r = lane >> 2;
    double bA[K::KS_A][K::NT_A], bB[4][2];
#pragma unroll
    for (int ks = 0; ks < K::KS_A; ++ks)
#pragma unroll
        for (int nt = 0; nt < K::NT_A; ++nt) {
            const int shi = 4 * ks + q, kl = 8 * nt + r;             // G_hi[kl][s_hi]
            double v = (kl < K::NHI) ? c_H[(kl & 3) * 4 + (shi & 3)] : 0.0;
            if (D == 4) v *= c_H[((kl >> 2) & 3) * 4 + (shi >> 2)];
            bA[ks][nt] = v;
        }
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
            const int slo = 4 * ks + q, ji = 8 * nt + r;             // G_lo[ji][s_lo]
            bB[ks][nt] = c_H[(ji & 3) * 4 + (slo & 3)] * c_H[(ji >> 2) * 4 + (slo >> 2)];
        }

    // ---- per-lane gather offsets of the b-vector element [s_hi = 4ks + q][s_lo = 8*half + r] ----
    // s = 2*flag + bit per axis; type = type_of_mask[flags]; the point is the cell corner + bits
    int lo_off[2], lo_mask[2], lo_c[2];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int slo = 8 * half + r, sx = slo & 3, sy = slo >> 2;
        lo_mask[half] = (sx >> 1) | ((sy >> 1) << 2);             // dm bit layout: fx | fz<<1 | fy<<2 | ft<<3
        lo_off[half] = (sy & 1) * S::PXS + (sx & 1);
        lo_c[half] = (sx & 1) | ((sy & 1) << 1);
    }
    int hi_off[K::KS_A], hi_mask[K::KS_A], hi_c[K::KS_A];
#pragma unroll
    for (int ks = 0; ks < K::KS_A; ++ks) {
        const int shi = 4 * ks + q, sz = shi & 3, st = shi >> 2;
        hi_mask[ks] = ((sz >> 1) << 1) | ((st >> 1) << 3);
        hi_off[ks] = ((st & 1) * S::PZ + (sz & 1)) * S::PY * S::PXS;
        hi_c[ks] = ((sz & 1) << 2) | ((st & 1) << 3);
    }
    double* U = ubase + warp * K::U_PER_WARP;

    for (int g = warp; g < K::NCELL / K::GC; g += Cfg::WARPS) {
        const int c0 = g * K::GC;
        // ---- step A: rows (cell, s_lo), K = s_hi, N = kl ----
#pragma unroll
        for (int ma = 0; ma < K::MT_A; ++ma) {
            const int ci = c0 + (ma >> 1), half = ma & 1;
            const int cx = ci % Cfg::TX, cy = (ci / Cfg::TX) % Cfg::TY, cz = (ci / (Cfg::TX * Cfg::TY)) % Cfg::TZ,
                      ct = ci / (Cfg::TX * Cfg::TY * Cfg::TZ);
            const int cell_off = ((ct * S::PZ + cz) * S::PY + cy) * S::PXS + cx;
            double acc[K::NT_A][2];
#pragma unroll
            for (int nt = 0; nt < K::NT_A; ++nt) acc[nt][0] = acc[nt][1] = 0.0;
#pragma unroll
            for (int ks = 0; ks < K::KS_A; ++ks) {
                const int mask = lo_mask[half] | hi_mask[ks];
                double a;
                if (D == 4 && p.quirk && mask == 15) {
                    // A.py:860: b[240 + c] = stencil at corner c-1, b[240] = 0
                    const int c = (lo_c[half] | hi_c[ks]) - 1;
                    const int off = (((c >> 3) & 1) * S::PZ + ((c >> 2) & 1)) * S::PY * S::PXS + ((c >> 1) & 1) * S::PXS + (c & 1);
                    a = (c < 0) ? 0.0 : deriv[15 * K::KTS + cell_off + off];
                } else {
                    a = deriv[mask * K::KTS + cell_off + lo_off[half] + hi_off[ks]];
                }
#pragma unroll
                for (int nt = 0; nt < K::NT_A; ++nt) dmma_884(acc[nt][0], acc[nt][1], a, bA[ks][nt]);
            }
            // C[row = s_lo = 8*half + r][col = kl = 8*nt + 2q + {0,1}]  ->  U[cell][kl][s_lo ^ swz(kl)].
            // swz(kl) = 4*((kl ^ kl>>1) & 3) makes both this store (lanes differ in q and r) and the step-B
            // load (lanes differ in kl = row and in s_lo & 3) hit 16 distinct bank pairs per half-warp.
#pragma unroll
            for (int nt = 0; nt < K::NT_A; ++nt) {
                const int kl = 8 * nt + 2 * q;
                if (kl < K::NHI) {
                    double* dst = U + ((ma >> 1) * K::NHI + kl) * K::US;
                    const int col = 8 * half + r;
                    dst[col ^ (4 * ((kl ^ (kl >> 1)) & 3))] = acc[nt][0];
                    dst[K::US + (col ^ (4 * (((kl + 1) ^ ((kl + 1) >> 1)) & 3)))] = acc[nt][1];
                }
            }
        }
        __syncwarp();
        // ---- step B: rows (cell, kl), K = s_lo, N = ji ----
#pragma unroll
        for (int mb = 0; mb < K::MT_B; ++mb) {
            const int R = mb * 8 + r;                         // row within the group
            const int cg = R / K::NHI, kl = R % K::NHI;
            double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
            const double* urow = U + (cg * K::NHI + kl) * K::US;
            const int swz = 4 * ((kl ^ (kl >> 1)) & 3);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const double a = urow[(4 * ks + q) ^ swz];
                dmma_884(acc[0][0], acc[0][1], a, bB[ks][0]);
                dmma_884(acc[1][0], acc[1][1], a, bB[ks][1]);
            }
            const int ci = c0 + cg;
            const int64_t gx = x0 + ci % Cfg::TX, gy = y0 + (ci / Cfg::TX) % Cfg::TY,
                          gz = z0 + (ci / (Cfg::TX * Cfg::TY)) % Cfg::TZ, gt = t0 + ci / (Cfg::TX * Cfg::TY * Cfg::TZ);
            bool ok = (gx < p.nc[0]) && (gy < p.nc[1]) && (gz < p.nc[2]);
            int64_t cell = gx + p.nc[0] * (gy + p.nc[1] * gz);
            if (D == 4) {
                ok = ok && (gt < p.nc[3]);
                cell += p.nc[0] * p.nc[1] * p.nc[2] * gt;
            }
            if (ok) {
                double* dst = p.table + (cell * p.ncomp + comp) * S::NM + kl * 16 + 2 * q;
                stg_stream_d2(dst, acc[0][0], acc[0][1]);
                stg_stream_d2(dst + 8, acc[1][0], acc[1][1]);
            }
        }
        __syncwarp();
    }
}


// ======================================================================================
// Separable build on the FP64 pipe (build variants 0 and 5..8; phases in arb_build_sep.cuh)
// ======================================================================================
// 3-D: one CTA = one tile of 8 x TY x TZ cells of one component.  TMA stages the grid tile, then the x, y
// and z passes of the 1-D line transform run back to back; the z pass writes the table.
template <typename S3, int MINB>
__global__ void __launch_bounds__(S3::THREADS, MINB)
build_sep3_kernel(const __grid_constant__ CUtensorMap tmap, const sep::SepParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* g = reinterpret_cast<double*>(smem_raw);
    double* X = g + S3::G_ELEMS;
    double* Y = X + S3::X_ELEMS;
    __shared__ uint64_t bar;

    int64_t tl = blockIdx.x;
    int comp = blockIdx.y;
    if (p.comp_fast) { comp = (int)(tl % p.ncomp); tl /= p.ncomp; }
    const int64_t tx = tl % p.ntile[0]; tl /= p.ntile[0];
    const int64_t ty = tl % p.ntile[1];
    const int64_t tz = tl / p.ntile[1];
    const int x0 = (int)(tx * 8), y0 = (int)(ty * S3::TY), z0 = (int)(tz * S3::TZ);
    const int tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
        mbar_expect_tx(&bar, S3::G_ELEMS * 8);
        tma_load_4d(g, &tmap, &bar, x0, y0, z0, comp);
    }
    __syncthreads();
    mbar_wait(&bar, 0);
    sep::pass_x(g, X, S3::NROW, tid, S3::THREADS);
    __syncthreads();
    sep::pass_y(X, Y, S3::GZ, S3::TY, tid, S3::THREADS);
    __syncthreads();
    S3::pass_z_emit(Y, p, x0, y0, z0, comp, tid, S3::THREADS);
}

// 4-D: one CTA = one column of 8 x 2 x 2 cells of one component, marching over p.lt cell layers along t.
// Grid planes arrive through a two-deep TMA pipeline (one mbarrier per buffer); every plane is transformed
// along x, y, z once, and the fused z/t phase emits the layer the plane completes.  Software pipeline: iteration
// s runs {x pass of plane s+1, first half of emit(s)} | barrier | {y pass of plane s+1, second half of emit(s)} |
// barrier, so the latency-bound passes always share an interval with table stores (Y and the corner values are double-buffered).
template <bool QUIRK>
__global__ void __launch_bounds__(sep::Sep4::THREADS, 2)
build_sep4_kernel(const __grid_constant__ CUtensorMap tmap, const sep::SepParams p) {
    using S4 = sep::Sep4;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* base = reinterpret_cast<double*>(smem_raw);
    double* plane = base + S4::OFF_PLANE;
    double* X = base + S4::OFF_X;
    double* Y = base + S4::OFF_Y;
    double* ring = base + S4::OFF_RING;
    double* w3ring = base + S4::OFF_W3;
    double* gbuf = base + S4::OFF_G;
    __shared__ uint64_t bar[2];

    int64_t tl = blockIdx.x;
    int comp = blockIdx.y;
    if (p.comp_fast) { comp = (int)(tl % p.ncomp); tl /= p.ncomp; }
    const int64_t tx = tl % p.ntile[0]; tl /= p.ntile[0];
    const int64_t ty = tl % p.ntile[1];
    const int64_t tz = tl / p.ntile[1];
    const int x0 = (int)(tx * 8), y0 = (int)(ty * S4::TY), z0 = (int)(tz * S4::TZ);
    const int64_t t0 = (int64_t)blockIdx.z * p.lt;                 // first cell layer of this CTA
    const int nlayer = (int)((p.nc[3] - t0 < p.lt) ? (p.nc[3] - t0) : p.lt);
    const int nstep = nlayer + 3;                                  // grid planes t0 .. t0 + nlayer + 2
    const int tid = threadIdx.x;
    auto fetch = [&](int q) {                                      // plane q -> buffer q & 1 (one thread)
        mbar_expect_tx(&bar[q & 1], S4::PLANE * 8);
        tma_load_5d(plane + (q & 1) * S4::PLANE_PITCH, &tmap, &bar[q & 1], x0, y0, z0, (int)t0 + q, comp);
    };
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        fence_mbar_init();
        fetch(0);
        fetch(1);
    }
    // the thread's two emit tasks (the same in every step: their ring elements are private to the thread)
    const S4::ETask task0 = S4::make_task(tid, p, x0, y0, z0, t0, comp);
    const S4::ETask task1 = S4::make_task(tid + S4::NTASK_E / 2, p, x0, y0, z0, t0, comp);
    const int64_t layer_stride = p.nc[0] * p.nc[1] * p.nc[2] * p.ncomp * 256;
    __syncthreads();
    // prologue: plane 0 through the x and y passes
    mbar_wait(&bar[0], 0);
    S4::phase_a<QUIRK>(plane, X, w3ring, tid, S4::THREADS);
    __syncthreads();
    if (tid == 0) { fence_proxy_async(); fetch(2); }              // nstep >= 4 always
    S4::phase_b<QUIRK>(X, Y, w3ring, gbuf, 0, tid, S4::THREADS);
    __syncthreads();
#pragma unroll 1
    for (int s = 0; s < nstep; ++s) {
        const int q = s + 1;                                       // plane whose x/y passes ride along
        const bool more = q < nstep;
        const double* Ys = Y + (s & 1) * S4::Y_ELEMS;
        const double* gs = gbuf + (s & 1) * S4::G_ELEMS;
        if (more) {
            mbar_wait(&bar[q & 1], (q >> 1) & 1);
            S4::phase_a<QUIRK>(plane + (q & 1) * S4::PLANE_PITCH, X, w3ring + (q & 3) * S4::W3_PITCH, tid, S4::THREADS);
        }
        S4::emit_task<QUIRK>(task0, Ys, ring, gs, p.table, layer_stride, s);
        __syncthreads();
        if (more) {
            if (tid == 0 && q + 2 < nstep) { fence_proxy_async(); fetch(q + 2); }
            S4::phase_b<QUIRK>(X, Y + (q & 1) * S4::Y_ELEMS, w3ring, gbuf + (q & 1) * S4::G_ELEMS, q, tid, S4::THREADS);
        }
        S4::emit_task<QUIRK>(task1, Ys, ring, gs, p.table, layer_stride, s);
        __syncthreads();
    }
}

__global__ void fill_nan_kernel(double* p, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = __longlong_as_double(0x7ff8000000000000LL);
}

// --------------------------------------------------------------------------------------
// host side
// --------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    });
    return fn;
}

// inv(B)^T fragments for mma m8n8k4 (B operand: element [k = lane&3][n = lane>>2]) per device
static std::mutex g_frag_mutex;
static double* g_frag[64][2];

static int get_bfrag(int d, const double** out) {
    int dev = 0;
    ARB_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_frag_mutex);
    double*& slot = g_frag[dev & 63][d - 3];
    if (!slot) {
        const int nm = 1 << (2 * d);
        std::vector<double> ib((size_t)nm * nm), frag((size_t)nm * nm);
        make_invB(d, ib.data());
        for (int ks = 0; ks < nm / 4; ++ks)
            for (int nt = 0; nt < nm / 8; ++nt)
                for (int lane = 0; lane < 32; ++lane) {
                    const int k = ks * 4 + (lane & 3), n = nt * 8 + (lane >> 2);
                    frag[((size_t)ks * (nm / 8) + nt) * 32 + lane] = ib[(size_t)n * nm + k];
                }
        ARB_CUDA(cudaMalloc(&slot, sizeof(double) * frag.size()));
        ARB_CUDA(cudaMemcpy(slot, frag.data(), sizeof(double) * frag.size(), cudaMemcpyHostToDevice));
    }
    *out = slot;
    return 0;
}

static std::atomic<int> g_build_variant{0};   // measurement knob (arb_set_build_variant)

// Tensor map over the grid [C][nt][nz][ny][nx] (x fastest) with the given box; TMA needs 16-byte global
// strides, so an odd nx is padded to even in a stream-ordered scratch copy (`padded`, freed by the caller
// after the launch).  Out-of-range box elements are zero-filled.
struct GridMap {
    CUtensorMap tmap;
    double* padded = nullptr;
    cudaStream_t stream = nullptr;
    GridMap() = default;
    GridMap(const GridMap&) = delete;
    GridMap& operator=(const GridMap&) = delete;
    // stream-ordered release on every exit path (error returns included): the free is queued behind the
    // kernels that read the copy
    void release() {
        if (padded) cudaFreeAsync(padded, stream);
        padded = nullptr;
    }
    ~GridMap() { release(); }
};

static int make_grid_map(int D, const double* grid, int ncomp, const int64_t* n, const cuuint32_t* box_dims,
                         cudaStream_t st, GridMap* out) {
    EncodeTiledFn encode = get_encode_fn();
    if (!encode) { set_error("arb_build_coeffs: cuTensorMapEncodeTiled not available from the driver"); return 2; }
    const double* src = grid;
    int64_t pitch = n[0];
    int64_t rows = ncomp;
    for (int a = 1; a < D; ++a) rows *= n[a];
    out->stream = st;
    if (n[0] & 1) {
        pitch = n[0] + 1;
        ARB_CUDA(cudaMallocAsync(&out->padded, sizeof(double) * pitch * rows, st));
        ARB_CUDA(cudaMemsetAsync(out->padded, 0, sizeof(double) * pitch * rows, st));
        ARB_CUDA(cudaMemcpy2DAsync(out->padded, pitch * 8, grid, n[0] * 8, n[0] * 8, rows, cudaMemcpyDeviceToDevice, st));
        src = out->padded;
    }
    if ((reinterpret_cast<uintptr_t>(src) & 15) != 0) {
        out->release();
        set_error("arb_build_coeffs: grid pointer must be 16-byte aligned");
        return 1;
    }
    cuuint64_t gdim[5], gstr[4];
    cuuint32_t box[5], estr[5];
    const int rank = D + 1;
    gdim[0] = (cuuint64_t)n[0];
    int64_t stride = pitch * 8;
    for (int a = 1; a < D; ++a) { gdim[a] = (cuuint64_t)n[a]; gstr[a - 1] = (cuuint64_t)stride; stride *= n[a]; }
    gdim[D] = (cuuint64_t)ncomp; gstr[D - 1] = (cuuint64_t)stride;
    for (int a = 0; a < D; ++a) box[a] = box_dims[a];
    box[D] = 1;
    for (int a = 0; a < rank; ++a) estr[a] = 1;
    CUresult cr = encode(&out->tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, rank, const_cast<double*>(src), gdim, gstr, box,
                         estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
        out->release();
        set_error("arb_build_coeffs: cuTensorMapEncodeTiled failed with CUresult %d", (int)cr);
        return 2;
    }
    if (getenv("ARB_DEBUG_TMAP")) {
        const unsigned long long* w = reinterpret_cast<const unsigned long long*>(&out->tmap);
        fprintf(stderr, "[arb] tensor map rank %d box %u %u %u:", rank, box[0], box[1], box[2]);
        for (int i = 0; i < 16; ++i) fprintf(stderr, " %016llx", w[i]);
        fprintf(stderr, "\n");
    }
    return 0;
}

// NaN sentinel row behind the last cell (A.py:45, 57, 70) and release of the padded scratch copy
static int finish_build(int D, int ncomp, int64_t ncell, double* table, GridMap* gm, cudaStream_t st) {
    const int64_t tail = (int64_t)ncomp * (1 << (2 * D));
    fill_nan_kernel<<<(unsigned)((tail + 255) / 256), 256, 0, st>>>(table + ncell * tail, tail);
    ARB_CUDA(cudaGetLastError());
    gm->release();
    return 0;
}

template <typename Cfg, bool KRON = false>
static int build_impl(const double* grid, int ncomp, const int64_t* n, double* table, int quirk, cudaStream_t st) {
    using S = BuildShape<Cfg>;
    constexpr int D = Cfg::D;
    BuildParams p;
    memset(&p, 0, sizeof(p));
    int64_t ncell = 1;
    const int tdim[4] = {Cfg::TX, Cfg::TY, Cfg::TZ, Cfg::TT};
    int64_t ntiles = 1;
    for (int a = 0; a < D; ++a) {
        if (n[a] < 4) { set_error("arb_build_coeffs: axis %d has %lld points, need >= 4", a, (long long)n[a]); return 1; }
        p.nc[a] = n[a] - 3;
        ncell *= p.nc[a];
        p.ntile[a] = (p.nc[a] + tdim[a] - 1) / tdim[a];
        ntiles *= p.ntile[a];
    }
    for (int a = D; a < 4; ++a) { p.nc[a] = 1; p.ntile[a] = 1; }
    if (ntiles > 0x7fffffffLL) { set_error("arb_build_coeffs: too many tiles (%lld)", (long long)ntiles); return 1; }
    p.table = table; p.ncomp = ncomp; p.quirk = quirk;
    for (int r = 0; r < (1 << D); ++r) {
        p.type_of_mask[deriv_mask(D, r)] = (unsigned char)r;
        p.type_nibbles |= (unsigned long long)r << (4 * deriv_mask(D, r));
    }
    { const int frc = get_bfrag(D, &p.bfrag); if (frc) return frc; }

    GridMap gm;
    const cuuint32_t box[4] = {S::GX, S::GY, S::GZ, S::GT};
    { const int rc = make_grid_map(D, grid, ncomp, n, box, st, &gm); if (rc) return rc; }

    dim3 gridDim((unsigned)ntiles, (unsigned)ncomp, 1);
    if (KRON) {
        auto k = build_kron_kernel<Cfg>;
        ARB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KronShape<Cfg>::SMEM));
        k<<<gridDim, S::THREADS, KronShape<Cfg>::SMEM, st>>>(gm.tmap, p);
    } else {
        auto k = build_kernel<Cfg>;
        ARB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::SMEM));
        k<<<gridDim, S::THREADS, S::SMEM, st>>>(gm.tmap, p);
    }
    ARB_CUDA(cudaGetLastError());
    return finish_build(D, ncomp, ncell, table, &gm, st);
}

static int sep_params(int D, int ncomp, const int64_t* n, double* table, int quirk, const int* tdim,
                      sep::SepParams* p, int64_t* ncell, int64_t* ntiles) {
    memset(p, 0, sizeof(*p));
    *ncell = 1;
    *ntiles = 1;
    for (int a = 0; a < D; ++a) {
        if (n[a] < 4) { set_error("arb_build_coeffs: axis %d has %lld points, need >= 4", a, (long long)n[a]); return 1; }
        p->nc[a] = n[a] - 3;
        *ncell *= p->nc[a];
        if (a < 3) {
            p->ntile[a] = (p->nc[a] + tdim[a] - 1) / tdim[a];
            *ntiles *= p->ntile[a];
        }
    }
    for (int a = D; a < 4; ++a) p->nc[a] = 1;
    if (*ntiles > 0x7fffffffLL) { set_error("arb_build_coeffs: too many tiles (%lld)", (long long)*ntiles); return 1; }
    p->table = table; p->ncomp = ncomp; p->quirk = quirk;
    return 0;
}

template <int TY, int TZ, int THREADS, int MINB>
static int build_sep3_impl(const double* grid, int ncomp, const int64_t* n, double* table, int comp_fast,
                           cudaStream_t st) {
    using S3 = sep::Sep3<TY, TZ, THREADS>;
    sep::SepParams p;
    int64_t ncell, ntiles;
    const int tdim[3] = {8, TY, TZ};
    { const int rc = sep_params(3, ncomp, n, table, 0, tdim, &p, &ncell, &ntiles); if (rc) return rc; }
    GridMap gm;
    const cuuint32_t box[4] = {sep::GX, S3::GY, S3::GZ, 1};
    { const int rc = make_grid_map(3, grid, ncomp, n, box, st, &gm); if (rc) return rc; }
    auto k = build_sep3_kernel<S3, MINB>;
    ARB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S3::SMEM));
    p.comp_fast = comp_fast && ntiles * ncomp <= 0x7fffffffLL;
    const dim3 blocks = p.comp_fast ? dim3((unsigned)(ntiles * ncomp), 1, 1) : dim3((unsigned)ntiles, (unsigned)ncomp, 1);
    k<<<blocks, THREADS, S3::SMEM, st>>>(gm.tmap, p);
    ARB_CUDA(cudaGetLastError());
    return finish_build(3, ncomp, ncell, table, &gm, st);
}

// layers_per_cta <= 0: long marches, cut only as far as needed to fill the GPU with a few waves of CTAs
static int build_sep4_impl(const double* grid, int ncomp, const int64_t* n, double* table, int quirk,
                           int layers_per_cta, int comp_fast, cudaStream_t st) {
    using S4 = sep::Sep4;
    sep::SepParams p;
    int64_t ncell, ntiles;
    const int tdim[3] = {8, S4::TY, S4::TZ};
    { const int rc = sep_params(4, ncomp, n, table, quirk, tdim, &p, &ncell, &ntiles); if (rc) return rc; }
    int64_t lt = layers_per_cta;
    if (lt <= 0) {
        const int64_t want = 8LL * 2 * num_sms();              // CTAs for ~8 waves at 2 CTAs per SM
        int64_t chunks = (want + ntiles * ncomp - 1) / (ntiles * ncomp);
        const int64_t max_chunks = (p.nc[3] + 5) / 6;            // keep marches >= 6 layers (3 planes of run-in each)
        if (chunks > max_chunks) chunks = max_chunks;
        if (chunks < 1) chunks = 1;
        lt = (p.nc[3] + chunks - 1) / chunks;
    }
    if (lt > p.nc[3]) lt = p.nc[3];
    p.lt = (int)lt;
    const int64_t nchunk = (p.nc[3] + lt - 1) / lt;
    if (nchunk > 65535) { set_error("arb_build_coeffs: too many t chunks (%lld)", (long long)nchunk); return 1; }
    GridMap gm;
    const cuuint32_t box[4] = {sep::GX, S4::GY, S4::GZ, 1};
    { const int rc = make_grid_map(4, grid, ncomp, n, box, st, &gm); if (rc) return rc; }
    auto k = quirk ? build_sep4_kernel<true> : build_sep4_kernel<false>;
    ARB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S4::SMEM));
    p.comp_fast = comp_fast && ntiles * ncomp <= 0x7fffffffLL;
    const dim3 blocks = p.comp_fast ? dim3((unsigned)(ntiles * ncomp), 1, (unsigned)nchunk)
                                    : dim3((unsigned)ntiles, (unsigned)ncomp, (unsigned)nchunk);
    k<<<blocks, S4::THREADS, S4::SMEM, st>>>(gm.tmap, p);
    ARB_CUDA(cudaGetLastError());
    return finish_build(4, ncomp, ncell, table, &gm, st);
}

}  // namespace arb

extern "C" {

int arb_build_coeffs(int d, const double* grid, int ncomp, const int64_t n[4], double* table, int reference_quirk,
                     void* stream) {
    if (!grid || !table || !n) { arb::set_error("arb_build_coeffs: null pointer"); return 1; }
    if (ncomp < 1 || ncomp > 4) { arb::set_error("arb_build_coeffs: ncomp=%d not in 1..4", ncomp); return 1; }
    cudaStream_t st = (cudaStream_t)stream;
    const int v = arb::g_build_variant.load(std::memory_order_relaxed);
    // 0 (default) and 5..8: separable FP64-pipe kernels (build_sep3_kernel / build_sep4_kernel) in different tile /
    //    march / block-order configurations; 1..3: dense 4^d x 4^d DMMA contraction (1 = its best tile);
    // 9, 4: Kronecker-factored DMMA solve (9 = its best tile).  profiles/r01_build_variants.log
    if (d == 3) {
        if (v == 1) return arb::build_impl<arb::Cfg3A>(grid, ncomp, n, table, reference_quirk, st);
        if (v == 2) return arb::build_impl<arb::Cfg3B>(grid, ncomp, n, table, reference_quirk, st);
        if (v == 3) return arb::build_impl<arb::Cfg3C>(grid, ncomp, n, table, reference_quirk, st);
        if (v == 4) return arb::build_impl<arb::Cfg3B, true>(grid, ncomp, n, table, reference_quirk, st);
        if (v == 9) return arb::build_impl<arb::Cfg3A, true>(grid, ncomp, n, table, reference_quirk, st);
        if (v == 5) return arb::build_sep3_impl<4, 4, 128, 4>(grid, ncomp, n, table, 0, st);
        if (v == 6) return arb::build_sep3_impl<4, 4, 128, 4>(grid, ncomp, n, table, 1, st);
        if (v == 7) return arb::build_sep3_impl<4, 8, 256, 2>(grid, ncomp, n, table, 1, st);
        return arb::build_sep3_impl<2, 4, 128, 6>(grid, ncomp, n, table, 1, st);
    }
    if (d == 4) {
        if (v == 1) return arb::build_impl<arb::Cfg4B>(grid, ncomp, n, table, reference_quirk, st);
        if (v == 2) return arb::build_impl<arb::Cfg4A>(grid, ncomp, n, table, reference_quirk, st);
        if (v == 3) return arb::build_impl<arb::Cfg4C>(grid, ncomp, n, table, reference_quirk, st);
        if (v == 4) return arb::build_impl<arb::Cfg4B, true>(grid, ncomp, n, table, reference_quirk, st);
        if (v == 9) return arb::build_impl<arb::Cfg4A, true>(grid, ncomp, n, table, reference_quirk, st);
        if (v == 5) return arb::build_sep4_impl(grid, ncomp, n, table, reference_quirk, 0, 0, st);
        if (v == 7) return arb::build_sep4_impl(grid, ncomp, n, table, reference_quirk, 8, 1, st);
        if (v == 8) return arb::build_sep4_impl(grid, ncomp, n, table, reference_quirk, 3, 1, st);
        return arb::build_sep4_impl(grid, ncomp, n, table, reference_quirk, 0, 1, st);
    }
    arb::set_error("arb_build_coeffs: d=%d not in {3,4}", d);
    return 1;
}

int arb_set_build_variant(int variant) {
    return arb::g_build_variant.exchange(variant);
}

int arb_build_coeffs_3d(const double* grid, int ncomp, int64_t nx, int64_t ny, int64_t nz, double* table,
                        void* stream) {
    const int64_t n[4] = {nx, ny, nz, 1};
    return arb_build_coeffs(3, grid, ncomp, n, table, 1, stream);
}

int arb_build_coeffs_4d(const double* grid, int ncomp, int64_t nx, int64_t ny, int64_t nz, int64_t nt, double* table,
                        void* stream) {
    const int64_t n[4] = {nx, ny, nz, nt};
    return arb_build_coeffs(4, grid, ncomp, n, table, 1, stream);
}
}
