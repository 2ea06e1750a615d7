// Separable coefficient build (the default; build variants 0 and 5..8): the phases of build_sep3_kernel / build_sep4_kernel.
//
// The Lekien-Marsden matrix of the reference, A = inv(B) D (A.py:175, 878), factorises: in 3-D
// A == M (x) M (x) M exactly, M = 1/2 [[0,2,0,0],[-1,0,1,0],[2,-5,4,-1],[-1,3,-3,1]] being the 1-D map from
// four consecutive grid values f(-1), f(0), f(1), f(2) to the cubic's monomial coefficients (central-
// difference slopes at 0 and 1 followed by the cubic Hermite inverse).  alpha = A f[4x4x4] is then three
// passes of that 4 -> 4 line transform, one per axis, and neighbouring cells share the early passes:
//     X[z][y][cell x][i]            from the grid tile        (pass along x)
//     Y[z][cell y][cell x][j][i]    from X                    (pass along y)
//     alpha[cell][k][j][i]          from Y, straight to HBM   (pass along z)
// ~170 FP64 instructions and ~10 shared-memory wavefronts per cell against 6 DMMAs + 30 wavefronts of the
// Kronecker/tensor-core kernel, so the build becomes a pure HBM-write stream.
//
// In 4-D the reference matrix is M^(x)4 plus a rank-16 term caused by A.py:860 (D's rows 241..255 use the
// stencil centre of the previous corner and row 240 stays zero): b_ref = b_true + e with
// e[240 + c] = fxyzt(corner c-1) - fxyzt(corner c) (fxyzt(-1) := 0), and
//     alpha_ref[l][k][j][i] = (M^(x)4 f)[l][k][j][i] + sum_c Hq[i][cx] Hq[j][cy] Hq[k][cz] Hq[l][ct] e[240 + c],
// Hq = [[0,0],[1,0],[-2,-1],[1,1]] the slope columns of the Hermite inverse.  The 4-D kernel marches along t:
// each new grid plane goes through the x, y and z passes once, the three previous planes' results wait in
// a shared-memory ring, and the t pass emits one layer of cells per step.  The march is software-pipelined:
// the x/y passes of plane s+1 share their two barrier intervals with the two halves of step s's emit phase.
//
// Everything here is __host__ __device__ and written as per-thread task loops ("for (e = tid; e < tasks; e += nthreads)")
// over shared arrays, so tests/host_emul/sep_host_emul.cu can run the very same phases on the CPU and compare
// them with A f (tests/test_host_logic.py::test_separable_build_phases_on_host).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ARB_HD __host__ __device__ __forceinline__
#else
#define ARB_HD inline
#endif
#if defined(__CUDA_ARCH__)
#define ARB_SEP_UNROLL _Pragma("unroll")
#else
#define ARB_SEP_UNROLL
#endif

namespace arb {
namespace sep {

struct SepParams {
    double* table;
    int64_t nc[4];        // cells per axis
    int64_t ntile[3];     // tiles along x, y, z
    int ncomp;
    int quirk;            // 4-D: reproduce A.py:860
    int lt;               // 4-D: cell layers per CTA along t
    int comp_fast;        // blockIdx.x = tile * ncomp + comp (the components of a tile are built side by side)
};

// one line: grid values at -1, 0, 1, 2 -> monomial coefficients a0..a3 (row e of M applied to f)
ARB_HD void cr_line(double f0, double f1, double f2, double f3, double& a0, double& a1, double& a2, double& a3) {
    a0 = f1;
    a1 = 0.5 * (f2 - f0);
    a2 = (f0 - 2.5 * f1) + (2.0 * f2 - 0.5 * f3);
    a3 = 0.5 * (f3 - f0) + 1.5 * (f1 - f2);
}

ARB_HD void emit(double* p, double v) {
#if defined(__CUDA_ARCH__)
    asm volatile("st.global.L1::no_allocate.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
#else
    *p = v;
#endif
}

// Shared layouts (doubles).  Pitches are chosen so that every 64-bit access of a half-warp hits 16
// distinct 8-byte bank pairs:
//   grid tile  g[row = z*GY + y][12]            (TMA box, 8 cells + 3 halo points, padded to 16 bytes)
//   X[row][i][cell x], row pitch XP = 36         two rows 2 apart differ by 72 == 8 (mod 16)
//   Y[z][cell y][cell x][j*4 + i], cell-x pitch YCX = 18   (2*cx + i distinct over a half-warp)
constexpr int GX = 12, XP = 36, YCX = 18, YROW = 8 * YCX;
struct alignas(16) Pair { double x, y; };

// x pass over `nrow` rows of a tile plane set: task = (row, cell x), 4 coefficients each.  Rows are
// visited in the order 0,2,1,3 within groups of four so that a half-warp holds rows two apart.
ARB_HD void pass_x(const double* g, double* X, int nrow, int tid, int nthr) {
    const int ntask = ((nrow + 3) / 4) * 32;
    for (int e = tid; e < ntask; e += nthr) {
        const int cx = e & 7, rsub = (e >> 3) & 3;
        const int row = (e >> 5) * 4 + (((rsub & 1) << 1) | (rsub >> 1));
        if (row >= nrow) continue;
        const double* s = g + row * GX + cx;
        double a0, a1, a2, a3;
        cr_line(s[0], s[1], s[2], s[3], a0, a1, a2, a3);
        double* d = X + row * XP + cx;
        d[0] = a0; d[8] = a1; d[16] = a2; d[24] = a3;
    }
}

// y pass: task = (z, cell y, i, cell x) -> 4 coefficients j.  GY = TY + 3 rows per z.
ARB_HD void pass_y(const double* X, double* Y, int nz, int TY, int tid, int nthr) {
    const int GY = TY + 3;
    const int ntask = nz * TY * 32;
    for (int e = tid; e < ntask; e += nthr) {
        const int cx = e & 7, i = (e >> 3) & 3, grp = e >> 5;
        const int cy = grp % TY, z = grp / TY;
        const double* s = X + (z * GY + cy) * XP + i * 8 + cx;
        double a0, a1, a2, a3;
        cr_line(s[0], s[XP], s[2 * XP], s[3 * XP], a0, a1, a2, a3);
        double* d = Y + grp * YROW + cx * YCX + i;
        d[0] = a0; d[4] = a1; d[8] = a2; d[12] = a3;
    }
}

// ---------------------------------------------------------------------------------------------------
// 3-D: static tile of 8 x TY x TZ cells
// ---------------------------------------------------------------------------------------------------
template <int TY_, int TZ_, int THREADS_>
struct Sep3 {
    static constexpr int TY = TY_, TZ = TZ_, THREADS = THREADS_;
    static constexpr int GY = TY + 3, GZ = TZ + 3, NROW = GY * GZ;
    static constexpr int G_ELEMS = NROW * GX, X_ELEMS = NROW * XP, Y_ELEMS = GZ * TY * YROW;
    static constexpr size_t SMEM = (size_t)(G_ELEMS + X_ELEMS + Y_ELEMS) * 8 + 128;

    // z pass: task = (cell, j*4 + i) -> the 4 coefficients k, written to the cell-major table
    ARB_HD static void pass_z_emit(const double* Y, const SepParams& p, int x0, int y0, int z0, int comp, int tid,
                                   int nthr) {
        constexpr int ntask = 8 * TY * TZ * 16;
        for (int e = tid; e < ntask; e += nthr) {
            const int ji = e & 15, cx = (e >> 4) & 7, cyz = e >> 7;
            const int cy = cyz % TY, cz = cyz / TY;
            const int64_t gx = x0 + cx, gy = y0 + cy, gz = z0 + cz;
            if (gx >= p.nc[0] || gy >= p.nc[1] || gz >= p.nc[2]) continue;
            const double* s = Y + (cz * TY + cy) * YROW + cx * YCX + ji;
            double a0, a1, a2, a3;
            cr_line(s[0], s[TY * YROW], s[2 * TY * YROW], s[3 * TY * YROW], a0, a1, a2, a3);
            const int64_t cell = gx + p.nc[0] * (gy + p.nc[1] * gz);
            double* d = p.table + (cell * p.ncomp + comp) * 64 + ji;
            emit(d, a0); emit(d + 16, a1); emit(d + 32, a2); emit(d + 48, a3);
        }
    }
};

// ---------------------------------------------------------------------------------------------------
// 4-D: 8 x 2 x 2 cells per t-layer, marching along t
// ---------------------------------------------------------------------------------------------------
struct Sep4 {
    static constexpr int TY = 2, TZ = 2, THREADS = 256;
    static constexpr int GY = 5, GZ = 5, NROW = 25, PLANE = NROW * GX;
    static constexpr int PLANE_PITCH = 304;                    // 128-byte multiple: TMA destination alignment
    static constexpr int NCELL = 8 * TY * TZ;                  // cells per layer
    static constexpr int NTASK_E = NCELL * 16;                 // emit tasks per layer: (cell, j*4 + i)
    static constexpr int X_ELEMS = NROW * XP, Y_ELEMS = GZ * TY * YROW;
    static constexpr int W3 = 81;                              // corner points of a plane: 9 x 3 x 3
    static constexpr int W3_PITCH = 82;
    static constexpr int RING_SLOT = NCELL * 64;
    static constexpr int G_ELEMS = 2 * W3_PITCH;               // fxyzt at the corner points of a layer: [ct][point]
    // offsets into dynamic shared memory (doubles); Y and the corner values are double-buffered (software pipeline)
    static constexpr int OFF_PLANE = 0, OFF_X = 2 * PLANE_PITCH, OFF_Y = OFF_X + X_ELEMS, OFF_RING = OFF_Y + 2 * Y_ELEMS,
                         OFF_W3 = OFF_RING + 3 * RING_SLOT, OFF_G = OFF_W3 + 4 * W3_PITCH, TOTAL = OFF_G + 2 * G_ELEMS;
    static constexpr size_t SMEM = (size_t)TOTAL * 8 + 128;

    // phase A: x pass of a new plane; with the quirk also dxdydz (central differences, unit spacing) at the
    // plane's 9 x 3 x 3 corner points (corner p <-> tile grid coordinate p + 1) -> this plane's w3 slot
    template <bool QUIRK>
    ARB_HD static void phase_a(const double* plane, double* X, double* w3, int tid, int nthr) {
        pass_x(plane, X, NROW, tid, nthr);
        if (QUIRK)
            for (int e = tid; e < W3; e += nthr) {
                const int px = e % 9, py = (e / 9) % 3, pz = e / 27;
                const double* s = plane + (pz * GY + py) * GX + px;
                const double lo = (s[2 * GX + 2] - s[2 * GX]) - (s[2] - s[0]);                  // z = pz     (x4)
                const double* u = s + 2 * GY * GX;
                const double hi = (u[2 * GX + 2] - u[2 * GX]) - (u[2] - u[0]);                  // z = pz + 2 (x4)
                w3[e] = 0.125 * (hi - lo);
            }
    }
    // phase B: y pass of local plane s; with the quirk (and s >= 3) also fxyzt at the corner points of the layer
    // that plane s completes: its corners sit on the local planes s-2 (ct = 0) and s-1 (ct = 1), and
    // fxyzt(plane q) = 0.5 * (w3[q + 1] - w3[q - 1]).
    template <bool QUIRK>
    ARB_HD static void phase_b(const double* X, double* Y, const double* w3ring, double* g, int s, int tid, int nthr) {
        pass_y(X, Y, GZ, TY, tid, nthr);
        if (QUIRK && s >= 3)
            for (int e = tid; e < 2 * W3; e += nthr) {
                const int ct = e >= W3, pt = e - ct * W3;
                const int q = s - 2 + ct;
                g[ct * W3_PITCH + pt] = 0.5 * (w3ring[((q + 1) & 3) * W3_PITCH + pt] - w3ring[((q - 1) & 3) * W3_PITCH + pt]);
            }
    }

    // Step-invariant part of one emit task (cell, j*4 + i); a thread owns the same tasks in every step.
    struct ETask {
        int y_off, ring_off, g_off;   // into a Y buffer, a ring slot, a corner-value plane
        int64_t out_off;              // table offset (doubles) of the task's first output in layer 0 of the march; -1: not stored
        double p00, p01, p10, p11;    // Hq[j][cy] Hq[i][cx]
    };
    ARB_HD static ETask make_task(int e, const SepParams& p, int x0, int y0, int z0, int64_t t0, int comp) {
        ETask t;
        const int ji = e & 15, cell = e >> 4;
        const int cx = cell & 7, cy = (cell >> 3) & 1, cz = cell >> 4;
        t.y_off = (cz * TY + cy) * YROW + cx * YCX + ji;
        t.ring_off = cell * 64 + ji;
        t.g_off = (cz * 3 + cy) * 9 + cx;
        const int64_t gx = x0 + cx, gy = y0 + cy, gz = z0 + cz;
        const bool ok = gx < p.nc[0] && gy < p.nc[1] && gz < p.nc[2];
        t.out_off = ok ? ((gx + p.nc[0] * (gy + p.nc[1] * (gz + p.nc[2] * t0))) * p.ncomp + comp) * 256 + ji : -1;
        const int i = ji & 3, j = ji >> 2;
        const double hi0 = (i == 1 || i == 3) ? 1.0 : (i == 2 ? -2.0 : 0.0);
        const double hi1 = (i == 3) ? 1.0 : (i == 2 ? -1.0 : 0.0);
        const double hj0 = (j == 1 || j == 3) ? 1.0 : (j == 2 ? -2.0 : 0.0);
        const double hj1 = (j == 3) ? 1.0 : (j == 2 ? -1.0 : 0.0);
        t.p00 = hj0 * hi0; t.p01 = hj0 * hi1; t.p10 = hj1 * hi0; t.p11 = hj1 * hi1;
        return t;
    }
    // phase E, one task: z pass of the new plane fused with the t pass of the layer it completes.  The 4 fresh
    // z-pass values (k = 0..3) replace the oldest ring plane in place after the thread has read it, so the ring
    // needs three slots and no barrier of its own.  layer_stride = doubles between consecutive t layers of the table.
    template <bool QUIRK>
    ARB_HD static void emit_task(const ETask& t, const double* Y, double* ring, const double* g, double* table,
                                 int64_t layer_stride, int s) {
        const double* sy = Y + t.y_off;
        double fresh[4];
        cr_line(sy[0], sy[TY * YROW], sy[2 * TY * YROW], sy[3 * TY * YROW], fresh[0], fresh[1], fresh[2], fresh[3]);
        const int slot_new = s % 3;                                   // holds local plane s-3, receives plane s
        double* r0 = ring + slot_new * RING_SLOT + t.ring_off;
        if (s >= 3 && t.out_off >= 0) {
            const double* r1 = ring + ((s + 1) % 3) * RING_SLOT + t.ring_off;       // plane s-2
            const double* r2 = ring + ((s + 2) % 3) * RING_SLOT + t.ring_off;       // plane s-1
            double* d = table + t.out_off + (int64_t)(s - 3) * layer_stride;
            double G0[4] = {0.0, 0.0, 0.0, 0.0}, G1[4] = {0.0, 0.0, 0.0, 0.0};
            if (QUIRK) {
                // e[c] = fxyzt(corner c-1) - fxyzt(corner c);  F[ct][cz] = sum_{cy,cx} Hq[j][cy] Hq[i][cx] e[ct][cz][cy][cx]
                const double* gc = g + t.g_off;
                double F[4];
                double prev = 0.0;
                ARB_SEP_UNROLL
                for (int m = 0; m < 4; ++m) {                                    // m = 2 ct + cz
                    const double* gm = gc + (m >> 1) * W3_PITCH + (m & 1) * 27;
                    const double g0 = gm[0], g1 = gm[1], g2 = gm[9], g3 = gm[10];
                    F[m] = (t.p00 * (prev - g0) + t.p01 * (g0 - g1)) + (t.p10 * (g1 - g2) + t.p11 * (g2 - g3));
                    prev = g3;
                }
                // G[ct][k] = Hq[k][0] F[ct][0] + Hq[k][1] F[ct][1]
                G0[1] = F[0]; G0[2] = -2.0 * F[0] - F[1]; G0[3] = F[0] + F[1];
                G1[1] = F[2]; G1[2] = -2.0 * F[2] - F[3]; G1[3] = F[2] + F[3];
            }
            ARB_SEP_UNROLL
            for (int k = 0; k < 4; ++k) {
                double a0, a1, a2, a3;
                cr_line(r0[k * 16], r1[k * 16], r2[k * 16], fresh[k], a0, a1, a2, a3);
                if (QUIRK) {   // + Hq[l][0] G[0][k] + Hq[l][1] G[1][k]
                    a1 += G0[k];
                    a2 += -2.0 * G0[k] - G1[k];
                    a3 += G0[k] + G1[k];
                }
                emit(d + k * 16, a0); emit(d + 64 + k * 16, a1); emit(d + 128 + k * 16, a2);
                emit(d + 192 + k * 16, a3);
            }
        }
        r0[0] = fresh[0]; r0[16] = fresh[1]; r0[32] = fresh[2]; r0[48] = fresh[3];
    }
};

}  // namespace sep
}  // namespace arb
