// CPU emulation of the node (Hermite) table math (test infrastructure): the __host__ __device__ pieces of
// arbinterp_b200/csrc/arb_nodes.cuh -- node_stencil (the build), eval3, eval4_lane (the query, combined over the
// four (cz, ct) lanes exactly as query_block_kernel<KIND = nodes> combines them) -- against the monomial evaluation
// of alpha = A f, A = inv(B) D from arb_core.cu (A.py:107-175, 726-878), with and without the A.py:860 quirk.
// Exit code 1 above 1e-12 scaled error.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../arbinterp_b200/csrc/arb_common.cuh"
#include "../../arbinterp_b200/csrc/arb_nodes.cuh"

using namespace arb;
using namespace arb::nodes;

static double rnd(uint64_t& s) {
    s = s * 6364136223846793005ULL + 1442695040888963407ULL;
    return (double)(s >> 11) / 9007199254740992.0;
}

template <int D, bool QUIRK>
static double run(const int64_t (&n)[4], int nq, uint64_t seed) {
    constexpr int NM = (D == 3) ? 64 : 256, T = (D == 3) ? 8 : 16;
    const int64_t nt = (D == 4) ? n[3] : 1;
    std::vector<double> grid((size_t)n[0] * n[1] * n[2] * nt);
    for (double& x : grid) x = 2.0 * rnd(seed) - 1.0;
    auto at = [&](int64_t x, int64_t y, int64_t z, int64_t t) { return grid[((t * n[2] + z) * n[1] + y) * n[0] + x]; };
    // ---- build: nodes 1..n-2 per axis
    const int64_t m0 = n[0] - 2, m1 = n[1] - 2, m2 = n[2] - 2, m3 = (D == 4) ? n[3] - 2 : 1;
    std::vector<double> tab((size_t)m0 * m1 * m2 * m3 * T);
    for (int64_t t = 0; t < m3; ++t)
        for (int64_t z = 0; z < m2; ++z)
            for (int64_t y = 0; y < m1; ++y)
                for (int64_t x = 0; x < m0; ++x) {
                    auto get = [&](int dx, int dy, int dz, int dt) {
                        return at(x + 1 + dx, y + 1 + dy, z + 1 + dz, (D == 4) ? t + 1 + dt : 0);
                    };
                    node_stencil<D>(get, &tab[((((size_t)t * m2 + z) * m1 + y) * m0 + x) * T]);
                }
    std::vector<double> A((size_t)NM * NM);
    make_A(D, QUIRK ? 1 : 0, A.data());
    double worst = 0.0;
    for (int qn = 0; qn < nq; ++qn) {
        int64_t c[4] = {0, 0, 0, 0};
        double fr[4] = {0, 0, 0, 0};
        for (int a = 0; a < D; ++a) { c[a] = (int64_t)(rnd(seed) * (n[a] - 3)); fr[a] = rnd(seed); }
        if (qn % 7 == 0) fr[qn % D] = 0.0;
        // ---- reference: alpha = A f on the 4^d neighbourhood, monomial evaluation
        std::vector<double> f(NM), alpha(NM);
        for (int m = 0; m < NM; ++m)
            f[m] = at(c[0] + (m & 3), c[1] + ((m >> 2) & 3), c[2] + ((m >> 4) & 3), (D == 4) ? c[3] + (m >> 6) : 0);
        for (int m = 0; m < NM; ++m) {
            double s = 0.0;
            for (int k = 0; k < NM; ++k) s += A[(size_t)m * NM + k] * f[k];
            alpha[m] = s;
        }
        double ref[5] = {0, 0, 0, 0, 0}, mag[5] = {0, 0, 0, 0, 0};
        for (int m = 0; m < NM; ++m) {
            const int e[4] = {m & 3, (m >> 2) & 3, (m >> 4) & 3, m >> 6};
            double pw[4] = {1, 1, 1, 1}, dpw[4] = {0, 0, 0, 0};
            for (int a = 0; a < D; ++a) { pw[a] = std::pow(fr[a], e[a]); dpw[a] = e[a] ? e[a] * std::pow(fr[a], e[a] - 1) : 0.0; }
            const double terms[5] = {pw[0] * pw[1] * pw[2] * pw[3], dpw[0] * pw[1] * pw[2] * pw[3], pw[0] * dpw[1] * pw[2] * pw[3],
                                     pw[0] * pw[1] * dpw[2] * pw[3], pw[0] * pw[1] * pw[2] * dpw[3]};
            for (int i = 0; i <= D; ++i) { ref[i] += alpha[m] * terms[i]; mag[i] += std::fabs(alpha[m] * terms[i]); }
        }
        // ---- the kernel's way
        double got[5] = {0, 0, 0, 0, 0};
        auto node = [&](int64_t x, int64_t y, int64_t z, int64_t t) {
            return &tab[((((size_t)t * m2 + z) * m1 + y) * m0 + x) * T];
        };
        if (D == 3) {
            alignas(16) double slot[64];
            for (int cz = 0; cz < 2; ++cz)
                for (int cy = 0; cy < 2; ++cy)
                    memcpy(slot + (cz * 2 + cy) * 16, node(c[0], c[1] + cy, c[2] + cz, 0), 16 * sizeof(double));   // x pair is contiguous
            eval3<true>(slot, fr, got);
        } else {
            alignas(16) double slot[4][64];
            for (int sl = 0; sl < 4; ++sl)
                for (int cy = 0; cy < 2; ++cy)
                    memcpy(slot[sl] + cy * 32, node(c[0], c[1] + cy, c[2] + (sl & 1), c[3] + (sl >> 1)), 32 * sizeof(double));
            for (int sl = 0; sl < 4; ++sl) {
                double g[5];
                const double prev = sl ? slot[sl - 1][48 + 15] : 0.0;            // the kernel's __shfl_up
                eval4_lane<true, QUIRK>(slot[sl], sl & 1, sl >> 1, fr, prev, g);
                for (int i = 0; i < 5; ++i) got[i] += g[i];
            }
        }
        for (int i = 0; i <= D; ++i) {
            const double err = std::fabs(got[i] - ref[i]) / std::fmax(mag[i], 1.0);
            if (!(err <= worst)) worst = err;
        }
    }
    return worst;
}

// 3-D interleaved-component forms: nodes::eval3_row (four (cy, cz) lanes) and gridfree::plane_il (four z-plane lanes)
// on four random component grids, against alpha = A f per component.
static double run_interleaved(const int64_t (&n)[3], int nq, uint64_t seed, bool grid_form) {
    const int64_t npt = n[0] * n[1] * n[2];
    std::vector<double> grid((size_t)npt * 4);
    for (double& x : grid) x = 2.0 * rnd(seed) - 1.0;
    auto at = [&](int c, int64_t x, int64_t y, int64_t z) { return grid[(size_t)c * npt + (z * n[1] + y) * n[0] + x]; };
    const int64_t m0 = n[0] - 2, m1 = n[1] - 2, m2 = n[2] - 2;
    std::vector<double> tab((size_t)m0 * m1 * m2 * 32);                   // [node][4][8]
    for (int64_t z = 0; z < m2; ++z)
        for (int64_t y = 0; y < m1; ++y)
            for (int64_t x = 0; x < m0; ++x)
                for (int c = 0; c < 4; ++c) {
                    auto get = [&](int dx, int dy, int dz, int) { return at(c, x + 1 + dx, y + 1 + dy, z + 1 + dz); };
                    node_stencil<3>(get, &tab[(((size_t)z * m1 + y) * m0 + x) * 32 + c * 8]);
                }
    std::vector<double> A(64 * 64);
    make_A(3, 0, A.data());
    double worst = 0.0;
    for (int qn = 0; qn < nq; ++qn) {
        int64_t c0[3];
        double fr[3];
        for (int a = 0; a < 3; ++a) { c0[a] = (int64_t)(rnd(seed) * (n[a] - 3)); fr[a] = rnd(seed); }
        double got[7] = {0, 0, 0, 0, 0, 0, 0};
        for (int sl = 0; sl < 4; ++sl) {
            alignas(16) double slot[64];
            double part[7];
            if (grid_form) {
                for (int j = 0; j < 4; ++j)
                    for (int i = 0; i < 4; ++i)
                        for (int c = 0; c < 4; ++c) slot[(j * 4 + i) * 4 + c] = at(c, c0[0] + i, c0[1] + j, c0[2] + sl);
                gridfree::plane_il<true>(slot, sl, fr, part);
            } else {
                memcpy(slot, &tab[((((size_t)c0[2] + (sl >> 1)) * m1 + c0[1] + (sl & 1)) * m0 + c0[0]) * 32], 64 * sizeof(double));
                eval3_row<true>(slot, sl & 1, sl >> 1, fr, part);
            }
            for (int i = 0; i < 7; ++i) got[i] += part[i];
        }
        for (int c = 0; c < 4; ++c) {
            double f[64], ref[4] = {0, 0, 0, 0}, mag[4] = {0, 0, 0, 0};
            for (int m = 0; m < 64; ++m) f[m] = at(c, c0[0] + (m & 3), c0[1] + ((m >> 2) & 3), c0[2] + (m >> 4));
            for (int m = 0; m < 64; ++m) {
                double al = 0.0;
                for (int k = 0; k < 64; ++k) al += A[m * 64 + k] * f[k];
                const int e[3] = {m & 3, (m >> 2) & 3, m >> 4};
                double pw[3], dpw[3];
                for (int a = 0; a < 3; ++a) { pw[a] = std::pow(fr[a], e[a]); dpw[a] = e[a] ? e[a] * std::pow(fr[a], e[a] - 1) : 0.0; }
                const double t[4] = {pw[0] * pw[1] * pw[2], dpw[0] * pw[1] * pw[2], pw[0] * dpw[1] * pw[2], pw[0] * pw[1] * dpw[2]};
                for (int i = 0; i < 4; ++i) { ref[i] += al * t[i]; mag[i] += std::fabs(al * t[i]); }
            }
            const int nout = (c == 3) ? 4 : 1;
            for (int i = 0; i < nout; ++i) {
                const double g = (i == 0) ? got[c] : got[3 + i];
                const double err = std::fabs(g - ref[i]) / std::fmax(mag[i], 1.0);
                if (!(err <= worst)) worst = err;
            }
        }
    }
    return worst;
}

int main() {
    int bad = 0;
    auto report = [&](const char* name, double e) {
        printf("%s: max scaled error %.3e\n", name, e);
        if (!(e <= 1e-12)) bad = 1;
    };
    report("3d nodes 9x8x7", run<3, false>({9, 8, 7, 1}, 500, 21));
    report("3d nodes 4x4x4", run<3, false>({4, 4, 4, 1}, 50, 22));
    report("3d nodes 12x5x6", run<3, false>({12, 5, 6, 1}, 500, 23));
    report("4d nodes 9x8x7x6 quirk", run<4, true>({9, 8, 7, 6}, 400, 24));
    report("4d nodes 5x6x4x7 quirk", run<4, true>({5, 6, 4, 7}, 400, 25));
    report("4d nodes 4x4x4x4 quirk", run<4, true>({4, 4, 4, 4}, 50, 26));
    report("4d nodes 9x8x7x6 fixed", run<4, false>({9, 8, 7, 6}, 400, 27));
    report("3d interleaved nodes 9x8x7", run_interleaved({9, 8, 7}, 400, 28, false));
    report("3d interleaved nodes 4x4x4", run_interleaved({4, 4, 4}, 50, 29, false));
    report("3d interleaved grid 9x8x7", run_interleaved({9, 8, 7}, 400, 30, true));
    report("3d interleaved grid 5x4x6", run_interleaved({5, 4, 6}, 200, 31, true));
    return bad;
}
