// CPU emulation of the 4-D table-free query math on the component-interleaved grid (test infrastructure): the
// __host__ __device__ pieces of arbinterp_b200/csrc/arb_gridil4.cuh are combined exactly as query_gridil4_kernel
// combines them -- four "lanes" (z-planes) per query, four passes (t-planes), the lanes' shares added up -- and compared, component by component, with sum_m alpha_m u^i v^j w^k s^l,
// alpha = A f, A = inv(B) D from arb_core.cu (A.py:726-878) with and without the A.py:860 quirk.
// Exit code 1 above 1e-12 scaled error.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../arbinterp_b200/csrc/arb_common.cuh"
#include "../../arbinterp_b200/csrc/arb_gridil4.cuh"

using namespace arb;
using namespace arb::gridil4;

static double rnd(uint64_t& s) {
    s = s * 6364136223846793005ULL + 1442695040888963407ULL;
    return (double)(s >> 11) / 9007199254740992.0;
}

template <bool BOTH, bool QUIRK>
static double run(int64_t nx, int64_t ny, int64_t nz, int64_t nt, int nq, uint64_t seed) {
    constexpr int NC = BOTH ? 4 : 3;
    std::vector<double> grid((size_t)nx * ny * nz * nt * 4);                 // [t][z][y][x][c]
    for (double& x : grid) x = 2.0 * rnd(seed) - 1.0;
    auto at = [&](int64_t x, int64_t y, int64_t z, int64_t t) { return &grid[(((t * nz + z) * ny + y) * nx + x) * 4]; };
    std::vector<double> A(256 * 256);
    make_A(4, QUIRK ? 1 : 0, A.data());
    double worst = 0.0;
    for (int qn = 0; qn < nq; ++qn) {
        const int64_t ix = (int64_t)(rnd(seed) * (nx - 3)), iy = (int64_t)(rnd(seed) * (ny - 3)),
                      iz = (int64_t)(rnd(seed) * (nz - 3)), it = (int64_t)(rnd(seed) * (nt - 3));
        double fr[4] = {rnd(seed), rnd(seed), rnd(seed), rnd(seed)};
        if (qn % 9 == 0) fr[qn % 4] = 0.0;
        // ---- reference per component: alpha = A f, evaluated with monomials
        double ref[4][5], mag[4][5];
        for (int c = 0; c < NC; ++c) {
            double f[256], alpha[256];
            for (int m = 0; m < 256; ++m) f[m] = at(ix + (m & 3), iy + ((m >> 2) & 3), iz + ((m >> 4) & 3), it + (m >> 6))[c];
            for (int m = 0; m < 256; ++m) {
                double s = 0.0;
                for (int k = 0; k < 256; ++k) s += A[m * 256 + k] * f[k];
                alpha[m] = s;
            }
            for (int i = 0; i < 5; ++i) ref[c][i] = mag[c][i] = 0.0;
            for (int m = 0; m < 256; ++m) {
                const int e[4] = {m & 3, (m >> 2) & 3, (m >> 4) & 3, m >> 6};
                double pw[4], dpw[4];
                for (int a = 0; a < 4; ++a) { pw[a] = std::pow(fr[a], e[a]); dpw[a] = e[a] ? e[a] * std::pow(fr[a], e[a] - 1) : 0.0; }
                const double terms[5] = {pw[0] * pw[1] * pw[2] * pw[3], dpw[0] * pw[1] * pw[2] * pw[3], pw[0] * dpw[1] * pw[2] * pw[3],
                                         pw[0] * pw[1] * dpw[2] * pw[3], pw[0] * pw[1] * pw[2] * dpw[3]};
                for (int i = 0; i < 5; ++i) { ref[c][i] += alpha[m] * terms[i]; mag[c][i] += std::fabs(alpha[m] * terms[i]); }
            }
        }
        // ---- the kernel's way
        Weights W;
        make_weights(fr, W);
        double acc[4][8];
        for (int k = 0; k < 4; ++k) {
            for (int i = 0; i < 8; ++i) acc[k][i] = 0.0;
            for (int l = 0; l < 4; ++l) {
                // the kernel's gather: 32 "lanes" bring 16 bytes each, addressed by the kernel's own helpers
                alignas(16) double slot[64];
                const int idx[4] = {(int)ix, (int)iy, (int)iz, (int)it};
                const uint32_t src32 = (uint32_t)plane_first_point(idx, k, l, nx, ny, nz);
                for (int lane = 0; lane < 32; ++lane)
                    memcpy(reinterpret_cast<char*>(slot) + lane * 16,
                           reinterpret_cast<const char*>(grid.data()) + lane_piece_bytes(lane, nx) + ((size_t)src32 << 5), 16);
                for (int j = 0; j < 4; ++j)
                    for (int i = 0; i < 4; ++i)
                        if (memcmp(slot + (j * 4 + i) * 4, at(ix + i, iy + j, iz + k, it + l), 4 * sizeof(double)) != 0) return 1.0;
                pass<BOTH, QUIRK>(acc[k], slot, k, l, fr, W);
            }
        }
        double got[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int k = 0; k < 4; ++k)
            for (int i = 0; i < 8; ++i) got[i] += acc[k][i];
        for (int c = 0; c < 3; ++c) {
            const double err = std::fabs(got[c] - ref[c][0]) / std::fmax(mag[c][0], 1.0);
            if (!(err <= worst)) worst = err;
        }
        if (BOTH)
            for (int i = 0; i < 5; ++i) {
                const double err = std::fabs(got[3 + i] - ref[3][i]) / std::fmax(mag[3][i], 1.0);
                if (!(err <= worst)) worst = err;
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
    report("4d interleaved table-free 9x8x7x6 both quirk", run<true, true>(9, 8, 7, 6, 300, 21));
    report("4d interleaved table-free 10x5x6x7 vector quirk", run<false, true>(10, 5, 6, 7, 300, 22));
    report("4d interleaved table-free 4x4x4x4 both quirk", run<true, true>(4, 4, 4, 4, 50, 23));
    report("4d interleaved table-free 9x8x7x6 both fixed", run<true, false>(9, 8, 7, 6, 300, 24));
    report("4d interleaved table-free 7x6x5x8 vector fixed", run<false, false>(7, 6, 5, 8, 200, 25));
    return bad;
}
