// CPU emulation of the 4-D table-free query math (test infrastructure): the __host__ __device__ pieces of
// arbinterp_b200/csrc/arb_gridfree.cuh are combined exactly as query_grid4_kernel combines them over its
// four lanes (shuffles replaced by array reads) and compared with sum_m alpha_m u^i v^j w^k s^l, alpha = A f,
// A = inv(B) D from arb_core.cu with and without the A.py:860 quirk.  Exit code 1 above 1e-12 scaled error.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../arbinterp_b200/csrc/arb_common.cuh"
#include "../../arbinterp_b200/csrc/arb_gridfree.cuh"

using namespace arb;
using namespace arb::gridfree;

static double rnd(uint64_t& s) {
    s = s * 6364136223846793005ULL + 1442695040888963407ULL;
    return (double)(s >> 11) / 9007199254740992.0;
}

template <bool QUIRK>
static double run(int64_t nx, int64_t ny, int64_t nz, int64_t nt, int nq, uint64_t seed) {
    std::vector<double> grid((size_t)nx * ny * nz * nt);
    for (double& x : grid) x = 2.0 * rnd(seed) - 1.0;
    auto at = [&](int64_t x, int64_t y, int64_t z, int64_t t) {
        if (x < 0 || y < 0 || z < 0 || t < 0 || x >= nx || y >= ny || z >= nz || t >= nt) return 0.0;
        return grid[((t * nz + z) * ny + y) * nx + x];
    };
    std::vector<double> A(256 * 256);
    make_A(4, QUIRK ? 1 : 0, A.data());
    double worst = 0.0;
    for (int qn = 0; qn < nq; ++qn) {
        const int64_t ix = (int64_t)(rnd(seed) * (nx - 3)), iy = (int64_t)(rnd(seed) * (ny - 3)),
                      iz = (int64_t)(rnd(seed) * (nz - 3)), it = (int64_t)(rnd(seed) * (nt - 3));
        const double fr[4] = {rnd(seed), rnd(seed), rnd(seed), rnd(seed)};
        // ---- reference: alpha = A f, evaluated with monomials
        double f[256], alpha[256];
        for (int m = 0; m < 256; ++m) f[m] = at(ix + (m & 3), iy + ((m >> 2) & 3), iz + ((m >> 4) & 3), it + (m >> 6));
        for (int m = 0; m < 256; ++m) {
            double s = 0.0;
            for (int k = 0; k < 256; ++k) s += A[m * 256 + k] * f[k];
            alpha[m] = s;
        }
        double ref[5] = {0, 0, 0, 0, 0}, mag[5] = {0, 0, 0, 0, 0};
        for (int m = 0; m < 256; ++m) {
            const int e[4] = {m & 3, (m >> 2) & 3, (m >> 4) & 3, m >> 6};
            double pw[4], dpw[4];
            for (int a = 0; a < 4; ++a) { pw[a] = std::pow(fr[a], e[a]); dpw[a] = e[a] ? e[a] * std::pow(fr[a], e[a] - 1) : 0.0; }
            const double terms[5] = {pw[0] * pw[1] * pw[2] * pw[3], dpw[0] * pw[1] * pw[2] * pw[3], pw[0] * dpw[1] * pw[2] * pw[3],
                                     pw[0] * pw[1] * dpw[2] * pw[3], pw[0] * pw[1] * pw[2] * dpw[3]};
            for (int i = 0; i < 5; ++i) { ref[i] += alpha[m] * terms[i]; mag[i] += std::fabs(alpha[m] * terms[i]); }
        }
        // ---- the kernel's way: four "lanes", one grid plane each
        const int off = (int)(ix & 1);
        double wx[4], dwx[4], wy[4], dwy[4], wz[4], dwz[4], wt[4], dwt[4];
        catmull_rom(fr[0], wx, dwx); catmull_rom(fr[1], wy, dwy); catmull_rom(fr[2], wz, dwz); catmull_rom(fr[3], wt, dwt);
        PlanePartial pp[4];
        for (int l = 0; l < 4; ++l) {
            alignas(16) double box[16 * 6];
            for (int k = 0; k < 4; ++k)
                for (int j = 0; j < 4; ++j)
                    for (int b = 0; b < 6; ++b) box[(4 * k + j) * 6 + b] = at(ix - off + b, iy + j, iz + k, it + l);
            const int lane = (qn * 4 + l) & 31;                     // any rotation must give the same answer
            memset(&pp[l], 0, sizeof(PlanePartial));
            plane_partial<true, QUIRK>(reinterpret_cast<const unsigned char*>(box), off, lane & 3, (lane >> 2) & 1, wx, dwx,
                                       wy, dwy, wz, dwz, pp[l]);
        }
        double got[5] = {0, 0, 0, 0, 0};
        for (int l = 0; l < 4; ++l) {
            double m[5] = {pp[l].val * wt[l], pp[l].gx * wt[l], pp[l].gy * wt[l], pp[l].gz * wt[l], pp[l].val * dwt[l]};
            if (QUIRK) {
                double g[8], g_otherct[8];
                for (int i = 0; i < 8; ++i) {
                    const double other = pp[l ^ 2].S[i];
                    g[i] = 0.0625 * ((l >= 2) ? (pp[l].S[i] - other) : (other - pp[l].S[i]));
                    const int lo = (l ^ 1) & 1;                    // what lane l ^ 1 computes for its ct
                    g_otherct[i] = 0.0625 * (pp[lo + 2].S[i] - pp[lo].S[i]);
                }
                const int ct = l & 1;
                double hx[2], dhx[2], hy[2], dhy[2], hz[2], dhz[2], ht[2], dht[2], c[4];
                hermite_slope(fr[0], hx, dhx); hermite_slope(fr[1], hy, dhy); hermite_slope(fr[2], hz, dhz);
                hermite_slope(fr[3], ht, dht);
                corner_term<true>(g, ct ? g_otherct[7] : 0.0, hx, dhx, hy, dhy, hz, dhz, c);
                if (l < 2) {
                    m[0] += ht[ct] * c[0]; m[1] += ht[ct] * c[1]; m[2] += ht[ct] * c[2]; m[3] += ht[ct] * c[3];
                    m[4] += dht[ct] * c[0];
                }
            }
            for (int i = 0; i < 5; ++i) got[i] += m[i];
        }
        for (int i = 0; i < 5; ++i) {
            const double err = std::fabs(got[i] - ref[i]) / std::fmax(mag[i], 1.0);
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
    report("4d table-free 9x8x7x6 quirk", run<true>(9, 8, 7, 6, 400, 11));
    report("4d table-free 10x5x6x7 quirk", run<true>(10, 5, 6, 7, 400, 12));
    report("4d table-free 4x4x4x4 quirk", run<true>(4, 4, 4, 4, 50, 13));
    report("4d table-free 9x8x7x6 fixed", run<false>(9, 8, 7, 6, 400, 14));
    return bad;
}
