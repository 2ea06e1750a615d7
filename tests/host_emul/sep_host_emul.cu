// CPU emulation of build_sep3_kernel / build_sep4_kernel (test infrastructure).
// Runs the very phases of arbinterp_b200/csrc/arb_build_sep.cuh on the host -- one "thread" after the
// other within a phase, phases in kernel order, the TMA box replaced by a zero-filled copy -- and compares
// every cell's coefficients with the dense product A f[4^d neighbourhood], A = inv(B) D generated exactly by
// arb_core.cu (make_A, including the A.py:860 quirk).  Prints "max scaled error <e>" per case; exit code 1
// when a case exceeds 1e-12.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../arbinterp_b200/csrc/arb_common.cuh"
#include "../../arbinterp_b200/csrc/arb_build_sep.cuh"

using namespace arb;

static double rnd(uint64_t& s) {
    s = s * 6364136223846793005ULL + 1442695040888963407ULL;
    return (double)(s >> 11) / 9007199254740992.0 * 2.0 - 1.0;
}

// grid [C][nt][nz][ny][nx]; zero outside (what the TMA unit fills in)
struct Grid {
    int64_t n[4];
    int ncomp;
    std::vector<double> v;
    double at(int c, int64_t x, int64_t y, int64_t z, int64_t t) const {
        if (x < 0 || y < 0 || z < 0 || t < 0 || x >= n[0] || y >= n[1] || z >= n[2] || t >= n[3]) return 0.0;
        return v[(((c * n[3] + t) * n[2] + z) * n[1] + y) * n[0] + x];
    }
};

static double check(int d, const Grid& g, const std::vector<double>& table, int quirk) {
    const int nm = 1 << (2 * d);
    std::vector<double> A((size_t)nm * nm);
    make_A(d, quirk, A.data());
    int64_t nc[4] = {g.n[0] - 3, g.n[1] - 3, g.n[2] - 3, d == 4 ? g.n[3] - 3 : 1};
    double scale = 0.0, worst = 0.0;
    for (double x : g.v) scale = std::fmax(scale, std::fabs(x));
    std::vector<double> f(nm);
    for (int c = 0; c < g.ncomp; ++c)
        for (int64_t ct = 0; ct < nc[3]; ++ct)
            for (int64_t cz = 0; cz < nc[2]; ++cz)
                for (int64_t cy = 0; cy < nc[1]; ++cy)
                    for (int64_t cx = 0; cx < nc[0]; ++cx) {
                        for (int m = 0; m < nm; ++m) {
                            const int i = m & 3, j = (m >> 2) & 3, k = (m >> 4) & 3, l = (m >> 6) & 3;
                            f[m] = g.at(c, cx + i, cy + j, cz + k, d == 4 ? ct + l : 0);
                        }
                        const int64_t cell = cx + nc[0] * (cy + nc[1] * (cz + nc[2] * ct));
                        const double* got = table.data() + (cell * g.ncomp + c) * nm;
                        for (int m = 0; m < nm; ++m) {
                            double ref = 0.0, mag = 0.0;
                            for (int q = 0; q < nm; ++q) { ref += A[(size_t)m * nm + q] * f[q]; }
                            mag = std::fmax(std::fabs(ref), scale);
                            const double err = std::fabs(got[m] - ref) / mag;
                            if (!(err <= worst)) worst = err;       // NaN-propagating max
                        }
                    }
    return worst;
}

template <typename S3>
static double run3(const Grid& g) {
    sep::SepParams p{};
    int64_t ncell = 1;
    const int tdim[3] = {8, S3::TY, S3::TZ};
    for (int a = 0; a < 3; ++a) { p.nc[a] = g.n[a] - 3; ncell *= p.nc[a]; p.ntile[a] = (p.nc[a] + tdim[a] - 1) / tdim[a]; }
    p.nc[3] = 1; p.ncomp = g.ncomp;
    std::vector<double> table((size_t)ncell * g.ncomp * 64, NAN);
    p.table = table.data();
    std::vector<double> gt(S3::G_ELEMS), X(S3::X_ELEMS, NAN), Y(S3::Y_ELEMS, NAN);
    for (int comp = 0; comp < g.ncomp; ++comp)
        for (int64_t tz = 0; tz < p.ntile[2]; ++tz)
            for (int64_t ty = 0; ty < p.ntile[1]; ++ty)
                for (int64_t tx = 0; tx < p.ntile[0]; ++tx) {
                    const int x0 = (int)tx * 8, y0 = (int)ty * S3::TY, z0 = (int)tz * S3::TZ;
                    for (int z = 0; z < S3::GZ; ++z)
                        for (int y = 0; y < S3::GY; ++y)
                            for (int x = 0; x < sep::GX; ++x)
                                gt[(z * S3::GY + y) * sep::GX + x] = g.at(comp, x0 + x, y0 + y, z0 + z, 0);
                    for (int t = 0; t < S3::THREADS; ++t) sep::pass_x(gt.data(), X.data(), S3::NROW, t, S3::THREADS);
                    for (int t = 0; t < S3::THREADS; ++t) sep::pass_y(X.data(), Y.data(), S3::GZ, S3::TY, t, S3::THREADS);
                    for (int t = 0; t < S3::THREADS; ++t) S3::pass_z_emit(Y.data(), p, x0, y0, z0, comp, t, S3::THREADS);
                }
    return check(3, g, table, 0);
}

// order = +1: threads 0..T-1 within a barrier interval, -1: T-1..0 (an intra-interval dependency between
// threads would make the two orders disagree)
template <bool QUIRK>
static double run4(const Grid& g, int lt, int order = 1) {
    const int quirk = QUIRK ? 1 : 0;
    using S4 = sep::Sep4;
    sep::SepParams p{};
    int64_t ncell = 1;
    const int tdim[3] = {8, S4::TY, S4::TZ};
    for (int a = 0; a < 4; ++a) { p.nc[a] = g.n[a] - 3; ncell *= p.nc[a]; if (a < 3) p.ntile[a] = (p.nc[a] + tdim[a] - 1) / tdim[a]; }
    p.ncomp = g.ncomp; p.quirk = quirk; p.lt = lt > p.nc[3] ? (int)p.nc[3] : lt;
    std::vector<double> table((size_t)ncell * g.ncomp * 256, NAN);
    p.table = table.data();
    const int64_t nchunk = (p.nc[3] + p.lt - 1) / p.lt;
    std::vector<double> sm(S4::TOTAL);
    const int T = S4::THREADS;
    auto each = [&](auto&& fn) {
        if (order > 0) for (int t = 0; t < T; ++t) fn(t);
        else for (int t = T - 1; t >= 0; --t) fn(t);
    };
    for (int comp = 0; comp < g.ncomp; ++comp)
        for (int64_t chunk = 0; chunk < nchunk; ++chunk)
            for (int64_t tz = 0; tz < p.ntile[2]; ++tz)
                for (int64_t ty = 0; ty < p.ntile[1]; ++ty)
                    for (int64_t tx = 0; tx < p.ntile[0]; ++tx) {
                        for (double& x : sm) x = NAN;            // uninitialised shared memory must never reach a result
                        const int x0 = (int)tx * 8, y0 = (int)ty * S4::TY, z0 = (int)tz * S4::TZ;
                        const int64_t t0 = chunk * p.lt;
                        const int nlayer = (int)((p.nc[3] - t0 < p.lt) ? (p.nc[3] - t0) : p.lt);
                        const int nstep = nlayer + 3;
                        double* base = sm.data();
                        double *X = base + S4::OFF_X, *Y = base + S4::OFF_Y, *ring = base + S4::OFF_RING,
                               *w3ring = base + S4::OFF_W3, *gbuf = base + S4::OFF_G;
                        auto fetch = [&](int q) {                 // what the TMA unit delivers for plane q
                            double* plane = base + S4::OFF_PLANE + (q & 1) * S4::PLANE_PITCH;
                            for (int z = 0; z < S4::GZ; ++z)
                                for (int y = 0; y < S4::GY; ++y)
                                    for (int x = 0; x < sep::GX; ++x)
                                        plane[(z * S4::GY + y) * sep::GX + x] = g.at(comp, x0 + x, y0 + y, z0 + z, t0 + q);
                        };
                        const int64_t layer_stride = p.nc[0] * p.nc[1] * p.nc[2] * p.ncomp * 256;
                        auto task = [&](int e) { return S4::make_task(e, p, x0, y0, z0, t0, comp); };
                        // same schedule as build_sep4_kernel
                        fetch(0); fetch(1);
                        each([&](int t) { S4::phase_a<QUIRK>(base + S4::OFF_PLANE, X, w3ring, t, T); });
                        fetch(2);
                        each([&](int t) { S4::phase_b<QUIRK>(X, Y, w3ring, gbuf, 0, t, T); });
                        for (int s = 0; s < nstep; ++s) {
                            const int q = s + 1;
                            const bool more = q < nstep;
                            const double* Ys = Y + (s & 1) * S4::Y_ELEMS;
                            const double* gs = gbuf + (s & 1) * S4::G_ELEMS;
                            each([&](int t) {
                                if (more)
                                    S4::phase_a<QUIRK>(base + S4::OFF_PLANE + (q & 1) * S4::PLANE_PITCH, X,
                                                       w3ring + (q & 3) * S4::W3_PITCH, t, T);
                                S4::emit_task<QUIRK>(task(t), Ys, ring, gs, p.table, layer_stride, s);
                            });
                            if (more && q + 2 < nstep) fetch(q + 2);
                            each([&](int t) {
                                if (more)
                                    S4::phase_b<QUIRK>(X, Y + (q & 1) * S4::Y_ELEMS, w3ring, gbuf + (q & 1) * S4::G_ELEMS, q, t, T);
                                S4::emit_task<QUIRK>(task(t + S4::NTASK_E / 2), Ys, ring, gs, p.table, layer_stride, s);
                            });
                        }
                    }
    return check(4, g, table, quirk);
}

static Grid make_grid(int64_t nx, int64_t ny, int64_t nz, int64_t nt, int ncomp, uint64_t seed) {
    Grid g;
    g.n[0] = nx; g.n[1] = ny; g.n[2] = nz; g.n[3] = nt; g.ncomp = ncomp;
    g.v.resize((size_t)nx * ny * nz * nt * ncomp);
    for (double& x : g.v) x = rnd(seed);
    return g;
}

int main() {
    int bad = 0;
    auto report = [&](const char* name, double e) {
        printf("%s: max scaled error %.3e\n", name, e);
        if (!(e <= 1e-12)) bad = 1;
    };
    {
        Grid g = make_grid(13, 9, 10, 1, 2, 1);
        report("3d 13x9x10 C=2 tile 8x4x4/128", run3<sep::Sep3<4, 4, 128>>(g));
        report("3d 13x9x10 C=2 tile 8x4x8/256", run3<sep::Sep3<4, 8, 256>>(g));
        report("3d 13x9x10 C=2 tile 8x2x4/128", run3<sep::Sep3<2, 4, 128>>(g));
    }
    {
        Grid g = make_grid(4, 4, 4, 1, 1, 2);
        report("3d 4x4x4 C=1 (one cell)", run3<sep::Sep3<4, 4, 128>>(g));
        Grid h = make_grid(20, 4, 7, 1, 3, 3);
        report("3d 20x4x7 C=3", run3<sep::Sep3<4, 4, 256>>(h));
    }
    {
        Grid g = make_grid(12, 6, 7, 9, 2, 4);
        report("4d 12x6x7x9 C=2 quirk lt=all", run4<true>(g, 1 << 20));
        report("4d 12x6x7x9 C=2 quirk lt=4", run4<true>(g, 4));
        report("4d 12x6x7x9 C=2 quirk lt=4, threads in reverse order", run4<true>(g, 4, -1));
        report("4d 12x6x7x9 C=2 quirk lt=1", run4<true>(g, 1));
        report("4d 12x6x7x9 C=2 fixed lt=5", run4<false>(g, 5));
        Grid h = make_grid(4, 4, 4, 4, 1, 5);
        report("4d 4x4x4x4 C=1 (one cell) quirk", run4<true>(h, 3));
        Grid k = make_grid(7, 9, 5, 6, 1, 6);
        report("4d 7x9x5x6 C=1 quirk lt=2", run4<true>(k, 2));
    }
    return bad;
}
