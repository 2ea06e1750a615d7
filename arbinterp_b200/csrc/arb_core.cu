// Error plumbing, version string and the exact constant matrices of the Lekien-Marsden scheme
// (reference: makeAMatrix, A.py:107-175 and A.py:726-878).  The reference builds B by evaluating
// monomial derivatives at the cube corners and inverts it with LAPACK; here inv(B) is written
// down directly as a Kronecker product of the 1-D cubic Hermite inverse, which is exact.
#include <stdarg.h>
#include <vector>
#include "arb_common.cuh"

namespace arb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return 0;
    set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    return 1000 + (int)e;
}

int num_sms() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (!cached[dev]) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

// Subsets of {0..d-1} ordered by size then lexicographically: f, fx, fy, fz, [ft,] fxy, ...
int deriv_mask(int d, int r) {
    int idx = 0;
    for (int size = 0; size <= d; ++size) {
        // lexicographic combinations of `size` axes
        int comb[4] = {0, 1, 2, 3};
        if (size == 0) {
            if (idx == r) return 0;
            ++idx;
            continue;
        }
        while (true) {
            if (idx == r) {
                int m = 0;
                for (int i = 0; i < size; ++i) m |= 1 << comb[i];
                return m;
            }
            ++idx;
            int i = size - 1;
            while (i >= 0 && comb[i] == d - size + i) --i;
            if (i < 0) break;
            ++comb[i];
            for (int j = i + 1; j < size; ++j) comb[j] = comb[j - 1] + 1;
        }
    }
    return -1;
}

// 1-D cubic Hermite inverse: coefficients a0..a3 of p(x) = sum a_e x^e from [p(0), p(1), p'(0), p'(1)].
static const double H1[4][4] = {{1, 0, 0, 0}, {0, 0, 1, 0}, {-3, 3, -2, -1}, {2, -2, 1, 1}};

void make_invB(int d, double* out) {
    const int ncorner = 1 << d, nm = 1 << (2 * d);
    for (int m = 0; m < nm; ++m)
        for (int r = 0; r < ncorner; ++r) {
            const int mask = deriv_mask(d, r);
            for (int c = 0; c < ncorner; ++c) {
                double v = 1.0;
                for (int a = 0; a < d; ++a) {
                    const int e = (m >> (2 * a)) & 3;
                    const int s = (((mask >> a) & 1) ? 2 : 0) + ((c >> a) & 1);
                    v *= H1[e][s];
                }
                out[(size_t)m * nm + r * ncorner + c] = v;
            }
        }
}

void make_D(int d, int quirk, double* out) {
    const int ncorner = 1 << d, nm = 1 << (2 * d);
    memset(out, 0, sizeof(double) * nm * nm);
    int stride[4] = {1, 4, 16, 64};
    int base = 0;
    for (int a = 0; a < d; ++a) base += stride[a];
    for (int r = 0; r < ncorner; ++r) {
        const int mask = deriv_mask(d, r);
        int axes[4], na = 0;
        for (int a = 0; a < d; ++a)
            if ((mask >> a) & 1) axes[na++] = a;
        double w = 1.0;
        for (int i = 0; i < na; ++i) w *= 0.5;
        for (int c = 0; c < ncorner; ++c) {
            int cc = c;
            if (d == 4 && na == 4 && quirk) {  // A.py:860: rows start at 241, row 240 stays zero
                if (c == 0) continue;
                cc = c - 1;
            }
            int centre = base;
            for (int a = 0; a < d; ++a) centre += ((cc >> a) & 1) * stride[a];
            double* row = out + (size_t)(r * ncorner + c) * nm;
            for (int sgn = 0; sgn < (1 << na); ++sgn) {
                int off = 0;
                double s = w;
                for (int i = 0; i < na; ++i) {
                    if ((sgn >> i) & 1) off += stride[axes[i]];
                    else { off -= stride[axes[i]]; s = -s; }
                }
                row[centre + off] = s;
            }
        }
    }
}

void make_A(int d, int quirk, double* out) {
    const int nm = 1 << (2 * d);
    std::vector<double> ib((size_t)nm * nm), D((size_t)nm * nm);
    make_invB(d, ib.data());
    make_D(d, quirk, D.data());
    // every partial sum is an integer multiple of 2^-d well inside 2^53: exact in any order
    for (int i = 0; i < nm; ++i)
        for (int j = 0; j < nm; ++j) {
            double s = 0.0;
            for (int k = 0; k < nm; ++k) s += ib[(size_t)i * nm + k] * D[(size_t)k * nm + j];
            out[(size_t)i * nm + j] = s;
        }
}

}  // namespace arb

extern "C" {

const char* arb_version(void) { return "arbinterp_b200 0.1.0 (sm_100a; ARBInterp 1.8 semantics)"; }
const char* arb_last_error(void) { return arb::g_err; }

int arb_get_matrix(int d, int which, int reference_quirk, double* out_host) {
    if ((d != 3 && d != 4) || !out_host || which < 0 || which > 2) {
        arb::set_error("arb_get_matrix: bad arguments (d=%d which=%d)", d, which);
        return 1;
    }
    if (which == 0) arb::make_invB(d, out_host);
    else if (which == 1) arb::make_D(d, reference_quirk, out_host);
    else arb::make_A(d, reference_quirk, out_host);
    return 0;
}
}
