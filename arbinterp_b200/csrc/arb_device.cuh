// Device-side pieces shared by the query kernels (arb_query.cu) and the particle pusher
// (arb_push.cu): kernel parameters, the reference-exact cell location and the nested-Horner
// evaluation of a 64-coefficient tricubic block.
#pragma once
#include "arb_common.cuh"

namespace arb {

struct QueryParams {
    const double* table;
    double* q;
    int64_t N, ldq;
    double* out_comps;
    double* out_norm;
    double* out_grad;
    int64_t* out_cell;
    int64_t* masked_rows;
    unsigned long long* masked_count;
    double mn[4], mx[4], h[4];
    int64_t nc[4];
    int64_t slab_lo, slab_hi;   // owned layers of the slowest axis
    int64_t total_cells;        // prod(nc): sentinel index (A.py:369)
    int64_t layer_cells;        // prod(nc[0..d-2])
    int64_t nn[4];              // node table (arb_nodes.cuh): nodes per axis = nc + 1
    int64_t node_comp_stride;   // doubles between the components of a node table
    // routed results (slab-sharded tables, arb_query_routed): rows [seg_start[h], seg_start[h+1]) came from rank h; row n's
    // outputs go to result row home_row[n] of rank h's buffer -- peer memory reached over NVLink -- laid out
    // [comps | norm | grad | cell (| pad)], peer_ld doubles per row
    int npeers;
    int64_t seg_start[ARB_MAX_PEERS + 1];
    const int64_t* home_row;
    double* peer[ARB_MAX_PEERS];
    int64_t peer_ld;
    // inbox form (arb_query_inbox): q is this rank's inbox, npeers segments of seg_cap rows [coords | home row], segment h
    // filled by rank h's route kernel (arb_route.cu) with inbox_counts[h] rows (device memory: no host round trip)
    const int64_t* inbox_counts;
    int64_t seg_cap;
};

// --------------------------------------------------------------------------------------
// locate: bounds mask, cell index, cell-fraction coordinates (A.py:350-373, 1069-1092).
// The arithmetic is the reference's, operation for operation: IEEE subtract, IEEE divide,
// floor, subtract.  (No fast-math, no reciprocal.)
// --------------------------------------------------------------------------------------
template <int D>
struct Located {
    double frac[D];
    int idx[D];           // per-axis cell index (valid when ok)
    int64_t cell_global;  // total_cells when the row yields NaN
    int64_t cell_local;   // row of `table`
    bool ok;              // evaluate; otherwise every output is NaN
    bool masked;          // row must be NaN-overwritten in place
};

template <int D>
__device__ __forceinline__ Located<D> locate_coords(const QueryParams& p, const double (&c)[D]) {
    Located<D> L;
    bool masked = false;
#pragma unroll
    for (int a = 0; a < D; ++a) masked |= (c[a] < p.mn[a]) | (c[a] > p.mx[a]);
    bool ok = !masked;
    int64_t lin = 0, mult = 1, islow = 0;
#pragma unroll
    for (int a = 0; a < D; ++a) {
        const double iu = __ddiv_rn(__dsub_rn(c[a], p.mn[a]), p.h[a]);
        const double fl = floor(iu);
        L.frac[a] = __dsub_rn(iu, fl);
        ok &= (iu == iu);                      // NaN coordinate: not masked in place, output NaN
        const int64_t ii = ok ? (int64_t)fl : 0;
        ok &= (ii < p.nc[a]);                  // exact upper edge rounding to n-3: NaN
        lin += ii * mult;
        mult *= p.nc[a];
        L.idx[a] = (int)ii;
        if (a == D - 1) islow = ii;
    }
    L.masked = masked;
    L.cell_global = ok ? lin : p.total_cells;
    ok &= (islow >= p.slab_lo) & (islow < p.slab_hi);   // slab-sharded table: not ours -> NaN
    L.ok = ok;
    L.cell_local = ok ? lin - p.slab_lo * p.layer_cells : 0;
    return L;
}

template <int D>
__device__ __forceinline__ Located<D> locate(const QueryParams& p, int64_t n) {
    const double* row = p.q + n * p.ldq;
    double c[D];
#pragma unroll
    for (int a = 0; a < D; ++a) c[a] = row[a];
    return locate_coords<D>(p, c);
}

__device__ __forceinline__ void mask_row_in_place(const QueryParams& p, int64_t n) {
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    double* row = p.q + n * p.ldq;
    for (int64_t c = 0; c < p.ldq; ++c) row[c] = nan;
    if (p.masked_rows) {
        unsigned long long slot = atomicAdd(p.masked_count, 1ULL);
        p.masked_rows[slot] = n;
    }
}

__device__ __forceinline__ double qnan() { return __longlong_as_double(0x7ff8000000000000LL); }

// x^j and d/dx x^j for j = 0..3 selected at run time (lane-dependent j)
__device__ __forceinline__ double pow_sel(double x, int j) {
    const double x2 = x * x;
    return j == 0 ? 1.0 : (j == 1 ? x : (j == 2 ? x2 : x2 * x));
}
__device__ __forceinline__ double dpow_sel(double x, int j) {
    return j == 0 ? 0.0 : (j == 1 ? 1.0 : (j == 2 ? 2.0 * x : 3.0 * (x * x)));
}

// ======================================================================================
// one-thread-per-query evaluation from a contiguous coefficient block (shared or global)
// ======================================================================================
template <bool SHARED>
__device__ __forceinline__ double2 ld_pair(const double* p) {
    if (SHARED) return *reinterpret_cast<const double2*>(p);
    return ldg_stream_d2(p);
}

// value only: nested Horner, highest power first
template <int D, bool SHARED>
__device__ __forceinline__ double eval_value(const double* blk, const double* f) {
    const double u = f[0], v = f[1], w = f[2];
    double out = 0.0;
    constexpr int NT = (D == 4) ? 4 : 1;
#pragma unroll
    for (int lt = NT - 1; lt >= 0; --lt) {
        double val = 0.0;
#pragma unroll
        for (int k = 3; k >= 0; --k) {
            double P = 0.0;
#pragma unroll
            for (int j = 3; j >= 0; --j) {
                const double* r = blk + ((lt * 4 + k) * 4 + j) * 4;
                const double2 lo = ld_pair<SHARED>(r), hi = ld_pair<SHARED>(r + 2);
                const double pp = fma(fma(fma(hi.y, u, hi.x), u, lo.y), u, lo.x);
                P = fma(P, v, pp);
            }
            val = fma(val, w, P);
        }
        out = (D == 4) ? fma(out, f[D - 1], val) : val;
    }
    return out;
}

// value + all partial derivatives (unit-cell coordinates): g[0]=value, g[1..D]=d/du, d/dv, ...
template <int D, bool SHARED>
__device__ __forceinline__ void eval_value_grad(const double* blk, const double* f, double* g) {
    const double u = f[0], v = f[1], w = f[2];
    const double u3 = 3.0 * u;
    double oV = 0.0, oX = 0.0, oY = 0.0, oZ = 0.0, oT = 0.0;
    constexpr int NT = (D == 4) ? 4 : 1;
#pragma unroll
    for (int lt = NT - 1; lt >= 0; --lt) {
        double val = 0.0, gx = 0.0, gy = 0.0, gz = 0.0;
#pragma unroll
        for (int k = 3; k >= 0; --k) {
            double P = 0.0, Px = 0.0, Py = 0.0;
#pragma unroll
            for (int j = 3; j >= 0; --j) {
                const double* r = blk + ((lt * 4 + k) * 4 + j) * 4;
                const double2 lo = ld_pair<SHARED>(r), hi = ld_pair<SHARED>(r + 2);
                const double pp = fma(fma(fma(hi.y, u, hi.x), u, lo.y), u, lo.x);
                const double dp = fma(fma(hi.y, u3, hi.x + hi.x), u, lo.y);
                Py = fma(Py, v, P);
                P = fma(P, v, pp);
                Px = fma(Px, v, dp);
            }
            gz = fma(gz, w, val);
            val = fma(val, w, P);
            gx = fma(gx, w, Px);
            gy = fma(gy, w, Py);
        }
        if (D == 4) {
            const double s = f[D - 1];
            oT = fma(oT, s, oV);
            oV = fma(oV, s, val);
            oX = fma(oX, s, gx);
            oY = fma(oY, s, gy);
            oZ = fma(oZ, s, gz);
        } else {
            oV = val; oX = gx; oY = gy; oZ = gz;
        }
    }
    g[0] = oV; g[1] = oX; g[2] = oY; g[3] = oZ;
    if (D == 4) g[4] = oT;
}


// host: argument validation + parameter block shared by every query-like entry point (arb_query.cu).
// Returns 0 = go, -1 = nothing to do (N == 0), > 0 = error code with arb_last_error() set.
int fill_params(const char* who, const arb_geom* g, bool need_table, const double* table, int mode, double* q,
                int64_t N, int64_t ldq, double* out_comps, double* out_norm, double* out_grad, int64_t* out_cell,
                int64_t* masked_rows, unsigned long long* masked_count, QueryParams& p, bool need_outputs = true);

}  // namespace arb
