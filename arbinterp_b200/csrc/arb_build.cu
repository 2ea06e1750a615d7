// Coefficient build: for every interpolatable cell, alpha = inv(B) * (D * f[4^d neighbourhood]).
// Replaces calcCoefficients1/2/3 + neighbourInd + allCoeffs of the reference
// (A.py:570-621, 523-525 tricubic; A.py:1322-1381, 1260-1262 quadcubic), which does the same with
// one Python-level dgemv of the fused matrix A = inv(B) D per cell and component.
//
// One CTA builds one tile of cells for one component:
//   1. TMA (cp.async.bulk.tensor, SASS UTMALDG) stages the grid tile plus its 3-point halo into
//      shared memory; out-of-range points are zero-filled by the TMA unit and only ever feed
//      cells that are not stored.
//   2. stencil stage: the 2^d derivative fields (f, fx, fy, ..., fxyz[t]; central differences in
//      unit-cell coordinates, the rows of D, A.py:129-173 / 762-876) are evaluated at every
//      cell-corner point of the tile and kept in shared memory.  A cell's b-vector is a gather
//      of 2^d corners x 2^d types from these fields, so it is never materialised.
//   3. solve stage: alpha^T[cells x 4^d] = b^T[cells x 4^d] * inv(B)^T on the FP64 tensor cores
//      (mma.sync m8n8k4 f64, SASS DMMA).  inv(B) is integer (|entries| <= 27 / 81) and is read as
//      pre-swizzled fragments from a small L1/L2-resident buffer; accumulators stay in
//      registers and go straight to the cell-major table with 16-byte stores.
// The 4-D reference matrix carries the A.py:860 off-by-one (row 240 of D is zero and rows
// 241..255 use the stencil centre of the previous corner); it is reproduced in the gather of
// step 3, where the quadruple-mixed type reads corner c-1 (and 0 for c = 0).
#include <vector>
#include <mutex>
#include "arb_common.cuh"

namespace arb {

template <int D> struct BuildCfg;
template <> struct BuildCfg<3> {
    static constexpr int TX = 8, TY = 4, TZ = 4, TT = 1;
    static constexpr int WARPS = 4, NT_W = 2;      // 8 n-tiles of 8 coefficients
};
template <> struct BuildCfg<4> {
    static constexpr int TX = 8, TY = 2, TZ = 2, TT = 2;
    static constexpr int WARPS = 8, NT_W = 4;      // 32 n-tiles
};

template <int D>
struct BuildShape {
    using Cfg = BuildCfg<D>;
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
    unsigned char mask_of_type[16];
};

template <int D>
__global__ void __launch_bounds__(BuildShape<D>::THREADS)
build_kernel(const __grid_constant__ CUtensorMap tmap, const BuildParams p) {
    using S = BuildShape<D>;
    using Cfg = BuildCfg<D>;
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
    __syncthreads();
    mbar_wait(&bar, 0);

    // ---- stencil stage -------------------------------------------------------------
    // corner point (px,py,pz,pt) of the tile sits at grid-tile coordinate (+1,+1,+1,+1)
    for (int e = threadIdx.x; e < S::NPOINT * S::NTYPE; e += S::THREADS) {
        const int type = e / S::NPOINT;
        int pt_ = e - type * S::NPOINT;
        const int px = pt_ % S::PX; pt_ /= S::PX;
        const int py = pt_ % S::PY; pt_ /= S::PY;
        const int pz = pt_ % S::PZ;
        const int pt = pt_ / S::PZ;
        const int centre = (((D == 4 ? (pt + 1) : 0) * S::GZ + (pz + 1)) * S::GY + (py + 1)) * S::GX + (px + 1);
        const int mask = p.mask_of_type[type];
        const int strides[4] = {1, S::GX, S::GX * S::GY, S::GX * S::GY * S::GZ};
        int axes[4], na = 0;
#pragma unroll
        for (int a = 0; a < D; ++a)
            if ((mask >> a) & 1) axes[na++] = a;
        double acc = 0.0, w = 1.0;
        for (int i = 0; i < na; ++i) w *= 0.5;
        for (int sgn = 0; sgn < (1 << na); ++sgn) {
            int off = 0;
            bool neg = false;
            for (int i = 0; i < na; ++i) {
                if ((sgn >> i) & 1) off += strides[axes[i]];
                else { off -= strides[axes[i]]; neg = !neg; }
            }
            const double v = gtile[centre + off];
            acc += neg ? -v : v;
        }
        deriv[type * S::TYPE_STRIDE + ((pt * S::PZ + pz) * S::PY + py) * S::PXS + px] = acc * w;
    }
    __syncthreads();

    // ---- solve stage (DMMA) ----------------------------------------------------------
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cellx = lane >> 2;              // A-fragment row = cell along x
    const int kq = lane & 3;                  // A-fragment column = corner (cx, cy) within a k-step
    double acc[S::MT][Cfg::NT_W][2];
#pragma unroll
    for (int m = 0; m < S::MT; ++m)
#pragma unroll
        for (int j = 0; j < Cfg::NT_W; ++j) acc[m][j][0] = acc[m][j][1] = 0.0;

    constexpr int KS_PER_TYPE = S::NTYPE / 4;     // k-steps per derivative type: 2 (3-D), 4 (4-D)
    constexpr int NKS = S::NM / 4;
    const int lane_off = (kq >> 1) * S::PXS + cellx + (kq & 1);   // (cy, cx) part of the address
#pragma unroll 1
    for (int ks = 0; ks < NKS; ++ks) {
        double bf[Cfg::NT_W];
#pragma unroll
        for (int j = 0; j < Cfg::NT_W; ++j)
            bf[j] = __ldg(p.bfrag + ((size_t)ks * (S::NM / 8) + warp * Cfg::NT_W + j) * 32 + lane);
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

template <int D>
static int build_impl(const double* grid, int ncomp, const int64_t* n, double* table, int quirk, cudaStream_t st) {
    using S = BuildShape<D>;
    using Cfg = BuildCfg<D>;
    EncodeTiledFn encode = get_encode_fn();
    if (!encode) { set_error("arb_build_coeffs: cuTensorMapEncodeTiled not available from the driver"); return 2; }

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
    for (int r = 0; r < (1 << D); ++r) p.mask_of_type[r] = (unsigned char)deriv_mask(D, r);
    { const int frc = get_bfrag(D, &p.bfrag); if (frc) return frc; }

    // TMA needs 16-byte global strides: pad odd nx to even in a scratch copy
    const double* src = grid;
    double* padded = nullptr;
    int64_t pitch = n[0];
    int64_t rows = ncomp;
    for (int a = 1; a < D; ++a) rows *= n[a];
    if (n[0] & 1) {
        pitch = n[0] + 1;
        ARB_CUDA(cudaMallocAsync(&padded, sizeof(double) * pitch * rows, st));
        ARB_CUDA(cudaMemsetAsync(padded, 0, sizeof(double) * pitch * rows, st));
        ARB_CUDA(cudaMemcpy2DAsync(padded, pitch * 8, grid, n[0] * 8, n[0] * 8, rows, cudaMemcpyDeviceToDevice, st));
        src = padded;
    }
    if ((reinterpret_cast<uintptr_t>(src) & 15) != 0) {
        if (padded) cudaFreeAsync(padded, st);
        set_error("arb_build_coeffs: grid pointer must be 16-byte aligned");
        return 1;
    }

    CUtensorMap tmap;
    cuuint64_t gdim[5], gstr[4];
    cuuint32_t box[5], estr[5];
    const int rank = D + 1;
    gdim[0] = (cuuint64_t)n[0];
    int64_t stride = pitch * 8;
    for (int a = 1; a < D; ++a) { gdim[a] = (cuuint64_t)n[a]; gstr[a - 1] = (cuuint64_t)stride; stride *= n[a]; }
    gdim[D] = (cuuint64_t)ncomp; gstr[D - 1] = (cuuint64_t)stride;
    box[0] = S::GX; box[1] = S::GY; box[2] = S::GZ;
    if (D == 4) box[3] = S::GT;
    box[D] = 1;
    for (int a = 0; a < rank; ++a) estr[a] = 1;
    CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, rank, const_cast<double*>(src), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) {
        if (padded) cudaFreeAsync(padded, st);
        set_error("arb_build_coeffs: cuTensorMapEncodeTiled failed with CUresult %d", (int)cr);
        return 2;
    }

    auto k = build_kernel<D>;
    ARB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::SMEM));
    dim3 gridDim((unsigned)ntiles, (unsigned)ncomp, 1);
    k<<<gridDim, S::THREADS, S::SMEM, st>>>(tmap, p);
    ARB_CUDA(cudaGetLastError());
    const int64_t tail = (int64_t)ncomp * S::NM;
    fill_nan_kernel<<<(unsigned)((tail + 255) / 256), 256, 0, st>>>(table + ncell * tail, tail);
    ARB_CUDA(cudaGetLastError());
    if (padded) ARB_CUDA(cudaFreeAsync(padded, st));
    return 0;
}

}  // namespace arb

extern "C" {

int arb_build_coeffs(int d, const double* grid, int ncomp, const int64_t n[4], double* table, int reference_quirk,
                     void* stream) {
    if (!grid || !table || !n) { arb::set_error("arb_build_coeffs: null pointer"); return 1; }
    if (ncomp < 1 || ncomp > 4) { arb::set_error("arb_build_coeffs: ncomp=%d not in 1..4", ncomp); return 1; }
    if (d == 3) return arb::build_impl<3>(grid, ncomp, n, table, reference_quirk, (cudaStream_t)stream);
    if (d == 4) return arb::build_impl<4>(grid, ncomp, n, table, reference_quirk, (cudaStream_t)stream);
    arb::set_error("arb_build_coeffs: d=%d not in {3,4}", d);
    return 1;
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
