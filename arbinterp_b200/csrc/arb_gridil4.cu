// 4-D table-free queries on the component-interleaved grid [nt][nz][ny][nx][4] (quadcubic(..., table=False), modes
// 'vector' and 'both'; rQuery1 / rQuery3 of the reference, A.py:1064-1127, 1190-1258).  The per-component form
// (query_grid4_kernel, arb_query.cu) asks the TMA unit for one 6x4x4 box per component and t-plane: 48-byte rows that
// cost two 32-byte sectors each, 12-16 requests per query.  Here a neighbourhood row is 4 x-points x 4 components =
// 128 aligned bytes that serve every component at once -- the same 8 KB per query as four cell-table blocks, from a
// grid 200x smaller than that table.
//
// Four lanes own a query (lane k = z-plane k), the four t-planes are passes; in a pass every lane's 512-byte slot
// (4 rows of 128 B) is filled by ONE warp-wide cp.async (SASS LDGSTS, 32 lanes x 16 B) whose source comes from the owning
// lane over a shuffle, exactly like the node-table and 3-D interleaved forms of query_block_kernel.  Lanes of different
// queries that want the same plane of the same cell share one slot (DEDUP).  The math, including the A.py:860 term as a
// second set of point weights, is in arb_gridil4.cuh.
#include "arb_device.cuh"
#include "arb_gridil4.cuh"

namespace arb {

// MINB = 3: the compiler keeps to 168 registers so that three CTAs are resident per SM -- 'vector' 0.96 of the cell table
// against 0.79 with two CTAs; the 'both' + quirk form spills ~190 bytes for it and is better off with MINB = 2 (~250
// registers, 0.94 against 0.71; profiles/r02_tablefree_bench_4d.log).  The launcher in arb_query.cu picks.
template <int MODE, bool QUIRK, bool DEDUP, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) query_gridil4_kernel(const QueryParams p) {
    static_assert(MODE == 0 || MODE == 2, "interleaved grid: 'vector' / 'both'");
    constexpr bool BOTH = (MODE == 2);
    constexpr int D = 4, SL = 4, QPW = 8, NV = BOTH ? 8 : 3;
    constexpr uint32_t SLOT = 528;                 // 512 + 16: LDS.128 of neighbouring lanes on distinct bank groups
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int qi = lane / SL, sl = lane % SL;
    unsigned char* const ring = smem + (size_t)wid * 32 * SLOT;
    const int64_t nx = p.nc[0] + 3, ny = p.nc[1] + 3, nz = p.nc[2] + 3;
    // this lane's 16 bytes of a slot: row j = lane / 8 of the plane (nx grid points of 32 B apart), piece lane % 8
    const char* const lane_base = reinterpret_cast<const char*>(p.table) + gridil4::lane_piece_bytes(lane, nx);
    const uint32_t ring0 = smem_u32(ring) + lane * 16;
    const int64_t warp_global = ((int64_t)blockIdx.x * THREADS + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * THREADS) >> 5;
    const int64_t nitem = (p.N + QPW - 1) / QPW;
    double cnext[D];
    {
        const int64_t n0 = warp_global * QPW + qi;
#pragma unroll
        for (int a = 0; a < D; ++a) cnext[a] = (warp_global < nitem && n0 < p.N) ? p.q[n0 * p.ldq + a] : 0.0;
    }
    for (int64_t item = warp_global; item < nitem; item += nwarps) {
        const int64_t n = item * QPW + qi;
        Located<D> L;
        L.ok = false; L.masked = false; L.cell_global = 0; L.cell_local = 0;
        L.idx[0] = L.idx[1] = L.idx[2] = L.idx[3] = 0;
        L.frac[0] = L.frac[1] = L.frac[2] = L.frac[3] = 0.0;
        if (n < p.N) L = locate_coords<D>(p, cnext);
        if (sl == 0 && n < p.N) {
            if (p.out_cell) p.out_cell[n] = L.cell_global;
            if (L.masked) mask_row_in_place(p, n);
        }
        gridil4::Weights W;
        gridil4::make_weights(L.frac, W);
        double acc[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        // lanes of different queries that want the same z-plane of the same cell share one slot in every pass
        int src_lane = lane;
        bool leader = L.ok;
        if (DEDUP) {
            const unsigned peers = __match_any_sync(0xffffffffu, L.ok ? L.cell_global * 4 + sl : (int64_t)(-1 - lane));
            src_lane = __ffs(peers) - 1;
            leader = L.ok && (src_lane == lane);
        }
        const unsigned fmask = __ballot_sync(0xffffffffu, leader);
#pragma unroll 1
        for (int l = 0; l < 4; ++l) {
            // grids < 128 GB (checked by the launcher): the count of 32-byte points fits 32 bits
            const uint32_t src32 = (uint32_t)gridil4::plane_first_point(L.idx, sl, l, nx, ny, nz);
#pragma unroll 8
            for (int o = 0; o < 32; ++o) {
                if ((fmask >> o) & 1u) {
                    const uint32_t s = __shfl_sync(0xffffffffu, src32, o);
                    cp_async_16(ring0 + o * SLOT, lane_base + ((size_t)s << 5));
                }
            }
            if (l == 0) {                         // next item's coordinates: in flight during the wait
                const int64_t n1 = (item + nwarps) * QPW + qi;
                if (item + nwarps < nitem && n1 < p.N) {
#pragma unroll
                    for (int a = 0; a < D; ++a) cnext[a] = p.q[n1 * p.ldq + a];
                }
            }
            cp_async_wait_all();
            __syncwarp();
            if (L.ok)
                gridil4::pass<BOTH, QUIRK>(acc, reinterpret_cast<const double*>(ring + (size_t)src_lane * SLOT), sl, l, L.frac, W);
            __syncwarp();                         // every lane is done with the slots before the next copies land
        }
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 1);
            acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 2);
        }
        if (n < p.N && sl == 0) {
            const double nan = qnan();
#pragma unroll
            for (int c = 0; c < 3; ++c) p.out_comps[n * 3 + c] = L.ok ? acc[c] : nan;
            if (BOTH) {
                p.out_norm[n] = L.ok ? acc[3] : nan;
#pragma unroll
                for (int a = 0; a < D; ++a) p.out_grad[n * D + a] = L.ok ? __ddiv_rn(acc[4 + a], p.h[a]) : nan;
            }
        }
    }
}

template <int MODE, bool QUIRK, bool DEDUP, int MINB>
static int launch_gridil4(const QueryParams& p, cudaStream_t st) {
    constexpr int THREADS = 128;
    const size_t smem = (size_t)(THREADS / 32) * 32 * 528;
    auto k = query_gridil4_kernel<MODE, QUIRK, DEDUP, THREADS, MINB>;
    static int occ_cache[16] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    int& occ = occ_cache[dev & 15];
    if (occ == 0) {
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int o = 1;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, k, THREADS, smem) != cudaSuccess || o < 1) o = 1;
        occ = o;
    }
    const int64_t items = (p.N + 7) / 8, blocks = (items + THREADS / 32 - 1) / (THREADS / 32);
    int64_t grid = (int64_t)num_sms() * occ;
    if (blocks < grid) grid = blocks;
    k<<<(unsigned)(grid < 1 ? 1 : grid), THREADS, smem, st>>>(p);
    return check_cuda(cudaGetLastError(), "query_gridil4_kernel launch");
}

// called by query_gridil_device (arb_query.cu) for d == 4; p is filled and checked there
template <int MODE, bool QUIRK>
static int pick(const QueryParams& p, bool dedup, bool wide, cudaStream_t st) {
    if (wide) return dedup ? launch_gridil4<MODE, QUIRK, true, 2>(p, st) : launch_gridil4<MODE, QUIRK, false, 2>(p, st);
    return dedup ? launch_gridil4<MODE, QUIRK, true, 3>(p, st) : launch_gridil4<MODE, QUIRK, false, 3>(p, st);
}

int query_gridil4_launch(const QueryParams& p, int mode, bool quirk, bool dedup, bool wide, cudaStream_t st) {
    if (mode == ARB_MODE_VECTOR) return quirk ? pick<0, true>(p, dedup, wide, st) : pick<0, false>(p, dedup, wide, st);
    return quirk ? pick<2, true>(p, dedup, wide, st) : pick<2, false>(p, dedup, wide, st);
}

}  // namespace arb
