// Node (Hermite) table build: the 2^d central-difference values f, fx, fy, fxy, ... of every interior grid node
// (the rows of the reference's D matrix, A.py:129-173 / 762-876, applied at the node instead of at the corners of
// every cell).  One thread per (node, component): 3^d neighbourhood reads served by L1/L2 (the grid is read once
// from HBM), 64 / 128 contiguous bytes written per thread with 16-byte streaming stores (3-D: twice, every node is
// the right half of one x-pair and the left half of the next) -- an HBM-write stream of 16 x the grid size, 4x (3-D) /
// 16x (4-D) smaller than the cell table the separable build kernels write.
#include "arb_common.cuh"
#include "arb_nodes.cuh"

namespace arb {

// LAYOUT: 0 = 4-D per component [C][nt'][nz'][ny'][nx'][16]; 1 = 3-D one component, aligned x-pairs [C][nz'][ny'][nx'-1][2][8];
//         2 = 3-D components interleaved [nz'][ny'][nx'][4][8].
// A block takes ROWS grid rows x 64 nodes along x.  Every thread evaluates the stencil of one (node, component) into a
// k-major shared-memory tile (conflict-free), then the block writes the tile's share of the table -- contiguous in every
// layout -- with 16-byte stores that are consecutive across the lanes.
template <int D, int LAYOUT>
__global__ void __launch_bounds__(256) build_nodes_kernel(const double* __restrict__ grid, int ncomp, int64_t nx,
                                                          int64_t ny, int64_t nz, int64_t nt, int64_t pitch_x,
                                                          double* __restrict__ out) {
    constexpr int T = (D == 3) ? 8 : 16;
    constexpr int NXB = 64;                                   // nodes along x per block
    constexpr int CI = (LAYOUT == 2) ? 4 : 1;                 // components held together in the tile
    constexpr int ROWS = 256 / (NXB * CI);                    // grid rows per block (4, or 1 when interleaved)
    constexpr int NODES = NXB + (LAYOUT == 1 ? 1 : 0);        // pairs need the node after the last one
    constexpr int K = T * CI;                                 // doubles per node in the tile
    constexpr int LDT = NODES + 1;                            // k-major tile: [row][k][node], odd pitch
    __shared__ double tile[ROWS * K * LDT];
    const int64_t m0 = nx - 2, m1 = ny - 2, m2 = nz - 2, m3 = (D == 4) ? nt - 2 : 1;
    const int64_t nrows = (int64_t)(LAYOUT == 2 ? 1 : ncomp) * m3 * m2 * m1;      // rows of nodes (all but x)
    const int64_t xblocks = (m0 + NXB - 1) / NXB;
    const int64_t rowgroups = (nrows + ROWS - 1) / ROWS;
    const int64_t sy = pitch_x, sz = pitch_x * ny, st = pitch_x * ny * nz, scomp = st * ((D == 4) ? nt : 1);
    const int tx = threadIdx.x % NXB, tc = (threadIdx.x / NXB) % CI, tr = threadIdx.x / (NXB * CI);
    for (int64_t blk = blockIdx.x; blk < rowgroups * xblocks; blk += gridDim.x) {
        const int64_t rg = blk / xblocks, x0 = (blk - rg * xblocks) * NXB;
        // ---- stencils into the tile
        for (int pass = 0; pass < (LAYOUT == 1 ? 2 : 1); ++pass) {
            const int node = (pass == 0) ? tx : NXB;          // pass 1: the extra node of the pair layout, lanes tx == 0
            if (pass == 1 && tx != 0) break;
            const int64_t row = rg * ROWS + tr, x = x0 + node;
            if (row < nrows && x < m0) {
                int64_t r = row;
                const int64_t y = r % m1; r /= m1;
                const int64_t z = r % m2; r /= m2;
                const int64_t t = (D == 4) ? r % m3 : 0;
                const int64_t c = (LAYOUT == 2) ? tc : ((D == 4) ? r / m3 : r);
                double v[T];
                if (c < ncomp) {
                    const double* centre = grid + c * scomp + ((D == 4) ? (t + 1) * st : 0) + (z + 1) * sz + (y + 1) * sy + (x + 1);
                    auto get = [&](int dx, int dy, int dz, int dt) { return __ldg(centre + dx + dy * sy + dz * sz + dt * st); };
                    nodes::node_stencil<D>(get, v);
                } else {
#pragma unroll
                    for (int k = 0; k < T; ++k) v[k] = 0.0;
                }
#pragma unroll
                for (int k = 0; k < T; ++k) tile[(tr * K + tc * T + k) * LDT + node] = v[k];
            }
        }
        __syncthreads();
        // ---- the tile's share of the table, 16 bytes per thread and trip, consecutive across the lanes
        for (int rr = 0; rr < ROWS; ++rr) {
            const int64_t row = rg * ROWS + rr;
            if (row >= nrows) break;
            const int64_t nvalid = (m0 - x0 < NXB) ? (m0 - x0) : NXB;                     // nodes of this block in the row
            if (LAYOUT == 1) {
                // pairs x0 .. : pair i = nodes (i, i + 1); the row holds m0 - 1 pairs of 16 doubles
                const int64_t npair = ((m0 - 1 - x0) < NXB) ? (m0 - 1 - x0) : NXB;
                double* dst = out + (row * (m0 - 1) + x0) * 16;
                for (int64_t ch = threadIdx.x; ch < npair * 8; ch += 256) {
                    const int pr = (int)(ch >> 3), w = (int)(ch & 7);
                    const int node = pr + (w >> 2), k = (w & 3) * 2;
                    stg_stream_d2(dst + ch * 2, tile[(rr * K + k) * LDT + node], tile[(rr * K + k + 1) * LDT + node]);
                }
            } else {
                double* dst = out + (row * m0 + x0) * K;
                for (int64_t ch = threadIdx.x; ch < nvalid * (K / 2); ch += 256) {
                    const int node = (int)(ch / (K / 2)), k = (int)(ch % (K / 2)) * 2;
                    stg_stream_d2(dst + ch * 2, tile[(rr * K + k) * LDT + node], tile[(rr * K + k + 1) * LDT + node]);
                }
            }
        }
        __syncthreads();
    }
}

int build_nodes_device(int d, const double* grid, int ncomp, const int64_t* npts, int64_t pitch_x, double* out,
                       cudaStream_t st) {
    if ((d != 3 && d != 4) || !grid || !out || ncomp < 1 || !npts) { set_error("arb_build_nodes: bad arguments"); return 1; }
    for (int a = 0; a < d; ++a)
        if (npts[a] < 4) { set_error("arb_build_nodes: axis %d has %lld points, need >= 4", a, (long long)npts[a]); return 1; }
    if (pitch_x < npts[0]) { set_error("arb_build_nodes: pitch_x < nx"); return 1; }
    if (reinterpret_cast<uintptr_t>(out) & 127) { set_error("arb_build_nodes: node table must be 128-byte aligned"); return 1; }
    if (d == 3 && ncomp >= 3 && ncomp > 4) { set_error("arb_build_nodes: at most 4 interleaved components"); return 1; }
    const bool il = (d == 3 && ncomp >= 3);
    int64_t rows = il ? 1 : ncomp;
    for (int a = 1; a < d; ++a) rows *= npts[a] - 2;
    const int rows_per_block = il ? 1 : 4;
    int64_t blocks = ((rows + rows_per_block - 1) / rows_per_block) * ((npts[0] - 2 + 63) / 64);
    const int64_t cap = (int64_t)num_sms() * 32;
    if (blocks > cap) blocks = cap;
    if (d == 4)
        build_nodes_kernel<4, 0><<<(int)blocks, 256, 0, st>>>(grid, ncomp, npts[0], npts[1], npts[2], npts[3], pitch_x, out);
    else if (il)
        build_nodes_kernel<3, 2><<<(int)blocks, 256, 0, st>>>(grid, ncomp, npts[0], npts[1], npts[2], 1, pitch_x, out);
    else
        build_nodes_kernel<3, 1><<<(int)blocks, 256, 0, st>>>(grid, ncomp, npts[0], npts[1], npts[2], 1, pitch_x, out);
    return check_cuda(cudaGetLastError(), "build_nodes_kernel launch");
}

}  // namespace arb

extern "C" int arb_build_nodes(int d, const double* grid, int ncomp, const int64_t* npts, int64_t pitch_x, double* nodes,
                               void* stream) {
    return arb::build_nodes_device(d, grid, ncomp, npts, pitch_x, nodes, (cudaStream_t)stream);
}
