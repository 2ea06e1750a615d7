// Node (Hermite) table build: the 2^d central-difference values f, fx, fy, fxy, ... of every interior grid node
// (the rows of the reference's D matrix, A.py:129-173 / 762-876, applied at the node instead of at the corners of
// every cell).  One thread per (node, component): 3^d neighbourhood reads served by L1/L2 (the grid is read once
// from HBM), 64 / 128 contiguous bytes written per thread with 16-byte streaming stores (3-D: twice, every node is
// the right half of one x-pair and the left half of the next) -- an HBM-write stream of 16 x the grid size, 4x (3-D) /
// 16x (4-D) smaller than the cell table the separable build kernels write.
#include "arb_common.cuh"
#include "arb_nodes.cuh"

namespace arb {

template <int D>
__global__ void __launch_bounds__(256) build_nodes_kernel(const double* __restrict__ grid, int ncomp, int64_t nx,
                                                          int64_t ny, int64_t nz, int64_t nt, int64_t pitch_x,
                                                          double* __restrict__ out) {
    constexpr int T = (D == 3) ? 8 : 16;
    const int64_t m0 = nx - 2, m1 = ny - 2, m2 = nz - 2, m3 = (D == 4) ? nt - 2 : 1;
    const bool interleaved = (D == 3 && ncomp >= 3);      // [node][4][8]: Bx, By, Bz, |B| (zeros when ncomp == 3)
    const int64_t per_comp = m0 * m1 * m2 * m3, total = per_comp * (interleaved ? 4 : ncomp);
    const int64_t sy = pitch_x, sz = pitch_x * ny, st = pitch_x * ny * nz;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = i / per_comp;
        int64_t r = i - c * per_comp;
        const int64_t x = r % m0; r /= m0;
        const int64_t y = r % m1; r /= m1;
        const int64_t z = r % m2;
        const int64_t t = r / m2;
        const double* centre = grid + c * (st * ((D == 4) ? nt : 1)) + ((D == 4) ? (t + 1) * st : 0) + (z + 1) * sz +
                               (y + 1) * sy + (x + 1);
        auto get = [&](int dx, int dy, int dz, int dt) { return __ldg(centre + dx + dy * sy + dz * sz + dt * st); };
        double v[T];
        if (c < ncomp) {
            nodes::node_stencil<D>(get, v);
        } else {
#pragma unroll
            for (int k = 0; k < T; ++k) v[k] = 0.0;
        }
        if (interleaved) {
            double* dst = out + ((i - c * per_comp) * 4 + c) * 8;
#pragma unroll
            for (int k = 0; k < T; k += 2) stg_stream_d2(dst + k, v[k], v[k + 1]);
        } else if (D == 4) {
            double* dst = out + i * T;
#pragma unroll
            for (int k = 0; k < T; k += 2) stg_stream_d2(dst + k, v[k], v[k + 1]);
        } else {
            // 3-D: x-pairs.  Pair ix holds nodes ix, ix + 1 in 128 aligned bytes, so a query's four rows are whole
            // 128-byte lines (a 64-byte node at an odd position would make the L2 fill two lines for one: measured
            // 730 instead of 536 DRAM bytes per query, profiles/r02_nodes3d_norm_ncu.txt before this layout)
            const int64_t npair = m0 - 1;
            double* row = out + ((c * m2 + z) * m1 + y) * npair * 16;
            if (x < npair) {
#pragma unroll
                for (int k = 0; k < T; k += 2) stg_stream_d2(row + x * 16 + k, v[k], v[k + 1]);
            }
            if (x > 0) {
#pragma unroll
                for (int k = 0; k < T; k += 2) stg_stream_d2(row + (x - 1) * 16 + 8 + k, v[k], v[k + 1]);
            }
        }
    }
}

int build_nodes_device(int d, const double* grid, int ncomp, const int64_t* npts, int64_t pitch_x, double* out,
                       cudaStream_t st) {
    if ((d != 3 && d != 4) || !grid || !out || ncomp < 1 || !npts) { set_error("arb_build_nodes: bad arguments"); return 1; }
    for (int a = 0; a < d; ++a)
        if (npts[a] < 4) { set_error("arb_build_nodes: axis %d has %lld points, need >= 4", a, (long long)npts[a]); return 1; }
    if (pitch_x < npts[0]) { set_error("arb_build_nodes: pitch_x < nx"); return 1; }
    if (reinterpret_cast<uintptr_t>(out) & 127) { set_error("arb_build_nodes: node table must be 128-byte aligned"); return 1; }
    int64_t total = (d == 3 && ncomp >= 3) ? 4 : ncomp;
    for (int a = 0; a < d; ++a) total *= npts[a] - 2;
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = (int64_t)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    if (d == 3)
        build_nodes_kernel<3><<<(int)blocks, 256, 0, st>>>(grid, ncomp, npts[0], npts[1], npts[2], 1, pitch_x, out);
    else
        build_nodes_kernel<4><<<(int)blocks, 256, 0, st>>>(grid, ncomp, npts[0], npts[1], npts[2], npts[3], pitch_x, out);
    return check_cuda(cudaGetLastError(), "build_nodes_kernel launch");
}

}  // namespace arb

extern "C" int arb_build_nodes(int d, const double* grid, int ncomp, const int64_t* npts, int64_t pitch_x, double* nodes,
                               void* stream) {
    return arb::build_nodes_device(d, grid, ncomp, npts, pitch_x, nodes, (cudaStream_t)stream);
}
