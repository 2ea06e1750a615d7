/* Plain-C client of the C ABI (include/arbinterp_b200.h): what a non-Python host would do.
 * Builds the coefficient table of a small quadratic field, queries it through the device entry point
 * (arb_query) and through the host-buffer entry point (arb_query_host), and checks both against the
 * analytic field: a tricubic interpolant with central-difference slopes reproduces per-axis quadratics
 * exactly, so value and gradient must match to round-off.  Also checks the NaN / cell-index conventions
 * (A.py:350-355, 368-370) and the error path.  Test infrastructure; prints "c_smoke ok". */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <cuda_runtime_api.h>
#include "arbinterp_b200.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s\n", cudaGetErrorString(e_), #x); return 2; } } while (0)
#define ARB(x) do { int r_ = (x); if (r_) { printf("arb error %d: %s at %s\n", r_, arb_last_error(), #x); return 3; } } while (0)

static double fval(double x, double y, double z) { return 1.0 + 2.0 * x - y + 0.5 * z + x * x - 2.0 * y * y + 3.0 * z * z; }

int main(void) {
    enum { NX = 14, NY = 11, NZ = 9, NQ = 1000 };
    const double x0 = -1.0, y0 = 0.5, z0 = 2.0, hx = 0.1, hy = 0.25, hz = 0.5;
    static double grid[NZ][NY][NX];
    for (int k = 0; k < NZ; ++k)
        for (int j = 0; j < NY; ++j)
            for (int i = 0; i < NX; ++i) grid[k][j][i] = fval(x0 + hx * i, y0 + hy * j, z0 + hz * k);

    arb_geom g;
    memset(&g, 0, sizeof g);
    g.d = 3; g.ncomp = 1;
    g.ncell[0] = NX - 3; g.ncell[1] = NY - 3; g.ncell[2] = NZ - 3; g.ncell[3] = 1;
    g.slab_lo = 0; g.slab_hi = NZ - 3;
    g.int_min[0] = x0 + hx; g.int_min[1] = y0 + hy; g.int_min[2] = z0 + hz;
    g.int_max[0] = x0 + hx * (NX - 2); g.int_max[1] = y0 + hy * (NY - 2); g.int_max[2] = z0 + hz * (NZ - 2);
    g.h[0] = hx; g.h[1] = hy; g.h[2] = hz; g.h[3] = 1.0;
    const long long ncell = (long long)(NX - 3) * (NY - 3) * (NZ - 3);

    printf("%s\n", arb_version());
    double *d_grid, *d_table, *d_q, *d_norm, *d_grad;
    long long* d_cell;
    CK(cudaMalloc((void**)&d_grid, sizeof grid));
    CK(cudaMalloc((void**)&d_table, (size_t)(ncell + 1) * 64 * sizeof(double)));
    CK(cudaMemcpy(d_grid, grid, sizeof grid, cudaMemcpyHostToDevice));
    ARB(arb_build_coeffs_3d(d_grid, 1, NX, NY, NZ, d_table, NULL));

    static double q[NQ][4], q2[NQ][4], norm[NQ], grad[NQ][3], norm2[NQ], grad2[NQ][3];
    static long long cell[NQ], cell2[NQ];
    unsigned s = 12345u;
    for (int n = 0; n < NQ; ++n) {
        for (int a = 0; a < 3; ++a) {
            s = s * 1664525u + 1013904223u;
            q[n][a] = g.int_min[a] + (g.int_max[a] - g.int_min[a]) * ((s >> 8) / 16777216.0) * 0.999999;
        }
        q[n][3] = 42.0;                               /* extra column: ignored, NaN-masked with the row */
    }
    q[7][1] = g.int_max[1] + 1.0;                     /* outside the volume */
    memcpy(q2, q, sizeof q);

    CK(cudaMalloc((void**)&d_q, sizeof q));
    CK(cudaMalloc((void**)&d_norm, sizeof norm));
    CK(cudaMalloc((void**)&d_grad, sizeof grad));
    CK(cudaMalloc((void**)&d_cell, sizeof cell));
    CK(cudaMemcpy(d_q, q, sizeof q, cudaMemcpyHostToDevice));
    ARB(arb_query(&g, d_table, ARB_MODE_NORM, d_q, NQ, 4, NULL, d_norm, d_grad, (int64_t*)d_cell, NULL, NULL, NULL));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(q, d_q, sizeof q, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(norm, d_norm, sizeof norm, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(grad, d_grad, sizeof grad, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(cell, d_cell, sizeof cell, cudaMemcpyDeviceToHost));
    ARB(arb_query_host(&g, d_table, ARB_MODE_NORM, &q2[0][0], NQ, 4, NULL, norm2, &grad2[0][0], (int64_t*)cell2, 0));

    double worst = 0.0;
    for (int n = 0; n < NQ; ++n) {
        if (n == 7) {
            if (!(isnan(norm[n]) && isnan(grad[n][2]) && isnan(q[n][0]) && isnan(q[n][3]) && cell[n] == ncell)) { printf("NaN convention broken\n"); return 4; }
            if (!(isnan(norm2[n]) && isnan(q2[n][3]) && cell2[n] == ncell)) { printf("NaN convention broken (host path)\n"); return 4; }
            continue;
        }
        const double x = q[n][0], y = q[n][1], z = q[n][2];
        const double ref[4] = {fval(x, y, z), 2.0 + 2.0 * x, -1.0 - 4.0 * y, 0.5 + 6.0 * z};
        const double got[4] = {norm[n], grad[n][0], grad[n][1], grad[n][2]};
        const double got2[4] = {norm2[n], grad2[n][0], grad2[n][1], grad2[n][2]};
        for (int a = 0; a < 4; ++a) {
            const double e = fabs(got[a] - ref[a]) / fmax(fabs(ref[a]), 50.0);
            if (e > worst) worst = e;
            if (got[a] != got2[a]) { printf("device and host entry points differ at row %d\n", n); return 5; }
        }
        const long long ix = (long long)floor((x - g.int_min[0]) / hx), iy = (long long)floor((y - g.int_min[1]) / hy),
                        iz = (long long)floor((z - g.int_min[2]) / hz);
        if (cell[n] != ix + (NX - 3) * (iy + (long long)(NY - 3) * iz) || cell2[n] != cell[n]) { printf("cell index mismatch at row %d\n", n); return 6; }
    }
    if (worst > 1e-12) { printf("max scaled error %.3e\n", worst); return 7; }
    if (arb_build_coeffs_3d(d_grid, 7, NX, NY, NZ, d_table, NULL) == 0 || strlen(arb_last_error()) == 0) { printf("bad ncomp accepted\n"); return 8; }
    cudaFree(d_grid); cudaFree(d_table); cudaFree(d_q); cudaFree(d_norm); cudaFree(d_grad); cudaFree(d_cell);
    printf("c_smoke ok: max scaled error %.3e\n", worst);
    return 0;
}
