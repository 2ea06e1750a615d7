/*
 * arbinterp_b200 -- C ABI of the B200-native ARBInterp interpolation hot path.
 *
 * This is the drop-in boundary: plain C, plain pointers and sizes, no torch types.
 * The reference (DurhamDecLab/ARBInterp) is pure Python/numpy and has no FFI of its own;
 * each entry point below names the reference method it replaces
 * (A.py = src/ARBInterp/ARBInterp.py in the reference tree).  The Python classes in
 * arbinterp_b200/interp.py bind these through ctypes; INTEGRATION.md shows the stub a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; arb_last_error() then
 *     returns a thread-local, NUL-terminated description.  No exceptions cross the ABI.
 *   - "device pointer" = CUDA device memory of the current device; caller-owned.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Device
 *     entry points only enqueue work; they never synchronise the host unless stated.
 *   - all floating point is IEEE float64; cell indices are int64.
 *   - re-entrant for distinct streams/buffers; a table is read-only during queries.
 *
 * Data layout
 *   grid   : [C][nt][nz][ny][nx] float64, x fastest (the reference's sorted row order,
 *            A.py:530-532 / 1266-1269, one dense plane per interpolated component).
 *   table  : [ncell_local + 1][C][4^d] float64, cell-major; coefficient m = i + 4j + 16k
 *            (+ 64l) multiplies u^i v^j w^k (s^l) (A.py:380-382 / 1107-1110).  Row
 *            `ncell_local` is the NaN sentinel (A.py:45,57,70).  Component order for C=3 is
 *            (x,y,z), for C=4 (x,y,z,norm), for C=1 the scalar / norm.
 *   query  : [N][ldq] float64, first d columns are coordinates, the rest is ignored
 *            (README "Further columns can be present").
 */
#ifndef ARBINTERP_B200_H
#define ARBINTERP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ARB_MODE_VECTOR 0 /* Query1 / rQuery1: components only        (A.py:344-397, 1064-1127) */
#define ARB_MODE_NORM   1 /* Query2 / rQuery2: value + gradient       (A.py:399-454, 1129-1188) */
#define ARB_MODE_BOTH   2 /* Query3 / rQuery3: components, value, grad (A.py:457-521, 1190-1258) */

/* Geometry the reference derives in getFieldParams (A.py:541-568, 1288-1320). */
typedef struct arb_geom {
    int32_t d;            /* 3 = tricubic, 4 = quadcubic                                        */
    int32_t ncomp;        /* C, components stored per cell in the table (1, 3 or 4)              */
    int64_t ncell[4];     /* interpolatable cells per axis, n-3 (A.py:545, 1288-1291)            */
    int64_t slab_lo;      /* first cell layer of the slowest axis held by `table` (0 = whole)    */
    int64_t slab_hi;      /* one past the last cell layer held (ncell[d-1] = whole)              */
    double  int_min[4];   /* xIntMin.. : second grid coordinate per axis (A.py:551-553)          */
    double  int_max[4];   /* xIntMax.. : second-to-last grid coordinate  (A.py:554-556)          */
    double  h[4];         /* hx.. = |axis[0]-axis[1]| (A.py:547-549)                             */
    int32_t flags;        /* ARB_GEOM_* bits; 0 = reference behaviour                            */
    int32_t reserved;
} arb_geom;

/* Table-free quadcubic queries evaluate with the corrected 4-D matrix (no A.py:860 off-by-one) when this
 * bit is set; coefficient tables carry the choice made at build time (arb_build_coeffs' reference_quirk). */
#define ARB_GEOM_FIXED_D4 1

const char* arb_version(void);
const char* arb_last_error(void);

/* Constant matrices of makeAMatrix (A.py:107-175, 726-878), generated exactly (no LAPACK).
 *   which: 0 = inv(B) (integer Hermite inverse), 1 = D (finite differences), 2 = A = inv(B)*D.
 *   reference_quirk != 0 reproduces the A.py:860 off-by-one in the 4-D matrix (the default
 *   everywhere in this library; parity with the reference needs it).
 *   out_host: HOST buffer of 4^d * 4^d doubles, row-major. */
int arb_get_matrix(int d, int which, int reference_quirk, double* out_host);

/* Coefficient build = allCoeffs() (A.py:523-525, 1260-1262) for every cell of `grid`:
 * finite-difference b-vector (A.py:129-173) in shared memory over a TMA-staged grid tile and
 * the Lekien-Marsden solve alpha = inv(B) b (A.py:175, 577-579), by default fused into separable
 * 1-D line transforms on the FP64 pipe (HBM-write bound); the dense and Kronecker-factored FP64
 * tensor-core (DMMA) contractions are build variants 1-3 and 9, 4.  Writes (prod(n-3) + 1) * C * 4^d doubles to `table` (device), including the
 * NaN sentinel row.  For slab sharding pass the sub-grid of planes [lo-1, hi+2] of the
 * slowest axis: the build is local to the planes it is given.
 *   n[a] = grid points per axis (x first).  grid/table: device pointers. */
int arb_build_coeffs(int d, const double* grid, int ncomp, const int64_t n[4], double* table,
                     int reference_quirk, void* stream);
int arb_build_coeffs_3d(const double* grid, int ncomp, int64_t nx, int64_t ny, int64_t nz,
                        double* table, void* stream);
int arb_build_coeffs_4d(const double* grid, int ncomp, int64_t nx, int64_t ny, int64_t nz, int64_t nt,
                        double* table, void* stream);

/* Range query = rQuery1/2/3 (A.py:344-521, 1064-1258) on device-resident data.
 *   q          : device [N][ldq]; rows with a coordinate < IntMin or > IntMax are overwritten
 *                with NaN IN PLACE across all ldq columns (A.py:350-355).
 *   out_comps  : device [N][3]   (modes VECTOR, BOTH; else NULL)
 *   out_norm   : device [N]      (modes NORM, BOTH;   else NULL)
 *   out_grad   : device [N][d]   (modes NORM, BOTH;   else NULL), physical units (divided by h)
 *   out_cell   : device [N] int64 or NULL -- global cell index, `total cells` for masked/NaN rows
 *                (queryInds, A.py:368-370).
 *   masked_rows/masked_count : optional device list (capacity N) + counter that receive the row
 *                numbers NaN-masked in place (unordered); used by the host path to mirror the
 *                side effect without copying q back.  Counter must be zeroed by the caller.
 * A coordinate exactly on the upper edge whose per-axis index rounds to n-3 yields NaN
 * (declared deviation, DESIGN.md "upper edge"). */
int arb_query(const arb_geom* g, const double* table, int mode, double* q, int64_t N, int64_t ldq,
              double* out_comps, double* out_norm, double* out_grad, int64_t* out_cell,
              int64_t* masked_rows, unsigned long long* masked_count, void* stream);

/* Same call with HOST pointers for q and the outputs: chunks the batch, stages it through
 * pinned memory when the caller's buffers are pageable, and overlaps H2D / kernel / D2H on
 * internal streams.  Synchronous: returns when the outputs (and the NaN-masked q rows) are
 * complete in host memory.  `table` stays a device pointer.  out_cell_host may be NULL, a host
 * buffer, or a DEVICE buffer of N int64 (then the indices stay on the GPU and cost no PCIe
 * traffic; the Python classes read them back lazily for `queryInds`). */
int arb_query_host(const arb_geom* g, const double* table, int mode, double* q_host, int64_t N, int64_t ldq,
                   double* out_comps_host, double* out_norm_host, double* out_grad_host,
                   int64_t* out_cell_host, int64_t chunk_rows);

/* Table-free query: evaluates straight from the 4^d grid neighbourhood -- no coefficient table, no
 * build, 4^d x less memory; about half the throughput of arb_query.  It is the reference's lazy path
 * taken to its limit (A.py:376-377 computes a cell's coefficients on first touch; here nothing is ever
 * stored) and relies on A = M(x)M(x)M being exact in 3-D; in 4-D the reference matrix is M^(x)4 plus a
 * rank-16 term (A.py:860), which the kernel adds from the cell's 16 corner values of fxyzt.
 *   grid    : device [C][nt][nz][ny][pitch_x] float64, nx = ncell[0]+3 etc.; pitch_x even, >= nx.
 * Other arguments as arb_query / arb_query_host. */
int arb_query_grid(const arb_geom* g, const double* grid, int64_t pitch_x, int mode, double* q, int64_t N,
                   int64_t ldq, double* out_comps, double* out_norm, double* out_grad, int64_t* out_cell,
                   int64_t* masked_rows, unsigned long long* masked_count, void* stream);
int arb_query_grid_host(const arb_geom* g, const double* grid, int64_t pitch_x, int mode, double* q_host, int64_t N,
                        int64_t ldq, double* out_comps_host, double* out_norm_host, double* out_grad_host,
                        int64_t* out_cell_host, int64_t chunk_rows);

/* Node (Hermite) table -- a 4x (3-D) / 16x (4-D) smaller alternative to the cell coefficient table with the same
 * answers to round-off: the 2^d central-difference values (f, fx, fy, fxy, ... = the rows of the reference's D
 * matrix, A.py:129-173 / 762-876) of every interior grid node; a query evaluates the tensor-product cubic Hermite
 * interpolant of its 2^d corner nodes, which is the polynomial alpha = inv(B) D f describes (A.py:112-125, 175),
 * including the A.py:860 term in 4-D unless ARB_GEOM_FIXED_D4 is set.  Replaces calcCoefficients* + rQuery* together.
 *   grid  : device [ncomp][nt][nz][ny][pitch_x] as for arb_build_coeffs (pitch_x >= nx)
 *   nodes : device, 128-byte aligned.  d = 4: [ncomp][nt-2][nz-2][ny-2][nx-2][16];
 *           d = 3, ncomp = 1: [nz-2][ny-2][nx-3][2][8] -- x-adjacent nodes i, i+1 stored as aligned pairs (every node twice);
 *           d = 3, ncomp = 3 or 4 ('vector' / 'both'): [nz-2][ny-2][nx-2][4][8] -- the components of a node together
 *           (Bx, By, Bz, |B|; the 4th block is zero for ncomp = 3), so one gather serves all of them
 * arb_query_nodes / arb_query_nodes_host take the node table where arb_query / arb_query_host take the cell table;
 * every other argument, the outputs, the in-place NaN rows and the cell indices are the same.  No slabs. */
int arb_build_nodes(int d, const double* grid, int ncomp, const int64_t* npts, int64_t pitch_x, double* nodes,
                    void* stream);
int arb_query_nodes(const arb_geom* g, const double* nodes, int mode, double* q, int64_t N, int64_t ldq,
                    double* out_comps, double* out_norm, double* out_grad, int64_t* out_cell, int64_t* masked_rows,
                    unsigned long long* masked_count, void* stream);
int arb_query_nodes_host(const arb_geom* g, const double* nodes, int mode, double* q_host, int64_t N, int64_t ldq,
                         double* out_comps_host, double* out_norm_host, double* out_grad_host, int64_t* out_cell_host,
                         int64_t chunk_rows);

/* Table-free 'vector' / 'both' queries on a component-interleaved grid: packed = device [nz][ny][nx][4] (3-D) or
 * [nt][nz][ny][nx][4] (4-D; rQuery1 / rQuery3 of quadcubic, A.py:1064-1127, 1190-1258, incl. the A.py:860 term unless
 * ARB_GEOM_FIXED_D4) with the point's values (Bx, By, Bz, |B|) together -- the 4th value is not read in mode VECTOR --
 * 32-byte aligned, smaller than 128 GB; g->ncomp = 3 or 4 as for arb_query_grid.  One 128-byte piece per grid row of
 * the neighbourhood serves every component (arb_query_grid: one 48-byte piece per row and component).  Same outputs
 * and conventions as arb_query_grid. */
int arb_query_gridil(const arb_geom* g, const double* packed, int mode, double* q, int64_t N, int64_t ldq,
                     double* out_comps, double* out_norm, double* out_grad, int64_t* out_cell, int64_t* masked_rows,
                     unsigned long long* masked_count, void* stream);
int arb_query_gridil_host(const arb_geom* g, const double* packed, int mode, double* q_host, int64_t N, int64_t ldq,
                          double* out_comps_host, double* out_norm_host, double* out_grad_host, int64_t* out_cell_host,
                          int64_t chunk_rows);

/* Slab-sharded tables, the return leg fused into the query kernel: the N rows are the ones the other ranks sent to
 * this rank (the owner of their slab), in segments by sender -- rows [seg_start[h], seg_start[h+1]) came from rank h --
 * and row n's outputs are stored straight into rank h's result buffer, peer memory reached over NVLink / NVSwitch, at
 * result row home_row[n] (the row it has in the sender's batch).  No result all-to-all and no re-ordering pass follow
 * (rQuery1/2/3 over a table sharded by slabs, A.py:1064-1258 / 344-521; global cell index A.py:1088).
 *   seg_start : host int64 [npeers + 1], seg_start[0] = 0, seg_start[npeers] = N;  home_row : device int64 [N]
 *   peers     : host array of npeers device pointers (npeers <= ARB_MAX_PEERS), peers[h] = rank h's result buffer mapped
 *               into this process (symmetric memory / CUDA IPC), 16-byte aligned; result row = ld doubles:
 *               [comps(3, modes VECTOR/BOTH) | norm(1) grad(d) (modes NORM/BOTH) | cell index (int64 bits) | pad to even]
 * Rows outside this rank's slab or outside the volume store NaN outputs and the sentinel index like arb_query.  The
 * caller orders the kernel against the readers (a barrier across the ranks after the kernel). */
#define ARB_MAX_PEERS 16
int arb_query_routed(const arb_geom* g, const double* table, int mode, double* q, int64_t N, int64_t ldq,
                     const int64_t* seg_start, const int64_t* home_row, double* const* peers, int npeers, int64_t ld,
                     void* stream);
/* Both legs fused (no NCCL all-to-all, no sort, no host round trip on the query path of a slab-sharded table).
 * arb_route_rows: every row of this rank's batch q [n][ldq] is stored into the inbox of the rank that owns its slab
 * (owner as in arb_owner_keys), followed by its row number here: inbox row = [coords (d) | home row (int64 bits) | pad to
 * an even count] = 4 (d = 3) or 6 (d = 4) doubles.  Rank o's inbox is nslab segments of seg_cap rows (n <= seg_cap),
 * segment r written by rank r only; when the kernel ends counts[o][my_rank] on rank o holds the rows this rank sent there.
 *   inboxes / counts : host arrays of nslab device pointers, entry o = rank o's inbox / int64 [nslab] counts mapped into
 *                      this process (symmetric memory / CUDA IPC); cursor: device uint64 [nslab], ticket: device uint32,
 *                      both local and ZERO on entry; outside: optional device [n], 1 = a coordinate outside the volume.
 * arb_query_inbox: the owner's side -- evaluates the rows of its inbox (counts read on the device) and stores every row's
 * outputs into peers[sender] at the row's home row, exactly like arb_query_routed (same result row layout, same ld).
 * The caller orders route kernels, query kernels and readers with a barrier across the ranks after each kernel. */
int arb_route_rows(const arb_geom* g, const double* q, int64_t n, int64_t ldq, const int64_t* slab_hi, int nslab,
                   int my_rank, double* const* inboxes, int64_t* const* counts, int64_t seg_cap,
                   unsigned long long* cursor, unsigned int* ticket, unsigned char* outside, void* stream);
int arb_query_inbox(const arb_geom* g, const double* table, int mode, double* inbox, const int64_t* inbox_counts,
                    int64_t seg_cap, double* const* peers, int npeers, int64_t ld, void* stream);
/* Routing keys of a slab-sharded table: owner[n] = the rank whose slab [slab_hi[r-1], slab_hi[r]) holds row n's
 * slowest-axis cell layer floor((t - tIntMin) / ht) (A.py:1081-1086; rows without a layer go to rank 0), outside[n] = 1
 * when a coordinate lies outside the interpolation volume (A.py:1069-1076).  q: device [n][ldq]; g->slab_* are ignored. */
int arb_owner_keys(const arb_geom* g, const double* q, int64_t n, int64_t ldq, const int64_t* slab_hi, int nslab,
                   int16_t* owner, unsigned char* outside, void* stream);
/* cudaDeviceEnablePeerAccess(peer_device) for the current device; "already enabled" is success. */
int arb_enable_peer_access(int peer_device);

/* Fused query + push: nsteps velocity-Verlet steps of dv/dt = kappa * grad(value)(x) + gravity for N
 * particles resident in device memory, the gradient being what Query2/Query3 return (A.py:452, 519)
 * for the table's last (norm / scalar) component; mode NORM or BOTH.  For d = 4 (time-dependent field) pos
 * carries each particle's own time as 4th coordinate, which advances by dt per step.
 *   pos : device [N][d], vel : device [N][3], updated in place; particles that leave the interpolation volume get NaN
 *              position and velocity (the Query convention) and are counted in *lost_count (device, may be NULL).
 *   gravity  : host pointer to 3 doubles or NULL. */
int arb_push(const arb_geom* g, const double* table, int mode, double* pos, double* vel, int64_t N, double dt,
             int64_t nsteps, double kappa, const double* gravity, unsigned long long* lost_count, void* stream);
/* Resumable form for slab-sharded tables (g->slab_lo/hi): step_io (device [N] int64, may be NULL = arb_push) holds
 * for every particle the index of the next step whose force has to be evaluated -- 0 for a fresh particle, a value
 * > nsteps for one that is finished.  A particle that is inside the volume but outside the table's slab is parked:
 * pos, vel and step_io are left as they are, so the rank owning that slab resumes it with the arithmetic of an
 * unsharded run.  On return step_io[n] = nsteps + 1 for finished or lost particles. */
int arb_push_steps(const arb_geom* g, const double* table, int mode, double* pos, double* vel, int64_t* step_io,
                   int64_t N, double dt, int64_t nsteps, double kappa, const double* gravity,
                   unsigned long long* lost_count, void* stream);
/* arb_push on a node table (arb_build_nodes) instead of the cell table; no slabs (a node table is replicated). */
int arb_push_nodes(const arb_geom* g, const double* nodes, int mode, double* pos, double* vel, int64_t N, double dt,
                   int64_t nsteps, double kappa, const double* gravity, unsigned long long* lost_count, void* stream);

/* Row permutation used by the slab-sharded routing path (no counterpart in the single-process reference):
 * gather  (scatter == 0): dst[i][:] = src[order[i]][:];  scatter (scatter != 0): dst[order[i]][:] = src[i][:].
 * Rows are `width` doubles; all pointers are device pointers. */
int arb_permute_rows(double* dst, const double* src, const int64_t* order, int64_t n, int width, int scatter,
                     void* stream);

/* Tuning knob for experiments/benchmarks: selects the query-kernel variant
 * (0 = default).  Returns the previous value. */
int arb_set_query_variant(int variant);
/* Same for the build kernel: 0 = default (separable FP64-pipe kernel), 5-8 its other tile / march
 * configurations, 1-3 dense DMMA contraction, 9 and 4 Kronecker-factored DMMA solve. */
int arb_set_build_variant(int variant);

#ifdef __cplusplus
}
#endif
#endif /* ARBINTERP_B200_H */
