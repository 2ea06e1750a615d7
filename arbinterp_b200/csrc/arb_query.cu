// Query kernels: locate cell -> gather the cell's coefficient block(s) -> Horner evaluation of
// the value and the analytic partial derivatives.  Replaces rQuery1/2/3 of the reference
// (A.py:344-521 tricubic, A.py:1064-1258 quadcubic).
//
// The path is HBM-bound (one 512 B / 2 KB coefficient block per interpolated component per
// query, no reuse for random queries), so the kernels are organised around keeping many
// full-line reads in flight per SM; the FP64 arithmetic (about 70-150 DFMA per query) is
// 10-20 % of the issue budget at roofline.
//
// Variants (arb_set_query_variant; everything but 0 is kept as a measured baseline, see
// profiles/r01_variant_sweep.log):
//   0  BLOCK: default, see query_block_kernel below (TMA block gather, one lane per 64-coefficient
//             block, all modes and both dimensionalities).  20-23: the same with other CTA sizes /
//             without warp de-duplication; 30: components looped inside a work item (the 4-D default).
//   10/11 COOP : G = 8 (3-D) / 32 (4-D) lanes share one query; lane l issues four 16-byte loads
//             that are contiguous across the group (one full 128 B line per group and
//             instruction), evaluates its 8 coefficients and the partial sums are combined
//             with a reduce-scatter over shuffles.  No shared memory, any block size.
//   1  BULK : first TMA design: one query per thread; each thread asks the TMA engine (cp.async.bulk, SASS
//             UBLKCP) to copy its cell's block into a private shared-memory slot, all copies
//             of a warp complete on one mbarrier, then the thread streams its slot through a
//             nested Horner scheme (LDS.128, conflict-free by a 16-byte slot skew).
//   2  DIRECT: one query per thread reading its block straight from global memory
//             (32 uncoalesced LDG.128 per component); kept as the naive baseline.
#include <atomic>
#include <mutex>
#include "arb_device.cuh"
#include "arb_gridfree.cuh"
#include "arb_nodes.cuh"

namespace arb {

// debugging / measurement knob (arb_set_query_variant); atomic so that concurrent callers at least see a whole value
static std::atomic<int> g_query_variant{0};
int current_query_variant() { return g_query_variant.load(std::memory_order_relaxed); }

template <int MODE, int D> struct OutCount { static constexpr int NV = MODE == 0 ? 3 : (MODE == 1 ? 1 + D : 4 + D); };
constexpr int next_pow2(int v) { return v <= 1 ? 1 : (v <= 2 ? 2 : (v <= 4 ? 4 : (v <= 8 ? 8 : 16))); }

// store output number idx of query n (value or NaN)
template <int D, int MODE>
__device__ __forceinline__ void store_output(const QueryParams& p, int64_t n, int idx, double v) {
    if (MODE == 0) {
        if (idx < 3) p.out_comps[n * 3 + idx] = v;
    } else if (MODE == 1) {
        if (idx == 0) p.out_norm[n] = v;
        else if (idx <= D) p.out_grad[n * D + (idx - 1)] = __ddiv_rn(v, p.h[idx - 1]);
    } else {
        if (idx < 3) p.out_comps[n * 3 + idx] = v;
        else if (idx == 3) p.out_norm[n] = v;
        else if (idx < 4 + D) p.out_grad[n * D + (idx - 4)] = __ddiv_rn(v, p.h[idx - 4]);
    }
}

// ======================================================================================
// Variants 10/11: cooperative sub-warp gather
// ======================================================================================
template <int D, int C, int MODE, int U>
__global__ void __launch_bounds__(256) query_coop_kernel(const QueryParams p) {
    constexpr int G = (D == 3) ? 8 : 32;       // lanes per query
    constexpr int NM = (D == 3) ? 64 : 256;    // coefficients per component
    constexpr int RS = NM / 4;                 // doubles between slow-index blocks
    constexpr int QPW = 32 / G;                // queries per warp and pass
    constexpr int NV = OutCount<MODE, D>::NV;
    constexpr int NVP = next_pow2(NV);
    static_assert(NVP <= G, "reduce-scatter needs at least NVP lanes");

    const int lane = threadIdx.x & 31;
    const int l = lane & (G - 1);
    const int grp = lane / G;
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;

    const int ju = l & 1;                 // lane holds i = 2*ju + {0,1}
    const int jv = (l >> 1) & 3;          // fixed j
    const int jw = (l >> 3) & 3;          // fixed k (4-D only)

    for (int64_t base = warp_global * (QPW * U); base < p.N; base += nwarps * (QPW * U)) {
        Located<D> L[U];
        int64_t n[U];
        double2 a[U][C][4];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            n[u] = base + u * QPW + grp;
            if (n[u] < p.N) {
                L[u] = locate<D>(p, n[u]);
            } else {
                L[u].ok = false; L[u].masked = false; L[u].cell_global = 0; L[u].cell_local = 0;
#pragma unroll
                for (int d_ = 0; d_ < D; ++d_) L[u].frac[d_] = 0.0;
            }
        }
        // issue every load of the pass before the first use
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const double* src = p.table + (L[u].cell_local * C) * NM + 2 * l;
#pragma unroll
            for (int c = 0; c < C; ++c)
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    if (L[u].ok) a[u][c][r] = ldg_stream_d2(src + c * NM + r * RS);
                    else a[u][c][r] = make_double2(0.0, 0.0);
                }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const double fu = L[u].frac[0], fv = L[u].frac[1], fw = L[u].frac[2];
            const double s = L[u].frac[D - 1];   // slow variable: Horner over the four loads
            const double u2 = fu * fu;
            const double pu0 = ju ? u2 : 1.0, pu1 = ju ? u2 * fu : fu;
            const double dpu0 = ju ? 2.0 * fu : 0.0, dpu1 = ju ? 3.0 * u2 : 1.0;
            const double pv = pow_sel(fv, jv), dpv = dpow_sel(fv, jv);
            double Pm = pv, Pm_dy = dpv, Pm_dz = 0.0;
            if (D == 4) {
                const double pw = pow_sel(fw, jw), dpw = dpow_sel(fw, jw);
                Pm = pv * pw; Pm_dy = dpv * pw; Pm_dz = pv * dpw;
            }
            double v[NVP];
#pragma unroll
            for (int i = 0; i < NVP; ++i) v[i] = 0.0;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                double t[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) t[r] = fma(a[u][c][r].y, pu1, a[u][c][r].x * pu0);
                const double T = fma(fma(fma(t[3], s, t[2]), s, t[1]), s, t[0]);
                const bool grad_comp = (MODE == 1) || (MODE == 2 && c == 3);
                if (!grad_comp) {
                    v[c] = T * Pm;
                } else {
                    constexpr int o = (MODE == 1) ? 0 : 3;
                    double tx[4];
#pragma unroll
                    for (int r = 0; r < 4; ++r) tx[r] = fma(a[u][c][r].y, dpu1, a[u][c][r].x * dpu0);
                    const double Tx = fma(fma(fma(tx[3], s, tx[2]), s, tx[1]), s, tx[0]);
                    const double Ts = fma(fma(3.0 * t[3], s, 2.0 * t[2]), s, t[1]);
                    v[o] = T * Pm;
                    v[o + 1] = Tx * Pm;
                    v[o + 2] = T * Pm_dy;
                    if (D == 4) v[o + 3] = T * Pm_dz;
                    v[o + D] = Ts * Pm;
                }
            }
            // reduce-scatter over the group: after step b a lane keeps half of its list
            int idx = 0;
#pragma unroll
            for (int b = 0, cnt = NVP; cnt > 1; ++b, cnt >>= 1) {
                const int half = cnt >> 1;
                const bool up = (l >> b) & 1;
#pragma unroll
                for (int i = 0; i < half; ++i) {
                    const double keep = up ? v[half + i] : v[i];
                    const double send = up ? v[i] : v[half + i];
                    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 1 << b);
                }
                idx += up ? half : 0;
            }
#pragma unroll
            for (int m = NVP; m < G; m <<= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], m);
            if (n[u] < p.N) {
                if (l < NVP && idx < NV) store_output<D, MODE>(p, n[u], idx, L[u].ok ? v[0] : qnan());
                if (l == G - 1) {
                    if (p.out_cell) p.out_cell[n[u]] = L[u].cell_global;
                    if (L[u].masked) mask_row_in_place(p, n[u]);
                }
            }
        }
    }
}

template <int D, int C, int MODE, bool SHARED>
__device__ __forceinline__ void eval_and_store(const QueryParams& p, int64_t n, const Located<D>& L,
                                               const double* blk) {
    constexpr int NM = (D == 3) ? 64 : 256;
    if (!L.ok) {
#pragma unroll
        for (int i = 0; i < OutCount<MODE, D>::NV; ++i) store_output<D, MODE>(p, n, i, qnan());
        return;
    }
    if (MODE == 0 || MODE == 2) {
#pragma unroll
        for (int c = 0; c < 3; ++c) store_output<D, MODE>(p, n, c, eval_value<D, SHARED>(blk + c * NM, L.frac));
    }
    if (MODE == 1 || MODE == 2) {
        constexpr int o = (MODE == 1) ? 0 : 3;
        double g[5];
        eval_value_grad<D, SHARED>(blk + (C - 1) * NM, L.frac, g);
#pragma unroll
        for (int i = 0; i <= D; ++i) store_output<D, MODE>(p, n, o + i, g[i]);
    }
}

// ======================================================================================
// Variant 1: TMA bulk gather into per-thread shared-memory slots
// ======================================================================================
template <int D, int C, int MODE, int THREADS>
__global__ void __launch_bounds__(THREADS) query_bulk_kernel(const QueryParams p) {
    constexpr int NM = (D == 3) ? 64 : 256;
    constexpr uint32_t BYTES = C * NM * 8;
    constexpr uint32_t SLOT = BYTES + 16;     // 16 B skew: LDS.128 of a quarter warp hits 8 distinct 16 B bank groups
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bars[THREADS / 32];

    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned char* slot = smem + (size_t)threadIdx.x * SLOT;
    uint64_t* bar = &bars[wid];
    if (lane == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncwarp();
    const int64_t warp_global = ((int64_t)blockIdx.x * THREADS + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * THREADS) >> 5;
    uint32_t phase = 0;
    for (int64_t base = warp_global * 32; base < p.N; base += nwarps * 32) {
        const int64_t n = base + lane;
        Located<D> L;
        L.ok = false; L.masked = false; L.cell_global = 0; L.cell_local = 0;
        if (n < p.N) L = locate<D>(p, n);
        const unsigned okmask = __ballot_sync(0xffffffffu, L.ok);
        if (lane == 0) mbar_expect_tx(bar, (uint32_t)__popc(okmask) * BYTES);
        __syncwarp();
        if (L.ok) bulk_g2s(slot, p.table + L.cell_local * (int64_t)(C * NM), BYTES, bar);
        if (n < p.N) {
            if (p.out_cell) p.out_cell[n] = L.cell_global;
            if (L.masked) mask_row_in_place(p, n);
        }
        mbar_wait(bar, phase);
        phase ^= 1;
        if (n < p.N) eval_and_store<D, C, MODE, true>(p, n, L, reinterpret_cast<const double*>(slot));
        __syncwarp();
    }
}

// ======================================================================================
// Variant 0 (default): TMA block gather, one lane per 64-coefficient block
// ======================================================================================
// Work item = (batch of queries, component).  In 3-D a lane owns (query, component); in 4-D the
// 256-coefficient block of a component is four tricubic blocks alpha[.., l], l = 0..3, and four
// adjacent lanes own (query, component, l): each evaluates its tricubic block in (u, v, w) and the
// four results are combined with weights s^l (value, d/dx, d/dy, d/dz) and l s^(l-1) (d/dt) over
// two shuffle steps.  Every lane therefore always gathers exactly one 512-byte block with one
// cp.async.bulk (UBLKCP) into its own 528-byte shared-memory slot, whatever the mode: 128-thread
// CTAs use 66 KB, three are resident per SM and up to 192 KB of coefficient reads are in flight
// per SM.  All copies of a warp complete on the warp's own mbarrier, so warps never wait for each
// other.  DEDUP: lanes of a warp that want the same block elect one leader to fetch it and read
// the leader's slot (warp-level binning by cell; free for clustered queries such as particle
// bunches, one MATCH instruction of overhead for random ones).
// SLOTS < 32 (clustered batches: cell-sorted rows, particle bunches, trajectories): the distinct blocks a warp item
// needs are numbered 0..K-1 and land in a ring of only SLOTS slots per warp (pass p fetches and evaluates the lanes
// whose block number is in [p SLOTS, (p+1) SLOTS)), so a warp costs SLOTS x 528 B of shared memory instead of 16.5 KB
// and 2-3 times as many warps are resident.  A clustered item has K <= SLOTS and takes one pass with every lane
// evaluating; what limits it is not DRAM any more but the latency of the locate -> fetch -> evaluate chain per warp,
// which the extra warps hide.  Uniformly random rows (K = 32) take 32 / SLOTS half-empty passes -- the launcher picks
// SLOTS from a sortedness probe of the batch (probe_kernel).
// PREFETCH: the coordinates of the warp's NEXT item are loaded while the current item's blocks are in flight.
// FETCH_LDGSTS: a slot is filled by ONE warp-wide cp.async (32 lanes x 16 bytes, SASS LDGSTS) whose addresses come from
// the owning lane over shuffles, instead of a per-lane cp.async.bulk -- the bulk copies take uniform registers, so
// the compiler issues them in a serial ELECT / R2UR / UBLKCP loop (one trip per lane), while the LDGSTS loop has no
// chain between its trips; it is also the only form that can gather a slot from several segments.
// KIND_NODES: `table` is a node (Hermite) table (arb_nodes.cuh): a 3-D slot is the 4 x-pairs of corner nodes
// (4 segments of 128 B, stored as 128-byte-aligned pairs), a 4-D slot the 2 x-pairs (cy = 0, 1) of one (cz, ct) (2 segments of 256 B); the four lanes
// of a 4-D query own (cz, ct) and their shares add up.  QUIRK4 reproduces A.py:860 there.
// KIND_NODES_IL / KIND_GRID_IL (3-D, modes 'vector' and 'both'): the components are interleaved in memory, so one
// gather serves all of them and its pieces are 512 / 128 contiguous bytes instead of 128 / 32.  Four lanes own a
// query, one 512-byte slot each, and their shares of every output add up over two shuffle steps:
//   NODES_IL: node table [nz-2][ny-2][nx-2][4][8]; lane (cy, cz) holds the x-pair of nodes of its row (one segment);
//   GRID_IL : raw grid   [nz][ny][nx][4] (table-free); lane k holds z-plane k of the 4x4x4 neighbourhood (4 segments
//             of 128 B = the four x-neighbours of all components).
constexpr int KIND_CELLS = 0, KIND_NODES = 1, KIND_NODES_IL = 2, KIND_GRID_IL = 3;
// ROUTED (slab-sharded tables): the rows are the ones other ranks sent here, in segments by sender; a row's outputs are
// stored into the SENDER's result buffer -- peer memory over NVLink -- at the row it has in the sender's batch
// (home_row[n]).  The outputs of a warp item are assembled in shared memory (the slot ring is free by then) and every
// result row leaves as 16-byte pieces stored by adjacent lanes, because small scattered remote stores are what NVLink
// is worst at (one 8-byte store per output: the kernel took twice as long at 8 ranks).
template <int D, int MODE, int THREADS, bool DEDUP, bool LOOPC = false, int SLOTS = 32, bool PREFETCH = false,
          bool FETCH_LDGSTS = false, int KIND = KIND_CELLS, bool QUIRK4 = true, bool ROUTED = false>
__global__ void __launch_bounds__(THREADS) query_block_kernel(const QueryParams p, const int* __restrict__ gate, int gate_want) {
    static_assert(KIND == KIND_CELLS || FETCH_LDGSTS, "node slots are gathered from several segments");
    constexpr bool QUAD = (KIND == KIND_NODES_IL || KIND == KIND_GRID_IL);
    static_assert(!QUAD || (D == 3 && MODE != 1), "interleaved layouts: 3-D 'vector' / 'both'");
    constexpr int C = QUAD ? 1 : (MODE == 0 ? 3 : (MODE == 1 ? 1 : 4));      // separately fetched components
    constexpr int SL = (D == 4 || QUAD) ? 4 : 1;  // 512-byte slots (lanes) per query and component
    constexpr int QPW = 32 / SL;                  // queries per warp item
    constexpr uint32_t BYTES = 512;
    constexpr uint32_t SLOT = BYTES + 16;         // skew keeps LDS.128 conflict-free across lanes
    static_assert(SLOTS == 32 || DEDUP, "the compact slot ring numbers the de-duplicated blocks");
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bars[THREADS / 32];
    if (gate && *gate != gate_want) return;       // two launches, the probe's verdict picks the one that runs

    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int qi = lane / SL, sl = lane % SL;
    uint64_t* bar = &bars[wid];
    if (lane == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncwarp();
    unsigned char* const ring = smem + (size_t)wid * SLOTS * SLOT;
    // LDGSTS: where this lane's 16 bytes of a slot come from, relative to the slot's first byte in global memory
    int64_t lane_src = lane * 16;
    if (KIND == KIND_NODES) {
        if (D == 3) lane_src = (((lane >> 4) * p.nn[1] + ((lane >> 3) & 1)) * p.nc[0]) * 128 + (lane & 7) * 16;
        else lane_src = (lane >> 4) * p.nn[0] * 128 + (lane & 15) * 16;
    }
    if (KIND == KIND_GRID_IL) lane_src = (lane >> 3) * (p.nc[0] + 3) * 32 + (lane & 7) * 16;      // row j = lane / 8
    const int64_t warp_global = ((int64_t)blockIdx.x * THREADS + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * THREADS) >> 5;
    const int64_t nbatch = (p.N + QPW - 1) / QPW;
    // LOOPC: a work item is a batch and its components are fetched one after the other (one locate per
    // query); otherwise (batch, component) pairs are separate items (more independent items in flight).
    int64_t nitem = LOOPC ? nbatch : nbatch * C;
    // inbox form: the rows sit in npeers segments of seg_cap rows, inbox_counts[h] of them filled; items are numbered
    // over the filled parts only (a segment's last item may be partly empty)
    // (kept in shared memory: per-thread arrays indexed by the segment went to local memory and, with the per-item
    // reads of the counts from global memory, cost the inbox form ~5 % against the plain kernel)
    __shared__ int64_t seg_item0[ROUTED ? ARB_MAX_PEERS + 1 : 1], seg_rows[ROUTED ? ARB_MAX_PEERS : 1];
    const bool inbox = ROUTED && p.inbox_counts != nullptr;
    if (ROUTED && inbox) {
        if (threadIdx.x == 0) {
            int64_t run = 0;
            for (int h = 0; h < p.npeers; ++h) {
                const int64_t c = p.inbox_counts[h];
                seg_rows[ROUTED ? h : 0] = c;
                seg_item0[ROUTED ? h : 0] = run;
                run += (c + QPW - 1) / QPW;
            }
            seg_item0[ROUTED ? p.npeers : 0] = run;
        }
        __syncthreads();
        nitem = seg_item0[ROUTED ? p.npeers : 0];
    }
    uint32_t phase = 0;
    double cnext[D];
    if (PREFETCH) {
        const int64_t n0 = (LOOPC ? warp_global : warp_global / C) * QPW + qi;
#pragma unroll
        for (int a = 0; a < D; ++a) cnext[a] = (warp_global < nitem && n0 < p.N) ? p.q[n0 * p.ldq + a] : 0.0;
    }
    for (int64_t item = warp_global; item < nitem; item += nwarps) {
        const int64_t batch = LOOPC ? item : item / C;
        int64_t n = batch * QPW + qi;
        int seg = 0;
        if (ROUTED && inbox) {                       // item -> (segment, row of the segment); rows past the count: n = N
            for (int h = 1; h < p.npeers; ++h) seg = (item >= seg_item0[ROUTED ? h : 0]) ? h : seg;
            const int64_t nl = (item - seg_item0[ROUTED ? seg : 0]) * QPW + qi;
            n = (nl < seg_rows[ROUTED ? seg : 0]) ? seg * p.seg_cap + nl : p.N;
        }
        Located<D> L;
        L.ok = false; L.masked = false; L.cell_global = 0; L.cell_local = 0;
        if (PREFETCH) {
            if (n < p.N) L = locate_coords<D>(p, cnext);
        } else {
            if (n < p.N) L = locate<D>(p, n);
        }
        // where this row's outputs go
        double *o_comps = nullptr, *o_norm = nullptr, *o_grad = nullptr;
        int64_t* o_cell = nullptr;
        constexpr int OFFN = (MODE == 1) ? 0 : 3;                           // routed result row: [comps | norm grad | cell | pad]
        constexpr int OFFC = (MODE == 0) ? 3 : OFFN + 1 + D;
        constexpr int LDS = (OFFC + 2) / 2 * 2;
        double rowv[ROUTED ? LDS : 1];
        if (ROUTED) {
            static_assert(!ROUTED || (LOOPC && !QUAD), "routed rows are assembled per query: components looped in the item");
#pragma unroll
            for (int i = 0; i < (ROUTED ? LDS : 1); ++i) rowv[i] = 0.0;
        } else if (n < p.N) {
            o_comps = p.out_comps + n * 3; o_norm = p.out_norm + n; o_grad = p.out_grad + n * D;
            o_cell = p.out_cell ? p.out_cell + n : nullptr;
        }
#pragma unroll 1
      for (int ci = 0; ci < (LOOPC ? C : 1); ++ci) {
        const int comp = LOOPC ? ci : (int)(item - batch * C);
        const int64_t blk = (L.cell_local * C + comp) * SL + sl;      // 512-byte block number in the table
        int src_lane = lane;
        bool leader = L.ok;
        if (DEDUP) {
            const unsigned peers = __match_any_sync(0xffffffffu, L.ok ? blk : (int64_t)(-1 - lane));
            src_lane = __ffs(peers) - 1;
            leader = L.ok && (src_lane == lane);
        }
        int myslot = src_lane, mypass = 0, npass = 1;
        if (SLOTS < 32) {
            const unsigned lm = __ballot_sync(0xffffffffu, leader);
            const int number = __popc(lm & ((1u << lane) - 1u));          // leaders: their block's number
            const int mine = __shfl_sync(0xffffffffu, number, src_lane);
            mypass = L.ok ? mine / SLOTS : 0;
            myslot = mine % SLOTS;
            npass = (__popc(lm) + SLOTS - 1) / SLOTS;
            if (npass < 1) npass = 1;
        }
        if (comp == 0 && sl == 0 && n < p.N) {
            if (ROUTED) rowv[ROUTED ? OFFC : 0] = __longlong_as_double(L.cell_global);
            else if (o_cell) *o_cell = L.cell_global;
            if (L.masked && !ROUTED) mask_row_in_place(p, n);      // routed rows are copies: the home rank masks the caller's
        }
        const bool grad_comp = (MODE == 1) || (MODE == 2 && (QUAD || comp == 3));     // warp-uniform
        double g[QUAD ? 7 : 5];
#pragma unroll
        for (int i = 0; i < (QUAD ? 7 : 5); ++i) g[i] = 0.0;
        double f15[4] = {0.0, 0.0, 0.0, 0.0};                               // 4-D nodes: fxyzt of this lane's corners
        const double* cb = reinterpret_cast<const double*>(ring + (size_t)myslot * SLOT);
        const double* src = p.table + blk * 64;
        if (KIND == KIND_NODES) {
            if (D == 3) {
                src = p.table + comp * p.node_comp_stride + ((L.idx[2] * p.nn[1] + L.idx[1]) * p.nc[0] + L.idx[0]) * 16;
            } else {
                const int64_t iz = L.idx[2] + (sl & 1), it = L.idx[D - 1] + (sl >> 1);
                src = p.table + comp * p.node_comp_stride +
                      (((it * p.nn[2] + iz) * p.nn[1] + L.idx[1]) * p.nn[0] + L.idx[0]) * 16;
            }
        }
        if (KIND == KIND_NODES_IL)      // lane (cy, cz) = (sl & 1, sl >> 1): nodes (ix, ix + 1) of its row, 2 x 256 B
            src = p.table + (((L.idx[2] + (sl >> 1)) * p.nn[1] + L.idx[1] + (sl & 1)) * p.nn[0] + L.idx[0]) * 32;
        if (KIND == KIND_GRID_IL)       // lane k = sl: rows (iy .. iy + 3) of grid plane iz + k, 4 x-points x 4 components each
            src = p.table + (((L.idx[2] + sl) * (p.nc[1] + 3) + L.idx[1]) * (p.nc[0] + 3) + L.idx[0]) * 4;
#pragma unroll 1
        for (int pass = 0; pass < npass; ++pass) {
            const bool active = (SLOTS == 32) || (mypass == pass);
            const bool fetch = leader && active;
            const unsigned fmask = __ballot_sync(0xffffffffu, fetch);
            if (FETCH_LDGSTS) {
                // one shuffle per slot: the owner's source as a 32-bit count of 2^USH-byte units from the table base (every
                // slot starts on such a unit: cell blocks 512 B, node pairs 128 / 256 B, interleaved grid points 32 B --
                // good for 2 TB / 512 GB / 128 GB of table), the destination packed beside it when the ring is compact;
                // the 32 trips are unrolled and independent of each other (warp-uniform predicate from the ballot)
                constexpr int USH = (KIND == KIND_CELLS) ? 9 : (KIND == KIND_GRID_IL ? 5 : 7);
                const uint32_t src16 = (uint32_t)((reinterpret_cast<const char*>(src) - reinterpret_cast<const char*>(p.table)) >> USH);
                const char* const lane_base = reinterpret_cast<const char*>(p.table) + lane_src;
                const uint32_t ring0 = smem_u32(ring) + lane * 16;
                // (interleaved nodes in mode 'vector': leaving the lanes idle that copy the unread |B| block of each node
                // was measured -- DRAM still moves whole 128-byte lines, 2065 bytes per query either way,
                // profiles/r02_nodes3d_vector_ncu.txt -- and dropped)
#pragma unroll 8
                for (int o = 0; o < 32; ++o) {
                    if ((fmask >> o) & 1u) {
                        const uint32_t s16 = __shfl_sync(0xffffffffu, src16, o);
                        const uint32_t slot_o = (SLOTS == 32) ? (uint32_t)o : (uint32_t)__shfl_sync(0xffffffffu, myslot, o);
                        cp_async_16(ring0 + slot_o * SLOT, lane_base + ((size_t)s16 << USH));
                    }
                }
            } else {
                if (lane == 0) mbar_expect_tx(bar, (uint32_t)__popc(fmask) * BYTES);
                __syncwarp();
                if (fetch) bulk_g2s(ring + (size_t)myslot * SLOT, src, BYTES, bar);
            }
            if (PREFETCH && pass == 0 && ci == 0) {       // next item's coordinates: in flight during the wait
                const int64_t nx_item = item + nwarps;
                const int64_t n1 = (LOOPC ? nx_item : nx_item / C) * QPW + qi;
                if (nx_item < nitem && n1 < p.N) {
#pragma unroll
                    for (int a = 0; a < D; ++a) cnext[a] = p.q[n1 * p.ldq + a];
                }
            }
            if (FETCH_LDGSTS) {
                cp_async_wait_all();
                __syncwarp();
            } else {
                mbar_wait(bar, phase);
                phase ^= 1;
            }
            if (L.ok && active) {
                if constexpr (KIND == KIND_NODES_IL) {
                    nodes::eval3_row<MODE == 2>(cb, sl & 1, sl >> 1, L.frac, g);
                } else if constexpr (KIND == KIND_GRID_IL) {
                    gridfree::plane_il<MODE == 2>(cb, sl, L.frac, g);
                } else if (KIND == KIND_NODES) {
                    if (D == 3) {
                        if (grad_comp) nodes::eval3<true>(cb, L.frac, g);
                        else nodes::eval3<false>(cb, L.frac, g);
                    } else {
                        if (grad_comp) nodes::eval4_lane<true, false>(cb, sl & 1, sl >> 1, L.frac, 0.0, g);
                        else nodes::eval4_lane<false, false>(cb, sl & 1, sl >> 1, L.frac, 0.0, g);
                        if (QUIRK4) { f15[0] = cb[15]; f15[1] = cb[31]; f15[2] = cb[47]; f15[3] = cb[63]; }
                    }
                } else {
                    if (grad_comp) eval_value_grad<3, true>(cb, L.frac, g);
                    else g[0] = eval_value<3, true>(cb, L.frac);
                }
            }
            __syncwarp();   // every lane is done with the slots before the next copies land
        }
        if constexpr (QUAD) {
#pragma unroll
            for (int i = 0; i < (MODE == 2 ? 7 : 3); ++i) {
                g[i] += __shfl_xor_sync(0xffffffffu, g[i], 1);
                g[i] += __shfl_xor_sync(0xffffffffu, g[i], 2);
            }
        } else if (D == 4 && KIND == KIND_NODES) {
            if (QUIRK4) {       // A.py:860: the fxyzt slot of corner c holds fxyzt(c - 1); the lane before owns c - 1 of our first corner
                const double up = __shfl_up_sync(0xffffffffu, f15[3], 1);
                if (grad_comp) nodes::quirk4_lane<true>(f15, sl ? up : 0.0, sl & 1, sl >> 1, L.frac, g);
                else nodes::quirk4_lane<false>(f15, sl ? up : 0.0, sl & 1, sl >> 1, L.frac, g);
            }
            const int nred = grad_comp ? 5 : 1;
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                if (i < nred) {
                    g[i] += __shfl_xor_sync(0xffffffffu, g[i], 1);
                    g[i] += __shfl_xor_sync(0xffffffffu, g[i], 2);
                }
            }
        } else if (D == 4) {
            const double s = L.frac[D - 1];
            const double w = pow_sel(s, sl), dw = dpow_sel(s, sl);
            g[4] = g[0] * dw;
#pragma unroll
            for (int i = 0; i < 4; ++i) g[i] *= w;
            const int nred = grad_comp ? 5 : 1;
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                if (i < nred) {
                    g[i] += __shfl_xor_sync(0xffffffffu, g[i], 1);
                    g[i] += __shfl_xor_sync(0xffffffffu, g[i], 2);
                }
            }
        }
        if constexpr (QUAD) {
            if (n < p.N && sl == 0) {
                const double nan = qnan();
#pragma unroll
                for (int c = 0; c < 3; ++c) o_comps[c] = L.ok ? g[c] : nan;
                if (MODE == 2) {
                    *o_norm = L.ok ? g[3] : nan;
#pragma unroll
                    for (int a = 0; a < 3; ++a) o_grad[a] = L.ok ? __ddiv_rn(g[4 + a], p.h[a]) : nan;
                }
            }
        } else if (ROUTED) {
            const double nan = qnan();
            if (!grad_comp) {
                const double v = L.ok ? g[0] : nan;
                if (comp == 0) rowv[0] = v;
                else if (comp == 1) rowv[ROUTED ? 1 : 0] = v;
                else rowv[ROUTED ? 2 : 0] = v;
            } else {
                rowv[ROUTED ? OFFN : 0] = L.ok ? g[0] : nan;
#pragma unroll
                for (int a = 0; a < D; ++a) rowv[ROUTED ? OFFN + 1 + a : 0] = L.ok ? __ddiv_rn(g[1 + a], p.h[a]) : nan;
            }
        } else if (n < p.N && sl == 0) {
            const double nan = qnan();
            if (!grad_comp) {
                o_comps[comp] = L.ok ? g[0] : nan;
            } else {
                *o_norm = L.ok ? g[0] : nan;
#pragma unroll
                for (int a = 0; a < D; ++a) o_grad[a] = L.ok ? __ddiv_rn(g[1 + a], p.h[a]) : nan;
            }
        }
      }
      if constexpr (ROUTED) {
        // assemble the item's result rows in shared memory, then store them with 16-byte pieces contiguous across lanes
        double* stage = reinterpret_cast<double*>(ring);
        if (sl == 0 && n < p.N) {
#pragma unroll
            for (int i = 0; i < LDS; ++i) stage[qi * LDS + i] = rowv[i];
        }
        unsigned long long dst = 0;
        if (n < p.N) {
            if (inbox) {                             // the home row number travelled with the row
                const int64_t home = __double_as_longlong(p.q[n * p.ldq + D]);
                dst = reinterpret_cast<unsigned long long>(p.peer[seg] + home * LDS);
            } else {
                int h = 0;
                for (int r = 1; r < p.npeers; ++r) h = (n >= p.seg_start[r]) ? r : h;
                dst = reinterpret_cast<unsigned long long>(p.peer[h] + p.home_row[n] * LDS);
            }
        }
        __syncwarp();
        constexpr int CPR = LDS / 2, NCH = QPW * CPR;
#pragma unroll
        for (int c0 = 0; c0 < NCH; c0 += 32) {
            const int c = c0 + lane;
            const bool valid = c < NCH;
            const int r = valid ? c / CPR : 0, part = c - r * CPR;
            const unsigned long long d0 = __shfl_sync(0xffffffffu, dst, r * SL);
            if (valid && d0) {
                const double2 v = *reinterpret_cast<const double2*>(stage + r * LDS + part * 2);
                stg_stream_d2(reinterpret_cast<double*>(d0) + part * 2, v.x, v.y);
            }
        }
        fence_proxy_async();   // generic-proxy writes to the ring above, async-proxy (TMA) writes to it next
        __syncwarp();          // the staging area is the slot ring: the next item's copies must not land before it is read
      }
    }
}

// Sortedness probe: 64 windows of 33 consecutive rows spread over the batch; a lane compares the cell of its row
// with the next row's.  *verdict = 1 when at least half of the sampled neighbours share a cell (cell-sorted or
// bunched batches: on average a warp item then needs <= 16 distinct blocks), else 0.
template <int D>
__global__ void __launch_bounds__(256) probe_kernel(const QueryParams p, int* verdict) {
    __shared__ int hits[8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int same = 0;
    for (int w = wid; w < 64; w += 8) {
        const int64_t base = (p.N - 33) * (int64_t)w / 63;
        const Located<D> a = locate<D>(p, base + lane), b = locate<D>(p, base + lane + 1);
        same += __popc(__ballot_sync(0xffffffffu, a.cell_global == b.cell_global && a.ok));
    }
    if (lane == 0) hits[wid] = same;
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
        for (int i = 0; i < 8; ++i) tot += hits[i];
        *verdict = (tot >= 64 * 32 / 2) ? 1 : 0;
    }
}

// ======================================================================================
// Variant 2: naive one-thread-per-query from global memory
// ======================================================================================
template <int D, int C, int MODE>
__global__ void __launch_bounds__(128) query_direct_kernel(const QueryParams p) {
    constexpr int NM = (D == 3) ? 64 : 256;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < p.N; n += stride) {
        const Located<D> L = locate<D>(p, n);
        if (p.out_cell) p.out_cell[n] = L.cell_global;
        if (L.masked) mask_row_in_place(p, n);
        eval_and_store<D, C, MODE, false>(p, n, L, p.table + L.cell_local * (int64_t)(C * NM));
    }
}


// ======================================================================================
// Table-free path: evaluate straight from the 4x4x4 grid neighbourhood
// ======================================================================================
// In 3-D the reference matrix is exactly A = M (x) M (x) M with M the 4x4 Catmull-Rom matrix
// (SURVEY fact 4), so value = sum_kji f[k][j][i] wz_k(w) wy_j(v) wx_i(u) with the Catmull-Rom
// weights w(t) = M^T [1,t,t^2,t^3] -- no coefficient table, 64x less memory, no build.  A lane owns
// (query, component) and asks the TMA unit for the neighbourhood as a tiled box of the grid tensor
// (cp.async.bulk.tensor, UTMALDG).  fp64 boxes must start on a 16-byte boundary (an odd x start traps
// with "illegal instruction", tools/micro/tma_dbg.cu), so the box is 6x4x4 starting at the even
// x <= ix and the x weights are placed at offset ix&1 inside a 6-vector whose other entries are 0.
// Slots are 768 B apart (TMA needs 128-byte alignment), which would put every lane on the same
// bank; rows are therefore visited in a per-lane rotated order (j by lane&3, k by (lane>>2)&1) with
// the y/z weight vectors pre-rotated to match, which makes the LDS.128 stream conflict-free.
struct GridQueryParams {
    QueryParams q;
};

using gridfree::catmull_rom;

template <int MODE, int THREADS, bool DEDUP>
__global__ void __launch_bounds__(THREADS) query_grid_kernel(const __grid_constant__ CUtensorMap tmap, const QueryParams p) {
    constexpr int D = 3;
    constexpr int C = MODE == 0 ? 3 : (MODE == 1 ? 1 : 4);
    constexpr uint32_t BYTES = 6 * 4 * 4 * 8;     // 768
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bars[THREADS / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint64_t* bar = &bars[wid];
    if (lane == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncwarp();
    const int64_t warp_global = ((int64_t)blockIdx.x * THREADS + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * THREADS) >> 5;
    const int64_t nbatch = (p.N + 31) / 32;
    const int64_t nitem = nbatch * C;
    const int rj = lane & 3, rk = (lane >> 2) & 1;     // row rotation of this lane
    unsigned char* slot = smem + (size_t)threadIdx.x * BYTES;
    uint32_t phase = 0;
    for (int64_t item = warp_global; item < nitem; item += nwarps) {
        const int64_t batch = item / C;
        const int comp = (int)(item - batch * C);
        const int64_t n = batch * 32 + lane;
        Located<D> L;
        L.ok = false; L.masked = false; L.cell_global = 0; L.cell_local = 0;
        L.idx[0] = L.idx[1] = L.idx[2] = 0;
        if (n < p.N) L = locate<D>(p, n);
        // DEDUP: lanes whose queries fall into the same cell share one box (the elected lane fetches it, the
        // others read its slot with their own row rotation -- still distinct banks or a broadcast)
        bool fetch = L.ok;
        const unsigned char* rd = slot;
        if (DEDUP) {
            const unsigned peers = __match_any_sync(0xffffffffu, L.ok ? L.cell_global : (int64_t)(-1 - lane));
            const int src_lane = __ffs(peers) - 1;
            fetch = L.ok && (src_lane == lane);
            rd = smem + (size_t)(threadIdx.x - lane + src_lane) * BYTES;
        }
        const unsigned fmask = __ballot_sync(0xffffffffu, fetch);
        if (lane == 0) mbar_expect_tx(bar, (uint32_t)__popc(fmask) * BYTES);
        __syncwarp();
        const int off = L.idx[0] & 1;
        if (fetch) tma_load_4d(slot, &tmap, bar, L.idx[0] - off, L.idx[1], L.idx[2], comp);
        if (comp == 0 && n < p.N) {
            if (p.out_cell) p.out_cell[n] = L.cell_global;
            if (L.masked) mask_row_in_place(p, n);
        }
        // weights while the box is in flight
        double wx[4], dwx[4], wy[4], dwy[4], wz[4], dwz[4];
        catmull_rom(L.frac[0], wx, dwx);
        catmull_rom(L.frac[1], wy, dwy);
        catmull_rom(L.frac[2], wz, dwz);
        double wx6[6], dwx6[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const int s = i - off;     // 0..3 inside the neighbourhood
            wx6[i] = (s == 0) ? wx[0] : (s == 1) ? wx[1] : (s == 2) ? wx[2] : (s == 3) ? wx[3] : 0.0;
            dwx6[i] = (s == 0) ? dwx[0] : (s == 1) ? dwx[1] : (s == 2) ? dwx[2] : (s == 3) ? dwx[3] : 0.0;
        }
        double wyr[4], dwyr[4], wzr[4], dwzr[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int jj = (j + rj) & 3, kk = (j + rk) & 3;
            wyr[j] = jj == 0 ? wy[0] : jj == 1 ? wy[1] : jj == 2 ? wy[2] : wy[3];
            dwyr[j] = jj == 0 ? dwy[0] : jj == 1 ? dwy[1] : jj == 2 ? dwy[2] : dwy[3];
            wzr[j] = kk == 0 ? wz[0] : kk == 1 ? wz[1] : kk == 2 ? wz[2] : wz[3];
            dwzr[j] = kk == 0 ? dwz[0] : kk == 1 ? dwz[1] : kk == 2 ? dwz[2] : dwz[3];
        }
        mbar_wait(bar, phase);
        phase ^= 1;
        const bool grad_comp = (MODE == 1) || (MODE == 2 && comp == 3);     // warp-uniform
        double val = 0.0, gx = 0.0, gy = 0.0, gz = 0.0;
        if (L.ok) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                double P = 0.0, Px = 0.0, Py = 0.0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int row = 4 * ((k + rk) & 3) + ((j + rj) & 3);
                    const double2* r = reinterpret_cast<const double2*>(rd + row * 48);
                    const double2 a = r[0], b = r[1], c = r[2];
                    double pp = a.x * wx6[0];
                    pp = fma(a.y, wx6[1], pp); pp = fma(b.x, wx6[2], pp); pp = fma(b.y, wx6[3], pp);
                    pp = fma(c.x, wx6[4], pp); pp = fma(c.y, wx6[5], pp);
                    P = fma(wyr[j], pp, P);
                    if (grad_comp) {
                        double dp = a.x * dwx6[0];
                        dp = fma(a.y, dwx6[1], dp); dp = fma(b.x, dwx6[2], dp); dp = fma(b.y, dwx6[3], dp);
                        dp = fma(c.x, dwx6[4], dp); dp = fma(c.y, dwx6[5], dp);
                        Px = fma(wyr[j], dp, Px);
                        Py = fma(dwyr[j], pp, Py);
                    }
                }
                val = fma(wzr[k], P, val);
                if (grad_comp) {
                    gx = fma(wzr[k], Px, gx);
                    gy = fma(wzr[k], Py, gy);
                    gz = fma(dwzr[k], P, gz);
                }
            }
        }
        if (n < p.N) {
            const double nan = qnan();
            if (!grad_comp) {
                p.out_comps[n * 3 + comp] = L.ok ? val : nan;
            } else {
                p.out_norm[n] = L.ok ? val : nan;
                p.out_grad[n * 3 + 0] = L.ok ? __ddiv_rn(gx, p.h[0]) : nan;
                p.out_grad[n * 3 + 1] = L.ok ? __ddiv_rn(gy, p.h[1]) : nan;
                p.out_grad[n * 3 + 2] = L.ok ? __ddiv_rn(gz, p.h[2]) : nan;
            }
        }
        __syncwarp();
    }
}

// 4-D table-free query (quadcubic(..., table=False)).  Four adjacent lanes own the four grid planes
// l = 0..3 of (query, component): each fetches its plane's 6x4x4 box through a 5-D tensor map, contracts it
// with the Catmull-Rom weights in (u, v, w) and scales by the t weight of its plane; with the reference quirk
// (A.py:860) the lanes also accumulate the plane's signed parity sums, exchange them over two shuffles to form
// fxyzt at the cell's 16 corners, and lanes 0/1 add the rank-16 term of their ct (arb_gridfree.cuh).  The four
// partial results are summed over two more shuffle steps.
template <int MODE, int THREADS, bool QUIRK, bool DEDUP>
__global__ void __launch_bounds__(THREADS) query_grid4_kernel(const __grid_constant__ CUtensorMap tmap, const QueryParams p) {
    constexpr int D = 4;
    constexpr int C = MODE == 0 ? 3 : (MODE == 1 ? 1 : 4);
    constexpr uint32_t BYTES = 6 * 4 * 4 * 8;     // 768
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bars[THREADS / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int qi = lane >> 2, l = lane & 3;
    uint64_t* bar = &bars[wid];
    if (lane == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncwarp();
    const int64_t warp_global = ((int64_t)blockIdx.x * THREADS + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * THREADS) >> 5;
    const int64_t nbatch = (p.N + 7) / 8;
    const int64_t nitem = nbatch * C;
    const int rj = lane & 3, rk = (lane >> 2) & 1;     // row rotation of this lane
    const unsigned char* slot = smem + (size_t)threadIdx.x * BYTES;
    uint32_t phase = 0;
    for (int64_t item = warp_global; item < nitem; item += nwarps) {
        const int64_t batch = item / C;
        const int comp = (int)(item - batch * C);
        const int64_t n = batch * 8 + qi;
        Located<D> L;
        L.ok = false; L.masked = false; L.cell_global = 0; L.cell_local = 0;
        L.idx[0] = L.idx[1] = L.idx[2] = L.idx[3] = 0;
        L.frac[0] = L.frac[1] = L.frac[2] = L.frac[3] = 0.0;
        if (n < p.N) L = locate<D>(p, n);
        bool fetch = L.ok;                                   // DEDUP: one box per distinct (cell, plane) in the warp
        const unsigned char* rd = slot;
        if (DEDUP) {
            const unsigned peers = __match_any_sync(0xffffffffu, L.ok ? L.cell_global * 4 + l : (int64_t)(-1 - lane));
            const int src_lane = __ffs(peers) - 1;
            fetch = L.ok && (src_lane == lane);
            rd = smem + (size_t)(threadIdx.x - lane + src_lane) * BYTES;
        }
        const unsigned fmask = __ballot_sync(0xffffffffu, fetch);
        if (lane == 0) mbar_expect_tx(bar, (uint32_t)__popc(fmask) * BYTES);
        __syncwarp();
        const int off = L.idx[0] & 1;
        if (fetch) tma_load_5d(const_cast<unsigned char*>(slot), &tmap, bar, L.idx[0] - off, L.idx[1], L.idx[2], L.idx[3] + l, comp);
        if (comp == 0 && l == 0 && n < p.N) {
            if (p.out_cell) p.out_cell[n] = L.cell_global;
            if (L.masked) mask_row_in_place(p, n);
        }
        // weights while the box is in flight
        double wx[4], dwx[4], wy[4], dwy[4], wz[4], dwz[4], wt[4], dwt[4];
        catmull_rom(L.frac[0], wx, dwx);
        catmull_rom(L.frac[1], wy, dwy);
        catmull_rom(L.frac[2], wz, dwz);
        catmull_rom(L.frac[3], wt, dwt);
        const double wt_l = gridfree::sel4(wt, l), dwt_l = gridfree::sel4(dwt, l);
        const bool grad_comp = (MODE == 1) || (MODE == 2 && comp == 3);     // warp-uniform
        mbar_wait(bar, phase);
        phase ^= 1;
        gridfree::PlanePartial pp;
        pp.val = pp.gx = pp.gy = pp.gz = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) pp.S[i] = 0.0;
        if (L.ok) {
            if (grad_comp) gridfree::plane_partial<true, QUIRK>(rd, off, rj, rk, wx, dwx, wy, dwy, wz, dwz, pp);
            else gridfree::plane_partial<false, QUIRK>(rd, off, rj, rk, wx, dwx, wy, dwy, wz, dwz, pp);
        }
        double m[5] = {pp.val * wt_l, pp.gx * wt_l, pp.gy * wt_l, pp.gz * wt_l, pp.val * dwt_l};
        if (QUIRK) {
            // fxyzt at the 8 corners of ct = l & 1: planes ct + 2 and ct are held by lanes l | 2 and l & 1
            double g[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const double other = __shfl_xor_sync(0xffffffffu, pp.S[i], 2);
                g[i] = 0.0625 * ((l >= 2) ? (pp.S[i] - other) : (other - pp.S[i]));
            }
            const double g7_other = __shfl_xor_sync(0xffffffffu, g[7], 1);
            const int ct = l & 1;
            double hx[2], dhx[2], hy[2], dhy[2], hz[2], dhz[2], ht[2], dht[2], c[4];
            gridfree::hermite_slope(L.frac[0], hx, dhx);
            gridfree::hermite_slope(L.frac[1], hy, dhy);
            gridfree::hermite_slope(L.frac[2], hz, dhz);
            gridfree::hermite_slope(L.frac[3], ht, dht);
            if (grad_comp) gridfree::corner_term<true>(g, ct ? g7_other : 0.0, hx, dhx, hy, dhy, hz, dhz, c);
            else gridfree::corner_term<false>(g, ct ? g7_other : 0.0, hx, dhx, hy, dhy, hz, dhz, c);
            if (l < 2) {          // lanes 2, 3 hold copies of the same corners
                const double h = ct ? ht[1] : ht[0], dh = ct ? dht[1] : dht[0];
                m[0] = fma(h, c[0], m[0]);
                m[1] = fma(h, c[1], m[1]);
                m[2] = fma(h, c[2], m[2]);
                m[3] = fma(h, c[3], m[3]);
                m[4] = fma(dh, c[0], m[4]);
            }
        }
        const int nred = grad_comp ? 5 : 1;
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            if (i < nred) {
                m[i] += __shfl_xor_sync(0xffffffffu, m[i], 1);
                m[i] += __shfl_xor_sync(0xffffffffu, m[i], 2);
            }
        }
        if (n < p.N && l == 0) {
            const double nan = qnan();
            if (!grad_comp) {
                p.out_comps[n * 3 + comp] = L.ok ? m[0] : nan;
            } else {
                p.out_norm[n] = L.ok ? m[0] : nan;
#pragma unroll
                for (int a = 0; a < D; ++a) p.out_grad[n * D + a] = L.ok ? __ddiv_rn(m[1 + a], p.h[a]) : nan;
            }
        }
        __syncwarp();
    }
}

// --------------------------------------------------------------------------------------
// launchers
// --------------------------------------------------------------------------------------
// Per-kernel, per-device launch set-up (dynamic shared-memory opt-in + occupancy) is done once and
// cached: it costs several microseconds per call, which matters for small batches.
struct LaunchCache { int occ[16]; };

template <typename K>
static int persistent_grid(K kernel, int threads, size_t smem, int64_t work_blocks, LaunchCache& cache) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    int& occ = cache.occ[dev & 15];
    if (occ == 0) {
        if (smem > 0) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int o = 1;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kernel, threads, smem) != cudaSuccess || o < 1) o = 1;
        occ = o;
    }
    int64_t g = (int64_t)num_sms() * occ;
    if (work_blocks < g) g = work_blocks;
    return (int)(g < 1 ? 1 : g);
}

template <int D, int C, int MODE, int U>
static int launch_coop(const QueryParams& p, cudaStream_t st) {
    constexpr int G = (D == 3) ? 8 : 32;
    constexpr int THREADS = 256;
    const int64_t per_block = (int64_t)(THREADS / G) * U;
    auto k = query_coop_kernel<D, C, MODE, U>;
    static LaunchCache cache = {};
    const int grid = persistent_grid(k, THREADS, 0, (p.N + per_block - 1) / per_block, cache);
    k<<<grid, THREADS, 0, st>>>(p);
    return check_cuda(cudaGetLastError(), "query_coop_kernel launch");
}

template <int D, int C, int MODE, int THREADS>
static int launch_bulk(const QueryParams& p, cudaStream_t st) {
    constexpr int NM = (D == 3) ? 64 : 256;
    const size_t smem = (size_t)THREADS * (C * NM * 8 + 16);
    auto k = query_bulk_kernel<D, C, MODE, THREADS>;
    static LaunchCache cache = {};
    const int grid = persistent_grid(k, THREADS, smem, (p.N + THREADS - 1) / THREADS, cache);
    k<<<grid, THREADS, smem, st>>>(p);
    return check_cuda(cudaGetLastError(), "query_bulk_kernel launch");
}

template <int D, int MODE, int THREADS, bool DEDUP, bool LOOPC = false, int SLOTS = 32, bool PREFETCH = false,
          bool FETCH_LDGSTS = false, int KIND = KIND_CELLS, bool QUIRK4 = true, bool ROUTED = false>
static int launch_block(const QueryParams& p, cudaStream_t st, const int* gate = nullptr, int gate_want = 0) {
    constexpr bool QUAD = (KIND == KIND_NODES_IL || KIND == KIND_GRID_IL);
    constexpr int C = QUAD ? 1 : (MODE == 0 ? 3 : (MODE == 1 ? 1 : 4));
    constexpr int QPW = (D == 4 || QUAD) ? 8 : 32;
    const size_t smem = (size_t)(THREADS / 32) * SLOTS * 528;
    auto k = query_block_kernel<D, MODE, THREADS, DEDUP, LOOPC, SLOTS, PREFETCH, FETCH_LDGSTS, KIND, QUIRK4, ROUTED>;
    static LaunchCache cache = {};
    const int64_t items = ((p.N + QPW - 1) / QPW) * (LOOPC ? 1 : C);
    const int grid = persistent_grid(k, THREADS, smem, (items + THREADS / 32 - 1) / (THREADS / 32), cache);
    k<<<grid, THREADS, smem, st>>>(p, gate, gate_want);
    return check_cuda(cudaGetLastError(), "query_block_kernel launch");
}

// One verdict word per launch, taken round-robin from a small per-device ring (launches of one stream are ordered;
// the host-buffer path keeps three streams busy, far fewer than the ring holds).
static int* probe_word() {
    static int* ring[16] = {};
    static std::atomic<unsigned> next{0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    int*& r = ring[dev & 15];
    if (!r) {
        static std::mutex m;
        std::lock_guard<std::mutex> lk(m);
        if (!r && cudaMalloc(&r, 256 * sizeof(int)) != cudaSuccess) { cudaGetLastError(); r = nullptr; return nullptr; }
    }
    return r + (next.fetch_add(1) & 255u);
}

// Default launch: large batches are probed for clustering on the device and the matching kernel runs (the other
// launch returns at once); no host synchronisation.
template <int D, int MODE, bool LOOPC, int SLOTS>
static int launch_auto(const QueryParams& p, cudaStream_t st) {
    int* word = (p.N >= (1 << 16)) ? probe_word() : nullptr;
    if (!word) return launch_block<D, MODE, 128, true, LOOPC>(p, st);
    probe_kernel<D><<<1, 256, 0, st>>>(p, word);
    int rc = check_cuda(cudaGetLastError(), "probe_kernel launch");
    if (rc) return rc;
    rc = launch_block<D, MODE, 128, true, LOOPC>(p, st, word, 0);
    if (rc) return rc;
    return launch_block<D, MODE, 128, true, LOOPC, SLOTS, true>(p, st, word, 1);
}

template <int D, int C, int MODE>
static int launch_direct(const QueryParams& p, cudaStream_t st) {
    auto k = query_direct_kernel<D, C, MODE>;
    static LaunchCache cache = {};
    const int grid = persistent_grid(k, 128, 0, (p.N + 127) / 128, cache);
    k<<<grid, 128, 0, st>>>(p);
    return check_cuda(cudaGetLastError(), "query_direct_kernel launch");
}

template <int D, int C, int MODE>
static int dispatch_variant(const QueryParams& p, cudaStream_t st, int variant) {
    constexpr int NM = (D == 3) ? 64 : 256;
    constexpr int BYTES = C * NM * 8;
    switch (variant) {
        case 1:   // per-thread slots holding all components of a query (first TMA design)
            if constexpr (BYTES <= 528) return launch_bulk<D, C, MODE, 128>(p, st);
            else if constexpr (BYTES <= 2048) return launch_bulk<D, C, MODE, 96>(p, st);
            else return launch_coop<D, C, MODE, 1>(p, st);
        case 2: return launch_direct<D, C, MODE>(p, st);
        case 10: return launch_coop<D, C, MODE, 1>(p, st);
        case 11: return launch_coop<D, C, MODE, (C == 1 ? 2 : 1)>(p, st);
        case 20: return launch_block<D, MODE, 128, false>(p, st);
        case 21: return launch_block<D, MODE, 64, true>(p, st);
        case 22: return launch_block<D, MODE, 192, true>(p, st);
        case 23: return launch_block<D, MODE, 384, true>(p, st);
        case 30: return launch_block<D, MODE, 128, true, true>(p, st);
        case 24: return launch_block<D, MODE, 128, true, D == 4, 32, true>(p, st);      // default + coordinate prefetch
        case 40: return launch_block<D, MODE, 128, true, D == 4, 16>(p, st);            // compact slot rings, forced
        case 41: return launch_block<D, MODE, 128, true, D == 4, 12>(p, st);
        case 42: return launch_block<D, MODE, 128, true, D == 4, 8>(p, st);
        case 43: return launch_block<D, MODE, 128, true, D == 4, 16, true>(p, st);
        case 44: return launch_block<D, MODE, 128, true, D == 4, 12, true>(p, st);
        case 45: return launch_block<D, MODE, 128, true, D == 4, 8, true>(p, st);
        case 60: return launch_block<D, MODE, 128, true, D == 4, 32, false, true>(p, st);  // warp-wide LDGSTS instead of bulk copies
        case 61: return launch_block<D, MODE, 128, true, D == 4, 32, true, true>(p, st);
        case 62: return launch_block<D, MODE, 128, true, D == 4, 16, true, true>(p, st);
        case 63: return launch_block<D, MODE, 128, true, D == 4, 12, true, true>(p, st);
        case 50: return launch_auto<D, MODE, D == 4, 16>(p, st);                        // probe picks 32 or 16 slots
        case 51: return launch_auto<D, MODE, D == 4, 12>(p, st);
        case 25:    // round-1 default: 32 slots per warp, no probe
            if constexpr (D == 4) return launch_block<D, MODE, 128, true, true>(p, st);
            else return launch_block<D, MODE, 128, true, false>(p, st);
        default:
            // batches of >= 2^16 rows are probed for clustering on the device: 32 slots per warp for scattered rows
            // (4-D: one locate per query and its components fetched in turn, +8 % in 'both'; 3-D: (batch, component)
            // items are independent, profiles/r01_variant_sweep.log), 16-slot rings + coordinate prefetch for
            // clustered ones (profiles/r02_cluster_bench.log)
            return launch_auto<D, MODE, D == 4, 16>(p, st);
    }
}

int fill_params(const char* who, const arb_geom* g, bool need_table, const double* table, int mode, double* q,
                int64_t N, int64_t ldq, double* out_comps, double* out_norm, double* out_grad, int64_t* out_cell,
                int64_t* masked_rows, unsigned long long* masked_count, QueryParams& p, bool need_outputs) {
    if (!g || (g->d != 3 && g->d != 4)) { set_error("%s: geometry missing or d not in {3,4}", who); return 1; }
    const int need_c = mode == ARB_MODE_VECTOR ? 3 : (mode == ARB_MODE_NORM ? 1 : (mode == ARB_MODE_BOTH ? 4 : -1));
    if (need_c < 0 || g->ncomp != need_c) {
        set_error("%s: mode %d needs %d components, geometry says %d", who, mode, need_c, g->ncomp);
        return 1;
    }
    if (N < 0 || ldq < g->d) { set_error("%s: need N >= 0 and ldq >= d (N=%lld ldq=%lld)", who, (long long)N, (long long)ldq); return 1; }
    if (N == 0) return -1;
    if ((need_table && !table) || !q) { set_error("%s: null table/grid or query pointer", who); return 1; }
    if (need_outputs && ((mode != ARB_MODE_NORM && !out_comps) || (mode != ARB_MODE_VECTOR && (!out_norm || !out_grad)))) {
        set_error("%s: output pointer missing for mode %d", who, mode);
        return 1;
    }
    if (masked_rows && !masked_count) { set_error("%s: masked_rows given without masked_count", who); return 1; }
    memset(&p, 0, sizeof(p));
    p.table = table; p.q = q; p.N = N; p.ldq = ldq;
    p.out_comps = out_comps; p.out_norm = out_norm; p.out_grad = out_grad; p.out_cell = out_cell;
    p.masked_rows = masked_rows; p.masked_count = masked_count;
    p.total_cells = 1; p.layer_cells = 1;
    for (int a = 0; a < g->d; ++a) {
        if (g->ncell[a] < 1 || !(g->h[a] > 0.0)) { set_error("%s: axis %d has ncell=%lld h=%g", who, a, (long long)g->ncell[a], g->h[a]); return 1; }
        p.mn[a] = g->int_min[a]; p.mx[a] = g->int_max[a]; p.h[a] = g->h[a]; p.nc[a] = g->ncell[a];
        p.total_cells *= g->ncell[a];
        if (a < g->d - 1) p.layer_cells *= g->ncell[a];
    }
    // node table (arb_nodes.cuh): 3-D [C][nz-2][ny-2][nx-3 x-pairs][2][8], 4-D [C][nt-2][nz-2][ny-2][nx-2][16]
    p.node_comp_stride = 16;
    for (int a = 0; a < 4; ++a) {
        p.nn[a] = (a < g->d) ? g->ncell[a] + 1 : 1;
        p.node_comp_stride *= (a == 0 && g->d == 3) ? g->ncell[0] : p.nn[a];
    }
    p.slab_lo = g->slab_lo; p.slab_hi = g->slab_hi;
    if (p.slab_lo < 0 || p.slab_hi > g->ncell[g->d - 1] || p.slab_lo >= p.slab_hi) {
        set_error("%s: bad slab [%lld,%lld)", who, (long long)p.slab_lo, (long long)p.slab_hi);
        return 1;
    }
    return 0;
}

int query_device(const arb_geom* g, const double* table, int mode, double* q, int64_t N, int64_t ldq,
                 double* out_comps, double* out_norm, double* out_grad, int64_t* out_cell, int64_t* masked_rows,
                 unsigned long long* masked_count, cudaStream_t st, int variant) {
    QueryParams p;
    const int rc = fill_params("arb_query", g, true, table, mode, q, N, ldq, out_comps, out_norm, out_grad, out_cell,
                               masked_rows, masked_count, p);
    if (rc) return rc < 0 ? 0 : rc;
    if (g->d == 3) {
        if (mode == ARB_MODE_VECTOR) return dispatch_variant<3, 3, 0>(p, st, variant);
        if (mode == ARB_MODE_NORM) return dispatch_variant<3, 1, 1>(p, st, variant);
        return dispatch_variant<3, 4, 2>(p, st, variant);
    }
    if (mode == ARB_MODE_VECTOR) return dispatch_variant<4, 3, 0>(p, st, variant);
    if (mode == ARB_MODE_NORM) return dispatch_variant<4, 1, 1>(p, st, variant);
    return dispatch_variant<4, 4, 2>(p, st, variant);
}

// Node (Hermite) table queries (arb_nodes.cuh).  variant: 0 = default, 1 = without coordinate prefetch,
// 2 = compact 16-slot rings.
template <int D, int MODE>
static int dispatch_nodes(const QueryParams& p, cudaStream_t st, int variant, bool quirk) {
    constexpr bool LC = (D == 4);
    if constexpr (D == 3 && MODE != 1) {       // 'vector' / 'both': components interleaved, four lanes per query
        switch (variant) {
            case 71: return launch_block<D, MODE, 128, true, true, 32, false, true, KIND_NODES_IL>(p, st);
            case 72: return launch_block<D, MODE, 128, true, true, 16, true, true, KIND_NODES_IL>(p, st);
            case 73: return launch_block<D, MODE, 128, false, true, 32, true, true, KIND_NODES_IL>(p, st);
            default: return launch_block<D, MODE, 128, true, true, 32, true, true, KIND_NODES_IL>(p, st);
        }
    } else {
        if (D == 4 && !quirk) return launch_block<D, MODE, 128, true, LC, 32, true, true, KIND_NODES, false>(p, st);
        switch (variant) {
            case 71: return launch_block<D, MODE, 128, true, LC, 32, false, true, KIND_NODES>(p, st);
            case 72: return launch_block<D, MODE, 128, true, LC, 16, true, true, KIND_NODES>(p, st);
            case 73: return launch_block<D, MODE, 128, false, LC, 32, true, true, KIND_NODES>(p, st);
            default: return launch_block<D, MODE, 128, true, LC, 32, true, true, KIND_NODES>(p, st);
        }
    }
}

// Table-free 'vector' / 'both' on the component-interleaved grid [nt][nz][ny][nx][4] (Bx, By, Bz, |B| or 0); the 4-D
// kernel lives in arb_gridil4.cu.
int query_gridil4_launch(const QueryParams& p, int mode, bool quirk, bool dedup, bool wide, cudaStream_t st);

int query_gridil_device(const arb_geom* g, const double* packed, int mode, double* q, int64_t N, int64_t ldq,
                        double* out_comps, double* out_norm, double* out_grad, int64_t* out_cell, int64_t* masked_rows,
                        unsigned long long* masked_count, cudaStream_t st) {
    QueryParams p;
    const int rc = fill_params("arb_query_gridil", g, true, packed, mode, q, N, ldq, out_comps, out_norm, out_grad,
                               out_cell, masked_rows, masked_count, p);
    if (rc) return rc < 0 ? 0 : rc;
    if (mode == ARB_MODE_NORM) { set_error("arb_query_gridil: 'vector' / 'both' only (one component has nothing to interleave)"); return 1; }
    if (g->slab_lo != 0 || g->slab_hi != g->ncell[g->d - 1]) { set_error("arb_query_gridil: slabs are not supported"); return 1; }
    if (reinterpret_cast<uintptr_t>(packed) & 31) { set_error("arb_query_gridil: grid must be 32-byte aligned"); return 1; }
    double bytes = 32.0;
    for (int a = 0; a < g->d; ++a) bytes *= (double)(g->ncell[a] + 3);
    if (bytes >= 137438953472.0) {
        set_error("arb_query_gridil: interleaved grids of 128 GB and more are not addressable by the gather");
        return 1;
    }
    const int v = current_query_variant();
    if (g->d == 4) {
        // 'both' with the A.py:860 term needs ~250 registers: two CTAs per SM without spills beat three with them (0.94
        // against 0.71 of the cell table); every other form fits 168 and runs three (variant 81 / 82 force either)
        const bool quirk = !(g->flags & ARB_GEOM_FIXED_D4);
        const bool wide = (v == 81) || (v != 82 && quirk && mode == ARB_MODE_BOTH);
        return query_gridil4_launch(p, mode, quirk, v != 73 && v != 20, wide, st);
    }
    if (mode == ARB_MODE_VECTOR) {
        if (v == 73) return launch_block<3, 0, 128, false, true, 32, true, true, KIND_GRID_IL>(p, st);
        return launch_block<3, 0, 128, true, true, 32, true, true, KIND_GRID_IL>(p, st);
    }
    if (v == 73) return launch_block<3, 2, 128, false, true, 32, true, true, KIND_GRID_IL>(p, st);
    return launch_block<3, 2, 128, true, true, 32, true, true, KIND_GRID_IL>(p, st);
}

int query_nodes_device(const arb_geom* g, const double* nodes, int mode, double* q, int64_t N, int64_t ldq,
                       double* out_comps, double* out_norm, double* out_grad, int64_t* out_cell, int64_t* masked_rows,
                       unsigned long long* masked_count, cudaStream_t st) {
    QueryParams p;
    const int rc = fill_params("arb_query_nodes", g, true, nodes, mode, q, N, ldq, out_comps, out_norm, out_grad,
                               out_cell, masked_rows, masked_count, p);
    if (rc) return rc < 0 ? 0 : rc;
    if (g->slab_lo != 0 || g->slab_hi != g->ncell[g->d - 1]) { set_error("arb_query_nodes: slabs are not supported (a node table is small enough to replicate)"); return 1; }
    if (reinterpret_cast<uintptr_t>(nodes) & 127) { set_error("arb_query_nodes: node table must be 128-byte aligned"); return 1; }
    const bool quirk = !(g->flags & ARB_GEOM_FIXED_D4);
    const int v = current_query_variant();
    if (g->d == 3) {
        if (mode == ARB_MODE_VECTOR) return dispatch_nodes<3, 0>(p, st, v, quirk);
        if (mode == ARB_MODE_NORM) return dispatch_nodes<3, 1>(p, st, v, quirk);
        return dispatch_nodes<3, 2>(p, st, v, quirk);
    }
    if (mode == ARB_MODE_VECTOR) return dispatch_nodes<4, 0>(p, st, v, quirk);
    if (mode == ARB_MODE_NORM) return dispatch_nodes<4, 1>(p, st, v, quirk);
    return dispatch_nodes<4, 2>(p, st, v, quirk);
}

// Slab-sharded tables: evaluate the rows this rank received and store every row's outputs into its home rank's result
// buffer (peer memory over NVLink) at its home row -- the return all-to-all and the re-ordering pass are gone.
int query_routed_device(const arb_geom* g, const double* table, int mode, double* q, int64_t N, int64_t ldq,
                        const int64_t* seg_start, const int64_t* home_row, double* const* peers, int npeers, int64_t ld,
                        cudaStream_t st) {
    QueryParams p;
    const int rc = fill_params("arb_query_routed", g, true, table, mode, q, N, ldq, nullptr, nullptr, nullptr, nullptr,
                               nullptr, nullptr, p, false);
    if (rc) return rc < 0 ? 0 : rc;
    const int d = g->d;
    const int ncomp_out = (mode == ARB_MODE_NORM) ? 0 : 3, ngrad = (mode == ARB_MODE_VECTOR) ? 0 : 1 + d;
    const int64_t want_ld = (ncomp_out + ngrad + 2) / 2 * 2;
    if (!seg_start || !home_row || !peers || npeers < 1 || npeers > ARB_MAX_PEERS || ld != want_ld) {
        set_error("arb_query_routed: need segments, 1..%d peers and ld == %lld (npeers=%d ld=%lld)", ARB_MAX_PEERS,
                  (long long)want_ld, npeers, (long long)ld);
        return 1;
    }
    p.npeers = npeers;
    for (int r = 0; r < npeers; ++r) {
        if (!peers[r] || (reinterpret_cast<uintptr_t>(peers[r]) & 15)) { set_error("arb_query_routed: peers[%d] is null or not 16-byte aligned", r); return 1; }
        if (seg_start[r] > seg_start[r + 1]) { set_error("arb_query_routed: bad segment %d", r); return 1; }
        p.peer[r] = peers[r];
        p.seg_start[r] = seg_start[r];
    }
    p.home_row = home_row;
    p.seg_start[npeers] = seg_start[npeers];
    if (seg_start[0] != 0 || seg_start[npeers] != N) { set_error("arb_query_routed: segments must cover the N rows"); return 1; }
    p.peer_ld = ld;
    if (d == 3) {
        if (mode == ARB_MODE_VECTOR) return launch_block<3, 0, 128, true, true, 32, false, false, KIND_CELLS, true, true>(p, st);
        if (mode == ARB_MODE_NORM) return launch_block<3, 1, 128, true, true, 32, false, false, KIND_CELLS, true, true>(p, st);
        return launch_block<3, 2, 128, true, true, 32, false, false, KIND_CELLS, true, true>(p, st);
    }
    if (mode == ARB_MODE_VECTOR) return launch_block<4, 0, 128, true, true, 32, false, false, KIND_CELLS, true, true>(p, st);
    if (mode == ARB_MODE_NORM) return launch_block<4, 1, 128, true, true, 32, false, false, KIND_CELLS, true, true>(p, st);
    return launch_block<4, 2, 128, true, true, 32, false, false, KIND_CELLS, true, true>(p, st);
}

// Inbox form of the routed query (both legs of a slab-sharded query fused into kernels): the rows were written into
// this rank's inbox by the senders' route kernels (arb_route_rows), their counts are in device memory.
int query_inbox_device(const arb_geom* g, const double* table, int mode, double* inbox, const int64_t* inbox_counts,
                       int64_t seg_cap, double* const* peers, int npeers, int64_t ld, cudaStream_t st) {
    if (!g || npeers < 1 || npeers > ARB_MAX_PEERS || seg_cap < 1 || !inbox || !inbox_counts || !peers) {
        set_error("arb_query_inbox: bad arguments (npeers=%d seg_cap=%lld)", npeers, (long long)seg_cap);
        return 1;
    }
    QueryParams p;
    const int d = g->d;
    const int64_t ld_in = (d + 2) / 2 * 2;
    const int rc = fill_params("arb_query_inbox", g, true, table, mode, inbox, (int64_t)npeers * seg_cap, ld_in, nullptr, nullptr,
                               nullptr, nullptr, nullptr, nullptr, p, false);
    if (rc) return rc < 0 ? 0 : rc;
    const int ncomp_out = (mode == ARB_MODE_NORM) ? 0 : 3, ngrad = (mode == ARB_MODE_VECTOR) ? 0 : 1 + d;
    const int64_t want_ld = (ncomp_out + ngrad + 2) / 2 * 2;
    if (ld != want_ld || (reinterpret_cast<uintptr_t>(inbox) & 15)) {
        set_error("arb_query_inbox: need ld == %lld and a 16-byte aligned inbox (ld=%lld)", (long long)want_ld, (long long)ld);
        return 1;
    }
    p.npeers = npeers;
    for (int r = 0; r < npeers; ++r) {
        if (!peers[r] || (reinterpret_cast<uintptr_t>(peers[r]) & 15)) { set_error("arb_query_inbox: peers[%d] is null or not 16-byte aligned", r); return 1; }
        p.peer[r] = peers[r];
    }
    p.peer_ld = ld;
    p.inbox_counts = inbox_counts;
    p.seg_cap = seg_cap;
    if (d == 3) {
        if (mode == ARB_MODE_VECTOR) return launch_block<3, 0, 128, true, true, 32, false, false, KIND_CELLS, true, true>(p, st);
        if (mode == ARB_MODE_NORM) return launch_block<3, 1, 128, true, true, 32, false, false, KIND_CELLS, true, true>(p, st);
        return launch_block<3, 2, 128, true, true, 32, false, false, KIND_CELLS, true, true>(p, st);
    }
    if (mode == ARB_MODE_VECTOR) return launch_block<4, 0, 128, true, true, 32, false, false, KIND_CELLS, true, true>(p, st);
    if (mode == ARB_MODE_NORM) return launch_block<4, 1, 128, true, true, 32, false, false, KIND_CELLS, true, true>(p, st);
    return launch_block<4, 2, 128, true, true, 32, false, false, KIND_CELLS, true, true>(p, st);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int MODE, bool DEDUP = true>
static int launch_grid(const CUtensorMap& tm, const QueryParams& p, cudaStream_t st) {
    constexpr int C = MODE == 0 ? 3 : (MODE == 1 ? 1 : 4);
    constexpr int THREADS = 128;
    const size_t smem = (size_t)THREADS * 768;
    auto k = query_grid_kernel<MODE, THREADS, DEDUP>;
    static LaunchCache cache = {};
    const int64_t items = ((p.N + 31) / 32) * C;
    const int grid = persistent_grid(k, THREADS, smem, (items + THREADS / 32 - 1) / (THREADS / 32), cache);
    k<<<grid, THREADS, smem, st>>>(tm, p);
    return check_cuda(cudaGetLastError(), "query_grid_kernel launch");
}

template <int MODE, bool QUIRK, bool DEDUP = true>
static int launch_grid4(const CUtensorMap& tm, const QueryParams& p, cudaStream_t st) {
    constexpr int C = MODE == 0 ? 3 : (MODE == 1 ? 1 : 4);
    constexpr int THREADS = 128;
    const size_t smem = (size_t)THREADS * 768;
    auto k = query_grid4_kernel<MODE, THREADS, QUIRK, DEDUP>;
    static LaunchCache cache = {};
    const int64_t items = ((p.N + 7) / 8) * C;
    const int grid = persistent_grid(k, THREADS, smem, (items + THREADS / 32 - 1) / (THREADS / 32), cache);
    k<<<grid, THREADS, smem, st>>>(tm, p);
    return check_cuda(cudaGetLastError(), "query_grid4_kernel launch");
}

int query_grid_device(const arb_geom* g, const double* grid, int64_t pitch_x, int mode, double* q, int64_t N,
                      int64_t ldq, double* out_comps, double* out_norm, double* out_grad, int64_t* out_cell,
                      int64_t* masked_rows, unsigned long long* masked_count, cudaStream_t st) {
    QueryParams p;
    const int rc = fill_params("arb_query_grid", g, true, grid, mode, q, N, ldq, out_comps, out_norm, out_grad,
                               out_cell, masked_rows, masked_count, p);
    if (rc) return rc < 0 ? 0 : rc;
    const int d = g->d;
    if (g->slab_lo != 0 || g->slab_hi != g->ncell[d - 1]) { set_error("arb_query_grid: slabs are not supported"); return 1; }
    int64_t npt[4] = {1, 1, 1, 1};
    for (int a = 0; a < d; ++a) npt[a] = g->ncell[a] + 3;
    if (pitch_x < npt[0] || (pitch_x & 1) || (reinterpret_cast<uintptr_t>(grid) & 15)) {
        set_error("arb_query_grid: row pitch must be even and >= nx, grid 16-byte aligned (pitch=%lld nx=%lld)",
                  (long long)pitch_x, (long long)npt[0]);
        return 1;
    }
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess) {
            set_error("arb_query_grid: cuTensorMapEncodeTiled not available from the driver");
            return 2;
        }
        encode = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    // grid [C][nt][nz][ny][pitch_x]: rank d + 1, box = 6 x 4 x 4 points of one plane of one component
    CUtensorMap tm;
    cuuint64_t gdim[5], gstr[4];
    cuuint32_t box[5] = {6, 4, 4, 1, 1}, estr[5] = {1, 1, 1, 1, 1};
    gdim[0] = (cuuint64_t)npt[0];
    uint64_t stride = (uint64_t)pitch_x * 8;
    for (int a = 1; a < d; ++a) { gdim[a] = (cuuint64_t)npt[a]; gstr[a - 1] = stride; stride *= (uint64_t)npt[a]; }
    gdim[d] = (cuuint64_t)g->ncomp; gstr[d - 1] = stride;
    const CUresult cr = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, d + 1, const_cast<double*>(grid), gdim, gstr, box,
                               estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                               CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { set_error("arb_query_grid: cuTensorMapEncodeTiled failed with CUresult %d", (int)cr); return 2; }
    const bool dedup = current_query_variant() != 20;      // variant 20: no warp de-duplication (here as in the table kernel)
    if (d == 3) {
        if (!dedup) {
            if (mode == ARB_MODE_VECTOR) return launch_grid<0, false>(tm, p, st);
            if (mode == ARB_MODE_NORM) return launch_grid<1, false>(tm, p, st);
            return launch_grid<2, false>(tm, p, st);
        }
        if (mode == ARB_MODE_VECTOR) return launch_grid<0>(tm, p, st);
        if (mode == ARB_MODE_NORM) return launch_grid<1>(tm, p, st);
        return launch_grid<2>(tm, p, st);
    }
    if (!dedup && !(g->flags & ARB_GEOM_FIXED_D4)) {
        if (mode == ARB_MODE_VECTOR) return launch_grid4<0, true, false>(tm, p, st);
        if (mode == ARB_MODE_NORM) return launch_grid4<1, true, false>(tm, p, st);
        return launch_grid4<2, true, false>(tm, p, st);
    }
    if (g->flags & ARB_GEOM_FIXED_D4) {       // corrected 4-D matrix: plain M^(x)4
        if (mode == ARB_MODE_VECTOR) return launch_grid4<0, false>(tm, p, st);
        if (mode == ARB_MODE_NORM) return launch_grid4<1, false>(tm, p, st);
        return launch_grid4<2, false>(tm, p, st);
    }
    if (mode == ARB_MODE_VECTOR) return launch_grid4<0, true>(tm, p, st);
    if (mode == ARB_MODE_NORM) return launch_grid4<1, true>(tm, p, st);
    return launch_grid4<2, true>(tm, p, st);
}

}  // namespace arb

extern "C" {

int arb_set_query_variant(int variant) {
    return arb::g_query_variant.exchange(variant);
}

int arb_query_grid(const arb_geom* g, const double* grid, int64_t pitch_x, int mode, double* q, int64_t N, int64_t ldq,
                   double* out_comps, double* out_norm, double* out_grad, int64_t* out_cell, int64_t* masked_rows,
                   unsigned long long* masked_count, void* stream) {
    return arb::query_grid_device(g, grid, pitch_x, mode, q, N, ldq, out_comps, out_norm, out_grad, out_cell,
                                  masked_rows, masked_count, (cudaStream_t)stream);
}

int arb_query_routed(const arb_geom* g, const double* table, int mode, double* q, int64_t N, int64_t ldq,
                     const int64_t* seg_start, const int64_t* home_row, double* const* peers, int npeers, int64_t ld,
                     void* stream) {
    return arb::query_routed_device(g, table, mode, q, N, ldq, seg_start, home_row, peers, npeers, ld,
                                    (cudaStream_t)stream);
}

int arb_query_inbox(const arb_geom* g, const double* table, int mode, double* inbox, const int64_t* inbox_counts,
                    int64_t seg_cap, double* const* peers, int npeers, int64_t ld, void* stream) {
    return arb::query_inbox_device(g, table, mode, inbox, inbox_counts, seg_cap, peers, npeers, ld, (cudaStream_t)stream);
}

int arb_query_gridil(const arb_geom* g, const double* packed, int mode, double* q, int64_t N, int64_t ldq,
                     double* out_comps, double* out_norm, double* out_grad, int64_t* out_cell, int64_t* masked_rows,
                     unsigned long long* masked_count, void* stream) {
    return arb::query_gridil_device(g, packed, mode, q, N, ldq, out_comps, out_norm, out_grad, out_cell, masked_rows,
                                    masked_count, (cudaStream_t)stream);
}

int arb_query_nodes(const arb_geom* g, const double* nodes, int mode, double* q, int64_t N, int64_t ldq,
                    double* out_comps, double* out_norm, double* out_grad, int64_t* out_cell, int64_t* masked_rows,
                    unsigned long long* masked_count, void* stream) {
    return arb::query_nodes_device(g, nodes, mode, q, N, ldq, out_comps, out_norm, out_grad, out_cell, masked_rows,
                                   masked_count, (cudaStream_t)stream);
}

int arb_query(const arb_geom* g, const double* table, int mode, double* q, int64_t N, int64_t ldq, double* out_comps,
              double* out_norm, double* out_grad, int64_t* out_cell, int64_t* masked_rows,
              unsigned long long* masked_count, void* stream) {
    return arb::query_device(g, table, mode, q, N, ldq, out_comps, out_norm, out_grad, out_cell, masked_rows,
                             masked_count, (cudaStream_t)stream, arb::current_query_variant());
}
}
