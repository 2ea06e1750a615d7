// Forward leg of slab-sharded queries fused into one kernel (round 2): every row of this rank's batch is written
// straight into the inbox of the rank that owns its slab -- peer memory over NVLink / NVSwitch -- together with the
// row number it has here, so that the owner's query kernel (arb_query_inbox) can store the results back at that row.
// Replaces, per step: the owner-key kernel, a 16-bit radix sort of the keys, the per-owner counts + their all-to-all
// (and the host synchronisation that reads them), the row-permutation kernel, and two NCCL all-to-alls (rows, home row
// numbers) -- by this kernel and one barrier.  The owning rank is the reference's cell location of the slowest axis,
// floor((t - tIntMin) / ht) (A.py:1081-1086), computed exactly like the query kernel's locate.
//
// Inbox of rank o: `nslab` segments of `seg_cap` rows, segment r filled by rank r only, so positions come from a
// counter that is local to the sender: no remote atomics.  A CTA takes tiles of 256 rows: lanes that go to the same
// owner are numbered with __match_any_sync / __popc, a shared-memory counter per owner numbers the warps' groups within
// the tile, one global atomicAdd per owner and tile reserves the tile's range in the segment, the tile is staged in
// shared memory ordered by owner, and every owner's run leaves as 16-byte pieces from consecutive threads.  The last CTA to finish
// publishes how many rows this rank put into every inbox (counts[o][my_rank] on rank o).
#include "arb_device.cuh"

namespace arb {

struct RouteParams {
    const double* q;
    int64_t n, ldq;
    int d, nslab, my_rank, ld_in;       // ld_in doubles per inbox row: [coords (d) | home row (int64 bits) | pad to even]
    double mn[4], mx[4], h_slow;
    int64_t hi[ARB_MAX_PEERS];
    double* inbox[ARB_MAX_PEERS];       // rank o's inbox, mapped here
    int64_t* counts[ARB_MAX_PEERS];     // rank o's per-sender row counts, mapped here
    int64_t seg_cap;
    unsigned long long* cursor;         // local [nslab], zero on entry: rows sent to every owner so far
    unsigned int* ticket;               // local, zero on entry: CTAs finished
    unsigned char* outside;             // local [n]: a coordinate outside the volume (A.py:1069-1076), optional
};

template <int D>
__global__ void __launch_bounds__(256) route_rows_kernel(const RouteParams p) {
    constexpr int LD = (D + 2) / 2 * 2;                   // 4 (d = 3) or 6 (d = 4) doubles per inbox row
    constexpr int CPR = LD / 2;                           // 16-byte pieces per row
    __shared__ unsigned int s_cnt[ARB_MAX_PEERS];
    __shared__ unsigned int s_off[ARB_MAX_PEERS + 1];     // where an owner's rows start in the staged tile
    __shared__ unsigned long long s_base[ARB_MAX_PEERS];  // ... and in this rank's segment of the owner's inbox
    __shared__ __align__(16) double s_rows[256 * LD];
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31;
    const int64_t ntile = (p.n + 255) / 256;
    for (int64_t tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
        if (threadIdx.x < ARB_MAX_PEERS) s_cnt[threadIdx.x] = 0;
        __syncthreads();
        const int64_t i = tile * 256 + threadIdx.x;
        const bool have = i < p.n;
        double c[D];
        int owner = -1 - lane;                            // rows beyond the batch match nobody
        if (have) {
            const double* row = p.q + i * p.ldq;
            bool out = false;
#pragma unroll
            for (int a = 0; a < D; ++a) {
                c[a] = row[a];
                out |= (c[a] < p.mn[a]) | (c[a] > p.mx[a]);
            }
            if (p.outside) p.outside[i] = out ? 1 : 0;
            const double slow = c[D - 1];
            owner = 0;                                    // no layer (outside the slow axis, NaN): rank 0 answers NaN
            if ((slow >= p.mn[D - 1]) & (slow <= p.mx[D - 1])) {
                int64_t layer = (int64_t)floor(__ddiv_rn(__dsub_rn(slow, p.mn[D - 1]), p.h_slow));
                const int64_t last = p.hi[p.nslab - 1] - 1;
                layer = layer < 0 ? 0 : (layer > last ? last : layer);
                while (owner < p.nslab - 1 && layer >= p.hi[owner]) ++owner;
            }
        }
        // number the rows of this tile per owner: lanes of a group, then the groups of the tile
        const unsigned peers = __match_any_sync(0xffffffffu, owner);
        const int in_group = __popc(peers & ((1u << lane) - 1u));
        const int leader = __ffs(peers) - 1;
        unsigned int group_base = 0;
        if (have && lane == leader) group_base = atomicAdd(&s_cnt[owner], (unsigned int)__popc(peers));
        group_base = __shfl_sync(0xffffffffu, group_base, leader);
        __syncthreads();
        if (threadIdx.x < p.nslab) {
            const unsigned int cnt = s_cnt[threadIdx.x];
            s_base[threadIdx.x] = cnt ? atomicAdd(&p.cursor[threadIdx.x], (unsigned long long)cnt) : 0ULL;
        }
        if (threadIdx.x == 0) {
            unsigned int run = 0;
            for (int o = 0; o < p.nslab; ++o) { s_off[o] = run; run += s_cnt[o]; }
            s_off[p.nslab] = run;
        }
        __syncthreads();
        // stage the tile ordered by owner, then store every owner's run as 16-byte pieces from consecutive threads:
        // small scattered remote stores are what NVLink is worst at (48-byte rows stored by their own threads: 0.68 ms
        // for 4 M rows at two ranks)
        if (have) {
            double* srow = s_rows + (size_t)(s_off[owner] + group_base + in_group) * LD;
#pragma unroll
            for (int a = 0; a < D; ++a) srow[a] = c[a];
            srow[D] = __longlong_as_double((long long)i);
            if (LD > D + 1) srow[D + 1] = 0.0;
        }
        __syncthreads();
        const int npiece = (int)s_off[p.nslab] * CPR;
        for (int k = threadIdx.x; k < npiece; k += 256) {
            const int r = k / CPR, part = k - r * CPR;
            int o = 0;
            for (int h = 1; h < p.nslab; ++h) o = (r >= (int)s_off[h]) ? h : o;
            double* dst = p.inbox[o] + ((int64_t)p.my_rank * p.seg_cap + (int64_t)s_base[o] + (r - (int)s_off[o])) * LD + part * 2;
            const double2 v = *reinterpret_cast<const double2*>(s_rows + (size_t)r * LD + part * 2);
            stg_stream_d2(dst, v.x, v.y);
        }
        __syncthreads();                                  // the shared arrays are reused by the next tile
    }
    // the last CTA publishes this rank's row counts in every owner's inbox header
    __threadfence_system();
    if (threadIdx.x == 0) s_last = (atomicAdd(p.ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (s_last && threadIdx.x < p.nslab) {
        const unsigned long long sent = atomicAdd(&p.cursor[threadIdx.x], 0ULL);
        p.counts[threadIdx.x][p.my_rank] = (int64_t)sent;
        __threadfence_system();
    }
}

}  // namespace arb

extern "C" int arb_route_rows(const arb_geom* g, const double* q, int64_t n, int64_t ldq, const int64_t* slab_hi, int nslab,
                              int my_rank, double* const* inboxes, int64_t* const* counts, int64_t seg_cap,
                              unsigned long long* cursor, unsigned int* ticket, unsigned char* outside, void* stream) {
    using namespace arb;
    if (!g || (g->d != 3 && g->d != 4) || n < 0 || ldq < g->d || !slab_hi || nslab < 1 || nslab > ARB_MAX_PEERS ||
        my_rank < 0 || my_rank >= nslab || !inboxes || !counts || !cursor || !ticket || n > seg_cap) {
        set_error("arb_route_rows: bad arguments (n=%lld seg_cap=%lld nslab=%d rank=%d)", (long long)n, (long long)seg_cap,
                  nslab, my_rank);
        return 1;
    }
    if (n > 0 && !q) { set_error("arb_route_rows: null rows"); return 1; }
    RouteParams p;
    memset(&p, 0, sizeof(p));
    p.q = q; p.n = n; p.ldq = ldq; p.d = g->d; p.nslab = nslab; p.my_rank = my_rank; p.ld_in = (g->d + 2) / 2 * 2;
    for (int a = 0; a < g->d; ++a) { p.mn[a] = g->int_min[a]; p.mx[a] = g->int_max[a]; }
    p.h_slow = g->h[g->d - 1];
    for (int r = 0; r < nslab; ++r) {
        if (!inboxes[r] || (reinterpret_cast<uintptr_t>(inboxes[r]) & 15) || !counts[r]) {
            set_error("arb_route_rows: inbox / counts of rank %d missing or not 16-byte aligned", r);
            return 1;
        }
        p.hi[r] = slab_hi[r]; p.inbox[r] = inboxes[r]; p.counts[r] = counts[r];
    }
    p.seg_cap = seg_cap; p.cursor = cursor; p.ticket = ticket; p.outside = outside;
    // the kernel always runs (n == 0 too): the counts must be published for every batch
    int64_t blocks = (n + 255) / 256;
    const int64_t cap = (int64_t)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    if (g->d == 3) route_rows_kernel<3><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p);
    else route_rows_kernel<4><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p);
    return check_cuda(cudaGetLastError(), "route_rows_kernel launch");
}
