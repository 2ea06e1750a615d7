// Fused query + push: particles stay resident in HBM and are advanced by many
// velocity-Verlet steps of  dv/dt = kappa * grad|B|(x) + g  inside one kernel, the gradient being the
// quantity the reference's Query2/Query3 return (A.py:452, 519).  This is the application ARBInterp
// was written for (CHANGELOG.md:22 "simulating particle motion in magnetic fields"; the reference
// dropped its own trajectory code in 1.8, CHANGELOG.md:8) and what removes the per-step host round trip
// of a Python loop around Query.
//
// One lane per particle (four in 4-D).  Each step locates the particle's cell with the reference's arithmetic and
// evaluates the norm component's tricubic block exactly like query_block_kernel (same TMA bulk copy
// into the lane's shared-memory slot, same nested Horner).  The slot persists across steps: while a
// particle stays in its cell no memory traffic is issued at all, so small time steps run at FP64
// issue rate instead of HBM rate.  Particles that leave the interpolation volume are marked lost:
// position and velocity become NaN (the Query convention for out-of-volume points).
//
// Slab-sharded tables (arb_push_steps): every particle carries the index of the next step whose force has to
// be evaluated.  A particle that is inside the volume but outside this table's slab is parked -- position,
// velocity (with the half kick still pending) and step index are written back unchanged -- so that the rank
// owning the neighbouring slab resumes it with exactly the arithmetic of an unsharded run
// (sharding.SlabShardedInterp.push moves the rows).
#include "arb_device.cuh"
#include "arb_nodes.cuh"

namespace arb {

struct PushParams {
    QueryParams q;          // geometry + table (q.q etc. unused)
    double* pos;            // [N][D] (D = 4: x, y, z and the particle's own time)
    double* vel;            // [N][3]
    double dt, kappa, g[3];
    int64_t nsteps;
    unsigned long long* lost;
    int64_t* step_io;       // optional [N]: in = first step to evaluate (0 = fresh), out = step reached (nsteps + 1 = done)
    int ncomp;              // table components per cell; the norm block is the last one
};

// D = 3: one lane per particle.  D = 4 (time-dependent field): the norm component's 256 coefficients are
// four tricubic blocks alpha[.., l]; four adjacent lanes own (particle, l), each evaluates its block in
// (u, v, w) and the spatial gradient is the s^l-weighted sum over the four lanes (as in
// query_block_kernel); all four lanes carry the particle state redundantly and update it identically.
// pos is [N][D] (the 4th coordinate is the particle's own time and advances by dt per step), vel [N][3].
// NODES: `table` is a node (Hermite) table (arb_nodes.cuh; one component, or the |B| block of the interleaved 3-D
// layout): the lane's slot is gathered from its 4 (3-D) / 2 (4-D) x-pair segments by one warp-wide LDGSTS per slot and
// evaluated with the Hermite basis; in 4-D the four lanes own (cz, ct) and the A.py:860 term is added unless QUIRK4 is off.
template <int D, int THREADS, bool NODES = false, bool QUIRK4 = true>
__global__ void __launch_bounds__(THREADS) push_kernel(const PushParams P) {
    constexpr int SL = (D == 4) ? 4 : 1;
    constexpr int PPW = 32 / SL;                 // particles per warp pass
    constexpr uint32_t BYTES = 512, SLOT = 528;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bars[THREADS / 32];
    const QueryParams& p = P.q;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int pi = lane / SL, sl = lane % SL;
    uint64_t* bar = &bars[wid];
    if (lane == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncwarp();
    unsigned char* slot = smem + (size_t)threadIdx.x * SLOT;
    const int64_t warp_global = ((int64_t)blockIdx.x * THREADS + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * THREADS) >> 5;
    uint32_t phase = 0;
    const double hdt = 0.5 * P.dt;
    // node tables: bytes between a slot's first byte and this lane's 16 bytes of it in global memory, and the layout's
    // strides -- one component [..][nx-3 pairs][2][8] / [..][nx-2][16], or the 4th block of the interleaved [..][nx-2][4][8]
    const bool il = NODES && D == 3 && P.ncomp >= 3;
    int64_t lane_src = 0;
    if (NODES) {
        if (D == 4) lane_src = (lane >> 4) * p.nn[0] * 128 + (lane & 15) * 16;
        else if (il) lane_src = ((lane >> 4) * p.nn[1] + ((lane >> 3) & 1)) * p.nn[0] * 256 + ((lane >> 2) & 1) * 256 + 192 + (lane & 3) * 16;
        else lane_src = ((lane >> 4) * p.nn[1] + ((lane >> 3) & 1)) * p.nc[0] * 128 + (lane & 7) * 16;
    }
    for (int64_t base = warp_global * PPW; base < p.N; base += nwarps * PPW) {
        const int64_t n = base + pi;
        const bool active = n < p.N;
        double x[D], v[3] = {0, 0, 0};
#pragma unroll
        for (int a = 0; a < D; ++a) x[a] = active ? P.pos[n * D + a] : 0.0;
        if (active) {
#pragma unroll
            for (int a = 0; a < 3; ++a) v[a] = P.vel[n * 3 + a];
        }
        // `step` = index of the next force evaluation of this particle; lanes of a warp may be at different steps
        // (resumed particles), so the loop runs until no lane has work left.
        int64_t step = (active && P.step_io) ? P.step_io[n] : 0;
        bool alive = active && step >= 0 && step <= P.nsteps;
        int64_t cur_blk = -1;
        while (__any_sync(0xffffffffu, alive)) {
            Located<D> L;
            L.ok = false; L.cell_global = p.total_cells; L.cell_local = 0;
            if (alive) L = locate_coords<D>(p, x);
            if (alive && !L.ok) {
                alive = false;
                if (P.step_io && L.cell_global != p.total_cells) {
                    // inside the volume, outside this slab: parked as is, the owner of that slab resumes at `step`
                } else {                        // left the volume (or NaN): lost from here on
#pragma unroll
                    for (int a = 0; a < 3; ++a) x[a] = v[a] = qnan();
                    step = P.nsteps + 1;
                    if (P.lost && sl == 0) atomicAdd(P.lost, 1ULL);
                }
            }
            const int64_t blk = (L.cell_local * P.ncomp + (P.ncomp - 1)) * SL + sl;
            const bool fetch = alive && (blk != cur_blk);
            const unsigned fmask = __ballot_sync(0xffffffffu, fetch);
            if (fmask) {                         // warp-uniform
                if (NODES) {
                    const double* src;
                    if (D == 4)
                        src = p.table + (P.ncomp - 1) * p.node_comp_stride +
                              ((((L.idx[3] + (sl >> 1)) * p.nn[2] + L.idx[2] + (sl & 1)) * p.nn[1] + L.idx[1]) * p.nn[0] + L.idx[0]) * 16;
                    else if (il)
                        src = p.table + ((L.idx[2] * p.nn[1] + L.idx[1]) * p.nn[0] + L.idx[0]) * 32;
                    else
                        src = p.table + ((L.idx[2] * p.nn[1] + L.idx[1]) * p.nc[0] + L.idx[0]) * 16;
                    const uint32_t src16 = (uint32_t)((reinterpret_cast<const char*>(src) - reinterpret_cast<const char*>(p.table)) >> 7);   // 128-byte units: 512 GB
                    const char* const lane_base = reinterpret_cast<const char*>(p.table) + lane_src;
                    const uint32_t ring0 = smem_u32(smem + (size_t)(threadIdx.x - lane) * SLOT) + lane * 16;
#pragma unroll 8
                    for (int o = 0; o < 32; ++o) {
                        if ((fmask >> o) & 1u) {
                            const uint32_t s16 = __shfl_sync(0xffffffffu, src16, o);
                            cp_async_16(ring0 + o * SLOT, lane_base + ((size_t)s16 << 7));
                        }
                    }
                    if (fetch) cur_blk = blk;
                    cp_async_wait_all();
                    __syncwarp();
                } else {
                    if (lane == 0) mbar_expect_tx(bar, (uint32_t)__popc(fmask) * BYTES);
                    __syncwarp();
                    if (fetch) {
                        bulk_g2s(slot, p.table + blk * 64, BYTES, bar);
                        cur_blk = blk;
                    }
                    mbar_wait(bar, phase);
                    phase ^= 1;
                }
            }
            double g[5] = {0, 0, 0, 0, 0};
            if (NODES) {
                const double* cb = reinterpret_cast<const double*>(slot);
                if (D == 3) {
                    if (alive) nodes::eval3<true>(cb, L.frac, g);
                } else {
                    double f15[4] = {0.0, 0.0, 0.0, 0.0};
                    if (alive) {
                        nodes::eval4_lane<true, false>(cb, sl & 1, sl >> 1, L.frac, 0.0, g);
                        if (QUIRK4) { f15[0] = cb[15]; f15[1] = cb[31]; f15[2] = cb[47]; f15[3] = cb[63]; }
                    }
                    if (QUIRK4) {
                        const double up = __shfl_up_sync(0xffffffffu, f15[3], 1);
                        if (alive) nodes::quirk4_lane<true>(f15, sl ? up : 0.0, sl & 1, sl >> 1, L.frac, g);
                    }
#pragma unroll
                    for (int c = 1; c <= 3; ++c) {
                        g[c] += __shfl_xor_sync(0xffffffffu, g[c], 1);
                        g[c] += __shfl_xor_sync(0xffffffffu, g[c], 2);
                    }
                }
            } else if (alive) {
                eval_value_grad<3, true>(reinterpret_cast<const double*>(slot), L.frac, g);
            }
            if (D == 4 && !NODES) {
                const double w = alive ? pow_sel(L.frac[D - 1], sl) : 0.0;
#pragma unroll
                for (int c = 1; c <= 3; ++c) {
                    g[c] *= w;
                    g[c] += __shfl_xor_sync(0xffffffffu, g[c], 1);
                    g[c] += __shfl_xor_sync(0xffffffffu, g[c], 2);
                }
            }
            if (alive) {
                double a[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) a[c] = fma(P.kappa, __ddiv_rn(g[1 + c], p.h[c]), P.g[c]);
                if (step > 0) {                  // second half kick of the previous step
#pragma unroll
                    for (int c = 0; c < 3; ++c) v[c] = fma(hdt, a[c], v[c]);
                }
                if (step == P.nsteps) {
                    alive = false;
                } else {
#pragma unroll
                    for (int c = 0; c < 3; ++c) {    // first half kick + drift
                        v[c] = fma(hdt, a[c], v[c]);
                        x[c] = fma(P.dt, v[c], x[c]);
                    }
                    if (D == 4) x[3] += P.dt;
                }
                ++step;
            }
            __syncwarp();
        }
        if (active && sl == 0) {
#pragma unroll
            for (int a = 0; a < D; ++a) P.pos[n * D + a] = x[a];
#pragma unroll
            for (int a = 0; a < 3; ++a) P.vel[n * 3 + a] = v[a];
            if (P.step_io) P.step_io[n] = step;
        }
        __syncwarp();
    }
}

template <int D, bool NODES = false, bool QUIRK4 = true>
static int launch_push(const PushParams& P, int64_t N, cudaStream_t st) {
    constexpr int THREADS = 128;
    constexpr int PPW = (D == 4) ? 8 : 32;
    const size_t smem = (size_t)THREADS * 528;
    auto k = push_kernel<D, THREADS, NODES, QUIRK4>;
    ARB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, THREADS, smem) != cudaSuccess || occ < 1) occ = 1;
    int64_t grid = (int64_t)num_sms() * occ;
    const int64_t per_block = (int64_t)(THREADS / 32) * PPW;
    const int64_t need = (N + per_block - 1) / per_block;
    if (need < grid) grid = need;
    k<<<(unsigned)grid, THREADS, smem, st>>>(P);
    return check_cuda(cudaGetLastError(), "push_kernel launch");
}

}  // namespace arb

extern "C" int arb_push_steps(const arb_geom* g, const double* table, int mode, double* pos, double* vel,
                              int64_t* step_io, int64_t N, double dt, int64_t nsteps, double kappa,
                              const double* gravity, unsigned long long* lost_count, void* stream) {
    using namespace arb;
    if (mode != ARB_MODE_NORM && mode != ARB_MODE_BOTH) {
        set_error("arb_push: needs a table with a norm component (mode norm or both), got mode %d", mode);
        return 1;
    }
    if (!pos || !vel || nsteps < 0) { set_error("arb_push: null pos/vel or negative nsteps"); return 1; }
    PushParams P;
    memset(&P, 0, sizeof(P));
    const int rc = fill_params("arb_push", g, true, table, mode, pos, N, g ? g->d : 3, nullptr, nullptr, nullptr, nullptr,
                               nullptr, nullptr, P.q, false);
    if (rc) return rc < 0 ? 0 : rc;
    P.pos = pos; P.vel = vel; P.dt = dt; P.kappa = kappa; P.nsteps = nsteps; P.lost = lost_count; P.ncomp = g->ncomp;
    P.step_io = step_io;
    for (int a = 0; a < 3; ++a) P.g[a] = gravity ? gravity[a] : 0.0;
    if (g->d == 3) return launch_push<3>(P, N, (cudaStream_t)stream);
    return launch_push<4>(P, N, (cudaStream_t)stream);
}

// The same integrator on a node (Hermite) table (arb_build_nodes): 4-16x less memory, so a field whose cell table must be
// slab-sharded (config 5: 402 GB) is pushed on every GPU from its own 26 GB replica, with no particle migration.
extern "C" int arb_push_nodes(const arb_geom* g, const double* nodes, int mode, double* pos, double* vel, int64_t N,
                              double dt, int64_t nsteps, double kappa, const double* gravity,
                              unsigned long long* lost_count, void* stream) {
    using namespace arb;
    if (mode != ARB_MODE_NORM && mode != ARB_MODE_BOTH) {
        set_error("arb_push_nodes: needs a table with a norm component (mode norm or both), got mode %d", mode);
        return 1;
    }
    if (!pos || !vel || nsteps < 0) { set_error("arb_push_nodes: null pos/vel or negative nsteps"); return 1; }
    PushParams P;
    memset(&P, 0, sizeof(P));
    const int rc = fill_params("arb_push_nodes", g, true, nodes, mode, pos, N, g ? g->d : 3, nullptr, nullptr, nullptr,
                               nullptr, nullptr, nullptr, P.q, false);
    if (rc) return rc < 0 ? 0 : rc;
    if (g->slab_lo != 0 || g->slab_hi != g->ncell[g->d - 1]) { set_error("arb_push_nodes: slabs are not supported"); return 1; }
    if (reinterpret_cast<uintptr_t>(nodes) & 127) { set_error("arb_push_nodes: node table must be 128-byte aligned"); return 1; }
    P.pos = pos; P.vel = vel; P.dt = dt; P.kappa = kappa; P.nsteps = nsteps; P.lost = lost_count; P.ncomp = g->ncomp;
    P.step_io = nullptr;
    for (int a = 0; a < 3; ++a) P.g[a] = gravity ? gravity[a] : 0.0;
    if (g->d == 3) return launch_push<3, true>(P, N, (cudaStream_t)stream);
    if (g->flags & ARB_GEOM_FIXED_D4) return launch_push<4, true, false>(P, N, (cudaStream_t)stream);
    return launch_push<4, true, true>(P, N, (cudaStream_t)stream);
}

extern "C" int arb_push(const arb_geom* g, const double* table, int mode, double* pos, double* vel, int64_t N,
                        double dt, int64_t nsteps, double kappa, const double* gravity,
                        unsigned long long* lost_count, void* stream) {
    return arb_push_steps(g, table, mode, pos, vel, nullptr, N, dt, nsteps, kappa, gravity, lost_count, stream);
}

// ---------------------------------------------------------------------------------------------------
// Row permutation for the slab-sharded routing path (sharding.exchange_and_query): query rows are 24/32
// bytes and result rows 24..64 bytes; torch's row-gather kernel moves them at ~60 GB/s (one tiny block
// per row), which cost more than the NVLink exchange itself.  One thread per double keeps it at HBM speed.
// ---------------------------------------------------------------------------------------------------
namespace arb {
template <bool SCATTER>
__global__ void permute_rows_kernel(double* __restrict__ dst, const double* __restrict__ src,
                                    const int64_t* __restrict__ order, int64_t n, int width) {
    const int64_t total = n * width;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = e / width;
        const int col = (int)(e - row * width);
        const int64_t other = order[row];
        if (SCATTER) dst[other * width + col] = src[e];
        else dst[e] = src[other * width + col];
    }
}
}  // namespace arb

extern "C" int arb_permute_rows(double* dst, const double* src, const int64_t* order, int64_t n, int width,
                                int scatter, void* stream) {
    using namespace arb;
    if (n < 0 || width < 1 || (n > 0 && (!dst || !src || !order))) { set_error("arb_permute_rows: bad arguments"); return 1; }
    if (n == 0) return 0;
    const int64_t total = n * width;
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = (int64_t)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    if (scatter) permute_rows_kernel<true><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(dst, src, order, n, width);
    else permute_rows_kernel<false><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(dst, src, order, n, width);
    return check_cuda(cudaGetLastError(), "permute_rows_kernel launch");
}

// ---------------------------------------------------------------------------------------------------
// Owner keys for the slab-sharded routing path: for every row the rank whose slab holds its slowest-axis cell layer
// (the reference's cell location, floor((t - tIntMin) / ht), A.py:1081-1086 -- same IEEE subtract / divide / floor as
// the query kernel's locate) and whether the row lies outside the volume (A.py:1069-1076).  One kernel instead of the
// dozen elementwise torch launches it replaces (0.48 -> 0.03 ms for 4 M rows).
// ---------------------------------------------------------------------------------------------------
namespace arb {
struct OwnerParams {
    const double* q;
    int64_t n, ld;
    int d, nslab;
    double mn[4], mx[4], h_slow;
    int64_t hi[ARB_MAX_PEERS];
    int16_t* owner;
    unsigned char* outside;
};
__global__ void owner_keys_kernel(const OwnerParams p) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += (int64_t)gridDim.x * blockDim.x) {
        const double* row = p.q + i * p.ld;
        bool out = false;
        double slow = 0.0;
        for (int a = 0; a < p.d; ++a) {
            const double c = row[a];
            out |= (c < p.mn[a]) | (c > p.mx[a]);
            if (a == p.d - 1) slow = c;
        }
        const bool valid = (slow >= p.mn[p.d - 1]) & (slow <= p.mx[p.d - 1]);
        int owner = 0;
        if (valid) {
            const double fl = floor(__ddiv_rn(__dsub_rn(slow, p.mn[p.d - 1]), p.h_slow));
            int64_t layer = (int64_t)fl;
            const int64_t last = p.hi[p.nslab - 1] - 1;
            layer = layer < 0 ? 0 : (layer > last ? last : layer);
            while (owner < p.nslab - 1 && layer >= p.hi[owner]) ++owner;
        }
        p.owner[i] = (int16_t)owner;
        p.outside[i] = out ? 1 : 0;
    }
}
}  // namespace arb

extern "C" int arb_owner_keys(const arb_geom* g, const double* q, int64_t n, int64_t ldq, const int64_t* slab_hi,
                              int nslab, int16_t* owner, unsigned char* outside, void* stream) {
    using namespace arb;
    if (!g || (g->d != 3 && g->d != 4) || n < 0 || ldq < g->d || !slab_hi || nslab < 1 || nslab > ARB_MAX_PEERS) {
        set_error("arb_owner_keys: bad arguments");
        return 1;
    }
    if (n == 0) return 0;
    if (!q || !owner || !outside) { set_error("arb_owner_keys: null pointer"); return 1; }
    OwnerParams p;
    memset(&p, 0, sizeof(p));
    p.q = q; p.n = n; p.ld = ldq; p.d = g->d; p.nslab = nslab; p.owner = owner; p.outside = outside;
    for (int a = 0; a < g->d; ++a) { p.mn[a] = g->int_min[a]; p.mx[a] = g->int_max[a]; }
    p.h_slow = g->h[g->d - 1];
    for (int r = 0; r < nslab; ++r) p.hi[r] = slab_hi[r];
    int64_t blocks = (n + 255) / 256;
    const int64_t cap = (int64_t)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    owner_keys_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p);
    return check_cuda(cudaGetLastError(), "owner_keys_kernel launch");
}
