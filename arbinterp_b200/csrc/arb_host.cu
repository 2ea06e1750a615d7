// Host-buffer query path: the call a numpy user makes (tricubic.Query(points) with points in
// host memory).  Chunks the batch and overlaps H2D copy / query kernel / D2H copy on a ring of
// streams.  Pinned (page-locked) caller buffers are copied directly; pageable ones are staged
// through pinned ring buffers.  The in-place NaN masking of out-of-volume rows (A.py:350-355)
// is mirrored on the host from a compact list of masked row numbers, so q is never copied back.
#include <atomic>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>
#include <unistd.h>
#include <sched.h>
#include "arb_common.cuh"

namespace arb {

int query_device(const arb_geom* g, const double* table, int mode, double* q, int64_t N, int64_t ldq,
                 double* out_comps, double* out_norm, double* out_grad, int64_t* out_cell, int64_t* masked_rows,
                 unsigned long long* masked_count, cudaStream_t st, int variant);
int current_query_variant();
int query_grid_device(const arb_geom* g, const double* grid, int64_t pitch_x, int mode, double* q, int64_t N,
                      int64_t ldq, double* out_comps, double* out_norm, double* out_grad, int64_t* out_cell,
                      int64_t* masked_rows, unsigned long long* masked_count, cudaStream_t st);
int query_nodes_device(const arb_geom* g, const double* nodes, int mode, double* q, int64_t N, int64_t ldq,
                       double* out_comps, double* out_norm, double* out_grad, int64_t* out_cell, int64_t* masked_rows,
                       unsigned long long* masked_count, cudaStream_t st);

int query_gridil_device(const arb_geom* g, const double* packed, int mode, double* q, int64_t N, int64_t ldq,
                        double* out_comps, double* out_norm, double* out_grad, int64_t* out_cell, int64_t* masked_rows,
                        unsigned long long* masked_count, cudaStream_t st);

// one device-side query on whatever `table` is: grid_pitch > 0 = raw grid planes (table-free), -1 = node table,
// -2 = component-interleaved grid (table-free 'vector' / 'both'), 0 = cell coefficient table
static int query_any_device(const arb_geom* g, const double* table, int64_t grid_pitch, int mode, double* q, int64_t N,
                            int64_t ldq, double* out_comps, double* out_norm, double* out_grad, int64_t* out_cell,
                            int64_t* masked_rows, unsigned long long* masked_count, cudaStream_t st) {
    if (grid_pitch > 0)
        return query_grid_device(g, table, grid_pitch, mode, q, N, ldq, out_comps, out_norm, out_grad, out_cell,
                                 masked_rows, masked_count, st);
    if (grid_pitch == -2)
        return query_gridil_device(g, table, mode, q, N, ldq, out_comps, out_norm, out_grad, out_cell, masked_rows,
                                   masked_count, st);
    if (grid_pitch < 0)
        return query_nodes_device(g, table, mode, q, N, ldq, out_comps, out_norm, out_grad, out_cell, masked_rows,
                                  masked_count, st);
    return query_device(g, table, mode, q, N, ldq, out_comps, out_norm, out_grad, out_cell, masked_rows, masked_count,
                        st, current_query_variant());
}

namespace {

constexpr int NSLOT = 3;

// Staging copies between pageable caller memory and the pinned rings are plain memcpy; one thread moves
// ~10-25 GB/s, less than the PCIe link, so they run on a few helper threads.  Two forms: copy() splits one
// copy over the pool and waits (results leaving the ring); submit()/wait() queues a copy and returns, so the
// query rows of the chunks AHEAD of the one being issued are staged while the caller thread feeds the GPU
// (round 1 staged each chunk in line: the link idled during every memcpy and a pageable batch ran at 0.3-1.0e9
// q/s against 1.46e9 for a page-locked one).
class CopyPool {
  public:
    struct Ticket { int pending = 0; };          // guarded by the pool mutex

    // queue `bytes` in up to `max_parts` pieces (>= 1 MiB each) and return
    void submit(void* dst, const void* src, size_t bytes, Ticket* t, int max_parts) {
        if (nworkers_.load() == 0 || bytes == 0) { if (bytes) memcpy(dst, src, bytes); return; }
        int parts = (int)(bytes >> 20);
        if (parts > max_parts) parts = max_parts;
        if (parts < 1) parts = 1;
        const size_t step = (((bytes + parts - 1) / parts) + 4095) & ~(size_t)4095;
        {
            std::lock_guard<std::mutex> lk(m_);
            for (size_t off = 0; off < bytes; off += step) {
                tasks_.push_back({(char*)dst + off, (const char*)src + off, std::min(step, bytes - off), t});
                ++t->pending;
            }
        }
        cv_.notify_all();
    }
    void wait(Ticket* t) {
        std::unique_lock<std::mutex> lk(m_);
        done_cv_.wait(lk, [&] { return t->pending == 0; });
    }
    void copy(void* dst, const void* src, size_t bytes) {
        if (bytes < (4u << 20)) { memcpy(dst, src, bytes); return; }
        Ticket t;
        submit(dst, src, bytes, &t, nworkers_.load() > 0 ? nworkers_.load() : 1);
        wait(&t);
    }

    // Called once per host-buffer query: (re)size the pool to the cores the CALLING thread may run on -- a process
    // whose affinity was widened since the last call (or that now drives several GPUs) gets more helpers.
    void ensure_started() {
        if (pid_ != getpid()) {                    // forked child: the parent's threads do not exist here
            workers_.clear(); tasks_.clear();      // (already detached)
            nworkers_ = 0;
            pid_ = getpid();
        }
        grow();
    }

  private:
    struct Task { char* dst; const char* src; size_t n; Ticket* t; };
    void grow() {
        // the pool grows with the cores this process may run on (sched_getaffinity: a rank bound to its GPU's
        // cores gets its share, not the whole box), one core is left to the caller thread; ARB_COPY_THREADS overrides
        int allowed = 0;
        cpu_set_t set;
        if (sched_getaffinity(0, sizeof(set), &set) == 0) allowed = CPU_COUNT(&set);
        if (allowed <= 0) allowed = (int)std::thread::hardware_concurrency();
        // half of the allowed cores, at most 16: one GPU's staging needs ~35 GB/s (4 threads; a process that drives
        // several GPUs needs that per GPU), and a pool as large as the core count starves the caller thread that feeds
        // the GPU (16 cores: 8 helpers 1.18e9 q/s, 15 helpers 0.55e9)
        int n = allowed / 2;
        if (n > 16) n = 16;
        if (n < 1) n = 1;
        if (const char* e = getenv("ARB_COPY_THREADS")) { const int v = atoi(e); if (v >= 0 && v <= 64) n = v; }
        std::lock_guard<std::mutex> grow_lock(grow_m_);
        for (int i = (int)workers_.size(); i < n; ++i) {
            workers_.emplace_back([this] {
                for (;;) {
                    Task t;
                    {
                        std::unique_lock<std::mutex> lk(m_);
                        cv_.wait(lk, [&] { return !tasks_.empty(); });
                        t = tasks_.front();
                        tasks_.pop_front();
                    }
                    memcpy(t.dst, t.src, t.n);
                    {
                        std::lock_guard<std::mutex> lk(m_);
                        if (--t.t->pending == 0) done_cv_.notify_all();
                    }
                }
            });
            workers_.back().detach();
            nworkers_ = (int)workers_.size();
        }
    }
    std::mutex m_, grow_m_;
    std::condition_variable cv_, done_cv_;
    std::deque<Task> tasks_;
    std::vector<std::thread> workers_;
    std::atomic<int> nworkers_{0};
    pid_t pid_ = 0;
};
// Never destroyed: the workers block on the condition variable for the life of the process, and
// glibc's pthread_cond_destroy would wait for them forever in a static destructor at exit.
CopyPool& g_pool = *new CopyPool;

struct Slot {
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    // device
    double *d_q = nullptr, *d_comps = nullptr, *d_norm = nullptr, *d_grad = nullptr;
    int64_t *d_cell = nullptr, *d_rows = nullptr;
    unsigned long long* d_count = nullptr;
    // pinned host staging
    double *h_comps = nullptr, *h_norm = nullptr, *h_grad = nullptr;
    int64_t *h_cell = nullptr, *h_rows = nullptr;
    unsigned long long* h_count = nullptr;
    // in-flight chunk
    bool busy = false;
    int64_t off = 0, rows = 0;
};

// pinned staging ring for pageable query rows: deeper than the stream ring, so the helper threads run
// NSTAGE - 1 chunks ahead of the chunk being issued; a buffer is refilled once its own H2D copy has finished
constexpr int NSTAGE = 6;
struct StageBuf {
    double* h = nullptr;
    cudaEvent_t h2d_done = nullptr;
    bool used = false;
    CopyPool::Ticket ticket;
};

struct HostCtx {
    Slot slot[NSLOT];
    StageBuf stage[NSTAGE];
    int64_t stage_rows = 0, stage_ldq = 0;
    int64_t cap_rows = 0, cap_ldq = 0;
    int device = -1;
    std::mutex mutex;
};

HostCtx g_ctx[16];

void free_slot(Slot& s) {
    cudaFree(s.d_q); cudaFree(s.d_comps); cudaFree(s.d_norm); cudaFree(s.d_grad); cudaFree(s.d_cell);
    cudaFree(s.d_rows); cudaFree(s.d_count);
    cudaFreeHost(s.h_comps); cudaFreeHost(s.h_norm); cudaFreeHost(s.h_grad);
    cudaFreeHost(s.h_cell); cudaFreeHost(s.h_rows); cudaFreeHost(s.h_count);
    s.d_q = s.d_comps = s.d_norm = s.d_grad = nullptr;
    s.d_cell = s.d_rows = nullptr; s.d_count = nullptr;
    s.h_comps = s.h_norm = s.h_grad = nullptr;
    s.h_cell = s.h_rows = nullptr; s.h_count = nullptr;
}

int ensure_capacity(HostCtx& c, int64_t rows, int64_t ldq) {
    if (rows <= c.cap_rows && ldq <= c.cap_ldq) return 0;
    if (rows < c.cap_rows) rows = c.cap_rows;
    if (ldq < c.cap_ldq) ldq = c.cap_ldq;
    c.cap_rows = 0; c.cap_ldq = 0;            // a failed (re)allocation must not leave a stale capacity behind
    for (int i = 0; i < NSLOT; ++i) {
        Slot& s = c.slot[i];
        if (!s.stream) {
            ARB_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
            ARB_CUDA(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
        }
        free_slot(s);
        ARB_CUDA(cudaMalloc(&s.d_q, sizeof(double) * rows * ldq));
        ARB_CUDA(cudaMalloc(&s.d_comps, sizeof(double) * rows * 3));
        ARB_CUDA(cudaMalloc(&s.d_norm, sizeof(double) * rows));
        ARB_CUDA(cudaMalloc(&s.d_grad, sizeof(double) * rows * 4));
        ARB_CUDA(cudaMalloc(&s.d_cell, sizeof(int64_t) * rows));
        ARB_CUDA(cudaMalloc(&s.d_rows, sizeof(int64_t) * rows));
        ARB_CUDA(cudaMalloc(&s.d_count, sizeof(unsigned long long)));
        ARB_CUDA(cudaMallocHost(&s.h_comps, sizeof(double) * rows * 3));
        ARB_CUDA(cudaMallocHost(&s.h_norm, sizeof(double) * rows));
        ARB_CUDA(cudaMallocHost(&s.h_grad, sizeof(double) * rows * 4));
        ARB_CUDA(cudaMallocHost(&s.h_cell, sizeof(int64_t) * rows));
        ARB_CUDA(cudaMallocHost(&s.h_rows, sizeof(int64_t) * rows));
        ARB_CUDA(cudaMallocHost(&s.h_count, sizeof(unsigned long long)));
    }
    c.cap_rows = rows; c.cap_ldq = ldq;
    return 0;
}

int ensure_stage(HostCtx& c, int64_t rows, int64_t ldq) {
    if (rows <= c.stage_rows && ldq <= c.stage_ldq) return 0;
    if (rows < c.stage_rows) rows = c.stage_rows;
    if (ldq < c.stage_ldq) ldq = c.stage_ldq;
    c.stage_rows = 0; c.stage_ldq = 0;
    for (int i = 0; i < NSTAGE; ++i) {
        StageBuf& b = c.stage[i];
        if (!b.h2d_done) ARB_CUDA(cudaEventCreateWithFlags(&b.h2d_done, cudaEventDisableTiming));
        cudaFreeHost(b.h);
        b.h = nullptr; b.used = false;
        // (write-combined staging buffers were measured: no gain, 1.19-1.21e9 against 1.20-1.27e9 q/s)
        ARB_CUDA(cudaMallocHost(&b.h, sizeof(double) * rows * ldq));
    }
    c.stage_rows = rows; c.stage_ldq = ldq;
    return 0;
}

bool is_device(const void* p) {
    if (!p) return false;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeDevice;
}

bool is_pinned(const void* p) {
    if (!p) return true;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

struct Call {
    const arb_geom* g; int mode; int d;
    double* q; int64_t ldq;
    double *comps, *norm, *grad; int64_t* cell;
    bool q_pinned, comps_pinned, norm_pinned, grad_pinned, cell_pinned;
    bool cell_on_device;   // out_cell is device memory: the kernel writes it in place, no D2H
};

// wait for the slot's chunk and finish its host-side part
int retire(Slot& s, const Call& c) {
    if (!s.busy) return 0;
    ARB_CUDA(cudaEventSynchronize(s.done));
    const int64_t n = s.rows, off = s.off;
    if (c.comps && !c.comps_pinned) g_pool.copy(c.comps + off * 3, s.h_comps, sizeof(double) * n * 3);
    if (c.norm && !c.norm_pinned) g_pool.copy(c.norm + off, s.h_norm, sizeof(double) * n);
    if (c.grad && !c.grad_pinned) g_pool.copy(c.grad + off * c.d, s.h_grad, sizeof(double) * n * c.d);
    if (c.cell && !c.cell_on_device && !c.cell_pinned) g_pool.copy(c.cell + off, s.h_cell, sizeof(int64_t) * n);
    const unsigned long long cnt = *s.h_count;
    if (cnt) {
        ARB_CUDA(cudaMemcpyAsync(s.h_rows, s.d_rows, sizeof(int64_t) * cnt, cudaMemcpyDeviceToHost, s.stream));
        ARB_CUDA(cudaStreamSynchronize(s.stream));
        const double nan = __builtin_nan("");
        for (unsigned long long i = 0; i < cnt; ++i) {
            double* row = c.q + (off + s.h_rows[i]) * c.ldq;
            for (int64_t k = 0; k < c.ldq; ++k) row[k] = nan;
        }
    }
    s.busy = false;
    return 0;
}

}  // namespace
}  // namespace arb

// ---------------------------------------------------------------------------------------------------
// Small batches (the reference's single-point and short line queries, A.py:213-342): latency matters,
// not overlap.  One pinned + one device scratch block laid out [q | count | comps | norm | grad | cell |
// rows]: one H2D (q and the zeroed counter), the kernel, one D2H (counter and all outputs), one sync.
// Tiny batches (<= ARB_ZEROCOPY_ROWS rows, default 256) skip both copies: the pinned block is mapped into the
// device's address space, so the kernel reads the rows and writes the results straight over PCIe -- one launch and
// one stream synchronise per call.  The NaN-masked rows come back with the scratch copy of q itself.
// ---------------------------------------------------------------------------------------------------
namespace arb {
namespace {
constexpr int64_t SMALL_ROWS = 8192;

int64_t zero_copy_rows() {
    static int64_t v = -1;
    if (v < 0) {
        const char* e = getenv("ARB_ZEROCOPY_ROWS");
        v = e ? atoll(e) : 256;
        if (v < 0) v = 0;
    }
    return v;
}

struct SmallCtx {
    cudaStream_t stream = nullptr;
    unsigned char* d = nullptr;
    unsigned char* h = nullptr;
    size_t cap = 0;
    std::mutex mutex;
};
SmallCtx g_small[16];

int query_host_small(const arb_geom* g, const double* table, int64_t grid_pitch, int mode, double* q_host, int64_t N,
                     int64_t ldq, double* comps, double* norm, double* grad, int64_t* cell) {
    int dev = 0;
    ARB_CUDA(cudaGetDevice(&dev));
    SmallCtx& c = g_small[dev & 15];
    std::lock_guard<std::mutex> lock(c.mutex);
    const int d = g->d;
    const size_t o_q = 0, o_count = o_q + sizeof(double) * N * ldq, o_comps = o_count + 8,
                 o_norm = o_comps + sizeof(double) * N * 3, o_grad = o_norm + sizeof(double) * N,
                 o_cell = o_grad + sizeof(double) * N * d, o_rows = o_cell + 8 * N, total = o_rows + 8 * N;
    if (total > c.cap) {
        if (!c.stream) ARB_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
        cudaFree(c.d); cudaFreeHost(c.h);
        c.d = nullptr; c.h = nullptr; c.cap = 0;
        const size_t want = total < (1u << 20) ? (1u << 20) : total;
        ARB_CUDA(cudaMalloc(&c.d, want));
        ARB_CUDA(cudaMallocHost(&c.h, want));
        c.cap = want;
    }
    memcpy(c.h + o_q, q_host, sizeof(double) * N * ldq);
    const bool cell_dev = is_device(cell);
    if (N <= zero_copy_rows()) {
        // zero-copy: every pointer the kernel sees is the mapped pinned block (the cell indices too, unless the
        // caller keeps them on the device)
        int64_t* z_cell = cell ? (cell_dev ? cell : reinterpret_cast<int64_t*>(c.h + o_cell)) : nullptr;
        double* zq = reinterpret_cast<double*>(c.h + o_q);
        const int zrc = query_any_device(g, table, grid_pitch, mode, zq, N, ldq, reinterpret_cast<double*>(c.h + o_comps),
                                         reinterpret_cast<double*>(c.h + o_norm), reinterpret_cast<double*>(c.h + o_grad),
                                         z_cell, nullptr, nullptr, c.stream);
        if (zrc) return zrc;
        ARB_CUDA(cudaStreamSynchronize(c.stream));
        if (comps) memcpy(comps, c.h + o_comps, sizeof(double) * N * 3);
        if (norm) memcpy(norm, c.h + o_norm, sizeof(double) * N);
        if (grad) memcpy(grad, c.h + o_grad, sizeof(double) * N * d);
        if (cell && !cell_dev) memcpy(cell, c.h + o_cell, 8 * N);
        memcpy(q_host, c.h + o_q, sizeof(double) * N * ldq);      // rows the kernel NaN-masked in place (A.py:350-355)
        return 0;
    }
    *reinterpret_cast<unsigned long long*>(c.h + o_count) = 0ULL;
    ARB_CUDA(cudaMemcpyAsync(c.d, c.h, o_count + 8, cudaMemcpyHostToDevice, c.stream));
    int64_t* d_cell = cell ? (cell_dev ? cell : reinterpret_cast<int64_t*>(c.d + o_cell)) : nullptr;
    double* dq = reinterpret_cast<double*>(c.d + o_q);
    const int rc = query_any_device(g, table, grid_pitch, mode, dq, N, ldq, reinterpret_cast<double*>(c.d + o_comps),
                                    reinterpret_cast<double*>(c.d + o_norm), reinterpret_cast<double*>(c.d + o_grad), d_cell,
                                    reinterpret_cast<int64_t*>(c.d + o_rows),
                                    reinterpret_cast<unsigned long long*>(c.d + o_count), c.stream);
    if (rc) return rc;
    const size_t back_end = (cell && !cell_dev) ? o_rows : o_cell;
    ARB_CUDA(cudaMemcpyAsync(c.h + o_count, c.d + o_count, back_end - o_count, cudaMemcpyDeviceToHost, c.stream));
    ARB_CUDA(cudaStreamSynchronize(c.stream));
    if (comps) memcpy(comps, c.h + o_comps, sizeof(double) * N * 3);
    if (norm) memcpy(norm, c.h + o_norm, sizeof(double) * N);
    if (grad) memcpy(grad, c.h + o_grad, sizeof(double) * N * d);
    if (cell && !cell_dev) memcpy(cell, c.h + o_cell, 8 * N);
    const unsigned long long cnt = *reinterpret_cast<unsigned long long*>(c.h + o_count);
    if (cnt) {
        ARB_CUDA(cudaMemcpyAsync(c.h + o_rows, c.d + o_rows, 8 * cnt, cudaMemcpyDeviceToHost, c.stream));
        ARB_CUDA(cudaStreamSynchronize(c.stream));
        const int64_t* rows = reinterpret_cast<const int64_t*>(c.h + o_rows);
        const double nan = __builtin_nan("");
        for (unsigned long long i = 0; i < cnt; ++i)
            for (int64_t k = 0; k < ldq; ++k) q_host[rows[i] * ldq + k] = nan;
    }
    return 0;
}
}  // namespace
}  // namespace arb

namespace arb {
namespace {
// error exit of the pipelined path: nothing may stay in flight that still points at this call's buffers
int abandon(HostCtx& ctx, int rc) {
    for (int i = 0; i < NSTAGE; ++i) {             // helper threads still read the caller's rows
        g_pool.wait(&ctx.stage[i].ticket);
        ctx.stage[i].used = false;
    }
    for (int i = 0; i < NSLOT; ++i) {
        if (ctx.slot[i].stream) cudaStreamSynchronize(ctx.slot[i].stream);
        ctx.slot[i].busy = false;
    }
    cudaGetLastError();
    return rc;
}
}  // namespace
}  // namespace arb

static int query_host_impl(const arb_geom* g, const double* table, int64_t grid_pitch, int mode, double* q_host,
                           int64_t N, int64_t ldq, double* out_comps_host, double* out_norm_host,
                           double* out_grad_host, int64_t* out_cell_host, int64_t chunk_rows) {
    using namespace arb;
    if (!g || !q_host || N < 0) { set_error("arb_query_host: bad arguments"); return 1; }
    if (N == 0) return 0;
    if (N <= SMALL_ROWS && chunk_rows <= 0)
        return query_host_small(g, table, grid_pitch, mode, q_host, N, ldq,
                                (mode != ARB_MODE_NORM) ? out_comps_host : nullptr,
                                (mode != ARB_MODE_VECTOR) ? out_norm_host : nullptr,
                                (mode != ARB_MODE_VECTOR) ? out_grad_host : nullptr, out_cell_host);
    Call c;
    c.g = g; c.mode = mode; c.d = g->d; c.q = q_host; c.ldq = ldq;
    c.comps = (mode != ARB_MODE_NORM) ? out_comps_host : nullptr;
    c.norm = (mode != ARB_MODE_VECTOR) ? out_norm_host : nullptr;
    c.grad = (mode != ARB_MODE_VECTOR) ? out_grad_host : nullptr;
    c.cell = out_cell_host;
    c.q_pinned = is_pinned(q_host);
    c.comps_pinned = is_pinned(c.comps); c.norm_pinned = is_pinned(c.norm);
    c.grad_pinned = is_pinned(c.grad); c.cell_pinned = is_pinned(c.cell);
    c.cell_on_device = is_device(c.cell);
    if (chunk_rows <= 0) {
        // measured optima (profiles/r01_e2e_chunk_sweep.log, r01_midsize_chunks.log): 1 Mi rows when every
        // buffer is page-locked, 256 Ki when something is staged through the ring (the staging memcpy then
        // overlaps the copies), and never fewer than ~6 chunks so that mid-size batches pipeline at all
        const bool all_pinned = c.q_pinned && c.comps_pinned && c.norm_pinned && c.grad_pinned &&
                                (c.cell_pinned || c.cell_on_device);
        const int64_t base = all_pinned ? (1 << 20) : (1 << 18);
        int64_t sixth = (N + 5) / 6;
        if (sixth < 65536) sixth = 65536;
        chunk_rows = sixth < base ? sixth : base;
    }
    if (chunk_rows > N) chunk_rows = N;
    int dev = 0;
    ARB_CUDA(cudaGetDevice(&dev));
    HostCtx& ctx = g_ctx[dev & 15];
    std::lock_guard<std::mutex> lock(ctx.mutex);
    g_pool.ensure_started();
    int rc = ensure_capacity(ctx, chunk_rows, ldq);
    if (rc) return rc;
    if (!c.q_pinned) {
        rc = ensure_stage(ctx, chunk_rows, ldq);
        if (rc) return rc;
    }

#define ARB_CUDA_OR_ABANDON(expr)                                  \
    do {                                                            \
        int _rc = ::arb::check_cuda((expr), #expr);                 \
        if (_rc) return abandon(ctx, _rc);                          \
    } while (0)
    const int64_t nchunks = (N + chunk_rows - 1) / chunk_rows;
    // pageable rows: chunk j is staged into ring buffer j % NSTAGE by the helper threads, NSTAGE - 1 chunks ahead
    auto stage_chunk = [&](int64_t j) -> int {
        StageBuf& b = ctx.stage[j % NSTAGE];
        if (b.used) ARB_CUDA(cudaEventSynchronize(b.h2d_done));
        const int64_t o = j * chunk_rows, n = (N - o < chunk_rows) ? (N - o) : chunk_rows;
        g_pool.submit(b.h, q_host + o * ldq, sizeof(double) * n * ldq, &b.ticket, 2);
        b.used = true;
        return 0;
    };
    if (!c.q_pinned)
        for (int64_t j = 0; j < nchunks && j < NSTAGE - 1; ++j) {
            rc = stage_chunk(j);
            if (rc) return abandon(ctx, rc);
        }
    int64_t chunk = 0;
    for (int64_t off = 0; off < N; off += chunk_rows, ++chunk) {
        Slot& s = ctx.slot[chunk % NSLOT];
        rc = retire(s, c);
        if (rc) return abandon(ctx, rc);
        const int64_t n = (N - off < chunk_rows) ? (N - off) : chunk_rows;
        const double* src = q_host + off * ldq;
        StageBuf* sb = nullptr;
        if (!c.q_pinned) {
            if (chunk + NSTAGE - 1 < nchunks) {
                rc = stage_chunk(chunk + NSTAGE - 1);
                if (rc) return abandon(ctx, rc);
            }
            sb = &ctx.stage[chunk % NSTAGE];
            g_pool.wait(&sb->ticket);
            src = sb->h;
        }
        ARB_CUDA_OR_ABANDON(cudaMemcpyAsync(s.d_q, src, sizeof(double) * n * ldq, cudaMemcpyHostToDevice, s.stream));
        if (sb) ARB_CUDA_OR_ABANDON(cudaEventRecord(sb->h2d_done, s.stream));
        ARB_CUDA_OR_ABANDON(cudaMemsetAsync(s.d_count, 0, sizeof(unsigned long long), s.stream));
        int64_t* cell_dst = c.cell ? (c.cell_on_device ? c.cell + off : s.d_cell) : nullptr;
        rc = query_any_device(g, table, grid_pitch, mode, s.d_q, n, ldq, s.d_comps, s.d_norm, s.d_grad, cell_dst,
                              s.d_rows, s.d_count, s.stream);
        if (rc) return abandon(ctx, rc);
        if (c.comps)
            ARB_CUDA_OR_ABANDON(cudaMemcpyAsync(c.comps_pinned ? c.comps + off * 3 : s.h_comps, s.d_comps, sizeof(double) * n * 3,
                                     cudaMemcpyDeviceToHost, s.stream));
        if (c.norm)
            ARB_CUDA_OR_ABANDON(cudaMemcpyAsync(c.norm_pinned ? c.norm + off : s.h_norm, s.d_norm, sizeof(double) * n,
                                     cudaMemcpyDeviceToHost, s.stream));
        if (c.grad)
            ARB_CUDA_OR_ABANDON(cudaMemcpyAsync(c.grad_pinned ? c.grad + off * c.d : s.h_grad, s.d_grad, sizeof(double) * n * c.d,
                                     cudaMemcpyDeviceToHost, s.stream));
        if (c.cell && !c.cell_on_device)
            ARB_CUDA_OR_ABANDON(cudaMemcpyAsync(c.cell_pinned ? c.cell + off : s.h_cell, s.d_cell, sizeof(int64_t) * n,
                                     cudaMemcpyDeviceToHost, s.stream));
        ARB_CUDA_OR_ABANDON(cudaMemcpyAsync(s.h_count, s.d_count, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s.stream));
        ARB_CUDA_OR_ABANDON(cudaEventRecord(s.done, s.stream));
        s.busy = true; s.off = off; s.rows = n;
    }
    for (int i = 0; i < NSLOT; ++i) {
        rc = retire(ctx.slot[i], c);
        if (rc) return abandon(ctx, rc);
    }
    for (int i = 0; i < NSTAGE; ++i) ctx.stage[i].used = false;     // every H2D has completed (retire waited)
    return 0;
}

extern "C" int arb_query_host(const arb_geom* g, const double* table, int mode, double* q_host, int64_t N, int64_t ldq,
                              double* out_comps_host, double* out_norm_host, double* out_grad_host,
                              int64_t* out_cell_host, int64_t chunk_rows) {
    return query_host_impl(g, table, 0, mode, q_host, N, ldq, out_comps_host, out_norm_host, out_grad_host,
                           out_cell_host, chunk_rows);
}

extern "C" int arb_query_nodes_host(const arb_geom* g, const double* nodes, int mode, double* q_host, int64_t N,
                                    int64_t ldq, double* out_comps_host, double* out_norm_host, double* out_grad_host,
                                    int64_t* out_cell_host, int64_t chunk_rows) {
    return query_host_impl(g, nodes, -1, mode, q_host, N, ldq, out_comps_host, out_norm_host, out_grad_host,
                           out_cell_host, chunk_rows);
}

extern "C" int arb_query_gridil_host(const arb_geom* g, const double* packed, int mode, double* q_host, int64_t N,
                                     int64_t ldq, double* out_comps_host, double* out_norm_host, double* out_grad_host,
                                     int64_t* out_cell_host, int64_t chunk_rows) {
    return query_host_impl(g, packed, -2, mode, q_host, N, ldq, out_comps_host, out_norm_host, out_grad_host,
                           out_cell_host, chunk_rows);
}

extern "C" int arb_query_grid_host(const arb_geom* g, const double* grid, int64_t pitch_x, int mode, double* q_host,
                                   int64_t N, int64_t ldq, double* out_comps_host, double* out_norm_host,
                                   double* out_grad_host, int64_t* out_cell_host, int64_t chunk_rows) {
    if (pitch_x <= 0) { arb::set_error("arb_query_grid_host: pitch_x must be positive"); return 1; }
    return query_host_impl(g, grid, pitch_x, mode, q_host, N, ldq, out_comps_host, out_norm_host, out_grad_host,
                           out_cell_host, chunk_rows);
}

// Let kernels on the current device store into memory of `peer_device` (the routed form of the query kernel writes
// results into the home rank's buffer over NVLink).  Already-enabled is not an error.
extern "C" int arb_enable_peer_access(int peer_device) {
    int dev = 0;
    ARB_CUDA(cudaGetDevice(&dev));
    if (dev == peer_device) return 0;
    int can = 0;
    ARB_CUDA(cudaDeviceCanAccessPeer(&can, dev, peer_device));
    if (!can) { arb::set_error("arb_enable_peer_access: device %d cannot access device %d", dev, peer_device); return 2; }
    const cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return 0; }
    return arb::check_cuda(e, "cudaDeviceEnablePeerAccess");
}
