// Microbenchmark: rate at which the TMA unit serves 4x4x4 fp64 boxes at arbitrary element offsets
// (cp.async.bulk.tensor.3d) versus 512-byte 1-D bulk copies -- decides whether a table-free query
// path (evaluate from the 64 grid values) can be TMA-fed.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>   // 0: tensor 4x4x4 box, 1: 1-D 512 B bulk
__global__ void __launch_bounds__(128) k(const __grid_constant__ CUtensorMap tm, const double* flat, const int4* coords,
                                         long n, int nx, int ny, double* out, int BOXB) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bars[4];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint64_t* bar = &bars[wid];
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    unsigned char* slot = smem + (size_t)threadIdx.x * BOXB;
    const long wg = ((long)blockIdx.x * 128 + threadIdx.x) >> 5, nw = ((long)gridDim.x * 128) >> 5;
    uint32_t phase = 0;
    double acc = 0;
    for (long base = wg * 32; base < n; base += nw * 32) {
        const long i = base + lane;
        const bool ok = i < n;
        int4 c = ok ? coords[i] : make_int4(0, 0, 0, 0);
        unsigned m = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"((MODE == 2 ? (int)(m & 1) : __popc(m)) * (MODE == 1 ? 512 : BOXB)) : "memory");
        __syncwarp();
        if (ok) {
            if (MODE == 0 || MODE == 2) {
            } else
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(s32(slot)), "l"(flat + ((long)c.x + (long)nx * (c.y + (long)ny * c.z)) / 64 * 64), "r"(512), "r"(s32(bar)) : "memory");
        }
        if (MODE == 0) {
            // manual waterfall: one elected lane issues the tensor copy of every active lane in turn
            for (unsigned rem = m; rem; rem &= rem - 1) {
                const int src = __ffs(rem) - 1;
                const int cx = __shfl_sync(0xffffffffu, c.x, src), cy = __shfl_sync(0xffffffffu, c.y, src), cz = __shfl_sync(0xffffffffu, c.z, src);
                const uint32_t dst = __shfl_sync(0xffffffffu, s32(slot), src);
                if (lane == 0)
                    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                                 ::"r"(dst), "l"(&tm), "r"(s32(bar)), "r"(cx), "r"(cy), "r"(cz), "r"(0) : "memory");
            }
        }
        if (MODE == 2 && lane == 0 && ok)
            asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                         ::"r"(s32(slot)), "l"(&tm), "r"(s32(bar)), "r"(c.x), "r"(c.y), "r"(c.z), "r"(0) : "memory");
        asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n"
                     ::"r"(s32(bar)), "r"(phase) : "memory");
        phase ^= 1;
        const double* d = (const double*)slot;
        acc += d[(lane * 2) & 63] + d[(lane * 2 + 33) & 63];
        __syncwarp();
    }
    if (acc == 12345.678) out[0] = acc;
}

int main(int argc, char** argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 256;
    const int bx = argc > 2 ? atoi(argv[2]) : 4, by = argc > 3 ? atoi(argv[3]) : 4, bz = argc > 4 ? atoi(argv[4]) : 4, prom = argc > 5 ? atoi(argv[5]) : 0;
    const int boxb = bx * by * bz * 8;
    const long nq = 1L << 24;
    double* grid; CK(cudaMalloc(&grid, sizeof(double) * n * n * n));
    CK(cudaMemset(grid, 0, sizeof(double) * n * n * n));
    std::vector<int4> h(nq);
    srand(1);
    for (long i = 0; i < nq; ++i) h[i] = make_int4((rand() % (n - 16)) & ~1, rand() % (n - 16), rand() % (n - 16), 0);
    int4* dc; CK(cudaMalloc(&dc, sizeof(int4) * nq)); CK(cudaMemcpy(dc, h.data(), sizeof(int4) * nq, cudaMemcpyHostToDevice));
    double* out; CK(cudaMalloc(&out, 8));
    void* fp = nullptr; cudaDriverEntryPointQueryResult qr;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qr));
    auto enc = (CUresult(*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                            const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill))fp;
    CUtensorMap tm;
    cuuint64_t gd[4] = {(cuuint64_t)n, (cuuint64_t)n, (cuuint64_t)n, 1}, gs[3] = {(cuuint64_t)n * 8, (cuuint64_t)n * n * 8, (cuuint64_t)n * n * n * 8};
    cuuint32_t box[4] = {(cuuint32_t)bx, (cuuint32_t)by, (cuuint32_t)bz, 1}, es[4] = {1, 1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, grid, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)prom, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    const size_t smem = 128 * (size_t)(boxb > 512 ? boxb : 512);
    CK(cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    if (smem > 200000) { printf("box too big\n"); return 1; }
    CK(cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int mode = 2; mode >= 0; --mode) {
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * 3, 128, smem>>>(tm, grid, dc, nq, n, n, out, boxb > 512 ? boxb : 512);
            else if (mode == 2) k<2><<<148 * 3, 128, smem>>>(tm, grid, dc, nq, n, n, out, boxb > 512 ? boxb : 512);
            else k<1><<<148 * 3, 128, smem>>>(tm, grid, dc, nq, n, n, out, boxb > 512 ? boxb : 512);
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("[tma_box] box %dx%dx%d prom %d grid %d^3 (%.0f MB) mode=%s rep%d: %.3f ms -> %.3e boxes/s\n", bx, by, bz, prom, n, n * (double)n * n * 8 / 1e6,
                   mode == 0 ? "tensor4x4x4" : (mode == 2 ? "tensor4x4x4-lane0only(1/32 boxes)" : "bulk512B"), rep, ms, (mode == 2 ? nq / 32 : nq) / (ms * 1e-3));
        }
    }
    return 0;
}
