// minimal TMA tensor-load probe: variants of who issues / which barrier / which slot
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("  -> CUDA error: %s\n", cudaGetErrorString(e)); return 1; } } while (0)
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// V: 0 = thread 0 only, CTA barrier; 1 = lane 0 of each warp, per-warp barrier, own slot; 2 = like 1 but all warps use bars[0] slot 0 (only warp 0 issues)
__global__ void __launch_bounds__(128) probe_g(const CUtensorMap* tmg, double* out, int boxb, int cx) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bars[4];
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bars[0])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bars[0])), "r"(boxb) : "memory");
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                     ::"r"(s32(smem)), "l"(tmg), "r"(s32(&bars[0])), "r"(cx), "r"(5), "r"(7), "r"(0) : "memory");
    }
    __syncthreads();
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(s32(&bars[0])), "r"(0) : "memory");
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = ((double*)smem)[0] + ((double*)smem)[boxb / 8 - 1];
}

template <int V>
__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap tm, double* out, int boxb) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bars[4];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (V == 0) {
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bars[0])));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bars[0])), "r"(boxb) : "memory");
            asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                         ::"r"(s32(smem)), "l"(&tm), "r"(s32(&bars[0])), "r"(3), "r"(5), "r"(7), "r"(0) : "memory");
        }
        __syncthreads();
        asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(s32(&bars[0])), "r"(0) : "memory");
        if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = ((double*)smem)[0] + ((double*)smem)[boxb / 8 - 1];
    } else {
        uint64_t* bar = &bars[wid];
        unsigned char* slot = smem + (size_t)wid * boxb;
        if (lane == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        if (lane == 0 && (V == 1 || wid == 0)) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(boxb) : "memory");
            asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                         ::"r"(s32(slot)), "l"(&tm), "r"(s32(bar)), "r"(3 + wid), "r"(5), "r"(7), "r"(0) : "memory");
        }
        __syncwarp();
        if (V == 1 || wid == 0)
            asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(s32(bar)), "r"(0) : "memory");
        if (lane == 0 && blockIdx.x == 0 && (V == 1 || wid == 0)) out[wid] = ((double*)slot)[0] + ((double*)slot)[boxb / 8 - 1];
    }
}

int main() {
    const int n = 64;
    double* grid; CK(cudaMalloc(&grid, sizeof(double) * n * n * n));
    double* h = (double*)malloc(sizeof(double) * n * n * n);
    for (int i = 0; i < n * n * n; ++i) h[i] = i;
    CK(cudaMemcpy(grid, h, sizeof(double) * n * n * n, cudaMemcpyHostToDevice));
    double* out; CK(cudaMalloc(&out, 64)); 
    void* fp = nullptr; cudaDriverEntryPointQueryResult qr;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qr));
    auto enc = (CUresult(*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                            const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill))fp;
    int boxes[3][3] = {{12, 7, 7}, {4, 4, 4}, {8, 4, 4}};
    for (auto& b : boxes) {
        CUtensorMap tm;
        cuuint64_t gd[4] = {(cuuint64_t)n, (cuuint64_t)n, (cuuint64_t)n, 1}, gs[3] = {(cuuint64_t)n * 8, (cuuint64_t)n * n * 8, (cuuint64_t)n * n * n * 8};
        cuuint32_t box[4] = {(cuuint32_t)b[0], (cuuint32_t)b[1], (cuuint32_t)b[2], 1}, es[4] = {1, 1, 1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, grid, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        const int boxb = b[0] * b[1] * b[2] * 8;
        const int slotb = (boxb + 127) / 128 * 128;
        printf("box %dx%dx%d (%d B, slot %d) encode=%d\n", b[0], b[1], b[2], boxb, slotb, (int)r);
        {
            unsigned long long* w = (unsigned long long*)&tm;
            printf("  desc words:");
            for (int i = 0; i < 16; ++i) printf(" %016llx", w[i]);
            printf("\n");
        }
        {
            CUtensorMap* dtm; CK(cudaMalloc(&dtm, 256)); CK(cudaMemcpy(dtm, &tm, sizeof(tm), cudaMemcpyHostToDevice));
            CK(cudaMemset(out, 0, 32));
            for (int cx : {0, 8, 4, 2, 6, 1, 3}) {
            probe_g<<<4, 128, 4 * slotb>>>(dtm, out, boxb, cx);
            cudaError_t e = cudaDeviceSynchronize();
            double ho = 0; cudaMemcpy(&ho, out, 8, cudaMemcpyDeviceToHost);
            printf("  global-descriptor variant cx=%d: %s out=%.0f\n", cx, cudaGetErrorString(e), ho);
            if (e != cudaSuccess) return 1;
            }
            return 0;
            unsigned long long* w = (unsigned long long*)&tm;
            printf("  desc words: %016llx %016llx %016llx %016llx %016llx %016llx\n", w[0], w[1], w[2], w[3], w[4], w[5]);
        }
        for (int v = 0; v < 3; ++v) {
            double ho[4] = {0, 0, 0, 0};
            CK(cudaMemset(out, 0, 32));
            if (v == 0) probe<0><<<4, 128, 4 * slotb>>>(tm, out, boxb);
            if (v == 1) probe<1><<<4, 128, 4 * slotb>>>(tm, out, boxb);
            if (v == 2) probe<2><<<4, 128, 4 * slotb>>>(tm, out, boxb);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("  variant %d: %s\n", v, cudaGetErrorString(e)); return 1; }
            cudaMemcpy(ho, out, 32, cudaMemcpyDeviceToHost);
            printf("  variant %d ok: out = %.0f %.0f %.0f %.0f (expect first+last of box at (3,5,7): %.0f)\n", v, ho[0], ho[1], ho[2], ho[3],
                   (double)(3 + n * (5 + n * 7)) + (double)((3 + b[0] - 1) + n * ((5 + b[1] - 1) + n * (7 + b[2] - 1))));
        }
    }
    return 0;
}
