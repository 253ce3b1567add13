// Microbenchmark: how fast does one SM's TMA engine deliver boxes made of 64-byte vs 128-byte rows?
// (Decides the K-slab width of the 3xTF32 kernels.)  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tma_rate tma_rate.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../deeperwin_b200/csrc/tc_common.cuh"
using namespace dpe;

constexpr int STAGES = 4;
struct Args { int inner, rows, n_tiles_per_cta, dims; int n2, n3; long tiles_total; };

__global__ void __launch_bounds__(64, 1) k_tma(const __grid_constant__ CUtensorMap map, Args a, int mode) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t box_bytes = a.inner * 4 * a.rows;
    const uint32_t stage_bytes = (box_bytes + 1023) & ~1023u;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + STAGES * stage_bytes);
    uint64_t *full = bars, *empty = bars + STAGES;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int stage = 0; uint32_t phase = 0;
        for (long t = blockIdx.x; t < a.tiles_total; t += gridDim.x) {
            mbar_wait(&empty[stage], phase ^ 1);
            mbar_expect_tx(&full[stage], box_bytes);
            if (mode == 0) {           // 2-D: [rows_total][K]; tile = (kb, row block)
                const int n_kb = a.n2;
                const long rb = t / n_kb; const int kb = (int)(t - rb * n_kb);
                tma_load_2d(smem + stage * stage_bytes, &map, &full[stage], kb * a.inner, (int)(rb * a.rows));
            } else {                   // 4-D mo-like: (col, i, c, b); tile = (det group, channel tile, walker)
                const int n_g = a.n2, n_ct = a.n3;
                long r = t; const int g = (int)(r % n_g); r /= n_g; const int ct = (int)(r % n_ct); const long b = r / n_ct;
                tma_load_4d(smem + stage * stage_bytes, &map, &full[stage], g * a.inner, 0, ct * 15, (int)b);
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
    } else if (threadIdx.x == 32) {
        int stage = 0; uint32_t phase = 0;
        for (long t = blockIdx.x; t < a.tiles_total; t += gridDim.x) {
            mbar_wait(&full[stage], phase);
            mbar_arrive(&empty[stage]);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
    }
}

int main() {
    EncodeTiledFn enc = get_encode();
    if (!enc) { printf("no encode\n"); return 1; }
    const size_t bytes = 4ull << 30;
    float *buf; cudaMalloc(&buf, bytes); cudaMemset(buf, 0, bytes);
    cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    struct Case { const char *name; int mode, inner, rows; CUtensorMapSwizzle sw; };
    Case cases[] = {
        {"2D K=320 box 16x256 SW64 ", 0, 16, 256, CU_TENSOR_MAP_SWIZZLE_64B},
        {"2D K=320 box 32x128 SW128", 0, 32, 128, CU_TENSOR_MAP_SWIZZLE_128B},
        {"2D K=320 box 32x256 SW128", 0, 32, 256, CU_TENSOR_MAP_SWIZZLE_128B},
        {"2D K=320 box 16x256 none ", 0, 16, 256, CU_TENSOR_MAP_SWIZZLE_NONE},
        {"2D K=320 box  8x256 SW32 ", 0, 8, 256, CU_TENSOR_MAP_SWIZZLE_32B},
        {"4D mo    box 16x14x15 SW64", 1, 16, 210, CU_TENSOR_MAP_SWIZZLE_64B},
        {"4D mo    box 32x14x15 SW128", 1, 32, 210, CU_TENSOR_MAP_SWIZZLE_128B},
        {"4D mo    box 64x14x15 none", 1, 64, 210, CU_TENSOR_MAP_SWIZZLE_NONE},
    };
    for (auto &c : cases) {
        CUtensorMap map; Args a; a.inner = c.inner; a.rows = c.rows;
        CUresult r;
        if (c.mode == 0) {
            const int K = 320; const long M = (long)(bytes / 4 / K) / 256 * 256;
            cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)M}; cuuint64_t str[1] = {K * 4ull};
            cuuint32_t box[2] = {(cuuint32_t)c.inner, (cuuint32_t)c.rows}, es[2] = {1, 1};
            r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, buf, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, c.sw,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            a.n2 = K / c.inner; a.tiles_total = (M / c.rows) * a.n2; a.n3 = 0;
        } else {
            const int N = 14, C = 44, cols = 448; const long B = (long)(bytes / 4 / ((long)N * C * cols));
            cuuint64_t dims[4] = {(cuuint64_t)cols, (cuuint64_t)N, (cuuint64_t)C, (cuuint64_t)B};
            cuuint64_t str[3] = {(cuuint64_t)C * cols * 4, (cuuint64_t)cols * 4, (cuuint64_t)N * C * cols * 4};
            cuuint32_t box[4] = {(cuuint32_t)c.inner, 14, 15, 1}, es[4] = {1, 1, 1, 1};
            r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, buf, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, c.sw,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            a.n2 = cols / c.inner; a.n3 = 3; a.tiles_total = B * a.n2 * a.n3;
        }
        if (r != CUDA_SUCCESS) { printf("%s: encode failed %d\n", c.name, (int)r); continue; }
        const uint32_t stage_bytes = (c.inner * 4 * c.rows + 1023) & ~1023u;
        const size_t smem = STAGES * stage_bytes + 1024 + 256;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            k_tma<<<148, 64, smem>>>(map, a, c.mode);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        cudaError_t err = cudaGetLastError();
        const double gb = (double)a.tiles_total * c.inner * 4 * c.rows / 1e9;
        const double rows_per_sm = (double)a.tiles_total * c.rows / 148;
        printf("%s: %8.3f ms  %7.1f GB/s  %6.1f clk/row (1.9 GHz)  %s\n", c.name, ms, gb / (ms * 1e-3), ms * 1e-3 * 1.9e9 / rows_per_sm,
               err == cudaSuccess ? "" : cudaGetErrorString(err));
    }
    return 0;
}
