// Microbenchmark: tcgen05.mma issue/execution rate on one SM as a function of N (M = 128, K = 16, bf16, cta_group::1),
// for SS (A, B from shared memory, K-major), TS (A from TMEM) and MN-major operands. One CTA, one converged warp issues
// `iters` back-to-back MMAs from an elected lane, commits to an mbarrier and waits; cycles / MMA are printed.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../visualbert_b200/csrc -o umma_rate umma_rate.cu
#include <cstdio>
#include "vb_common.cuh"
using namespace vb;

__device__ __forceinline__ uint32_t idesc(int m, int n, bool a_mn, bool b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__global__ void __launch_bounds__(128, 1) k(int n, int mode, int iters, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw), base = (raw + 1023u) & ~1023u;
    __shared__ uint64_t bar_storage;
    __shared__ uint32_t tptr;
    const int warp = threadIdx.x >> 5;
    const uint32_t bar = smem_u32(&bar_storage);
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(smem_u32(&tptr), 512);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tm = tptr;
    if (warp == 1) {
        const bool a_mn = mode == 2, b_mn = mode >= 1;
        const uint32_t id = idesc(128, n, a_mn, b_mn);
        const UmmaDesc da = make_umma_desc_sw128(base, a_mn ? 16384 : 0, 1024);
        const UmmaDesc db = make_umma_desc_sw128(base + 65536, 0, 1024);
        long long t0 = clock64();
        if (elect_one()) {
            for (int i = 0; i < iters; ++i) {
                const uint32_t ka = (i & 3) * (a_mn ? 2048 : 32), kb = (i & 3) * (b_mn ? 2048 : 32);
                if (mode == 3) umma_bf16_ts(tm + 256, tm + (i & 3) * 8, db.at((i & 3) * 2048), idesc(128, n, false, true), i > 0);
                else umma_bf16(tm + 256, da.at(ka), db.at(kb), id, i > 0);
            }
            umma_commit(bar);
        }
        __syncwarp();
        long long t1 = clock64();
        mbar_wait(bar, 0);
        long long t2 = clock64();
        if ((threadIdx.x & 31) == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) { tcgen05_fence_after(); tmem_dealloc(tm, 512); }
}

int main() {
    long long* out;
    cudaMallocManaged(&out, 16);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const char* names[4] = {"SS K-major A, K-major B", "SS K-major A, MN-major B", "SS MN-major A, MN-major B", "TS (A in TMEM), MN-major B"};
    for (int mode = 0; mode < 4; ++mode)
        for (int n : {16, 32, 64, 96, 128, 176, 256}) {
            for (int rep = 0; rep < 2; ++rep) { k<<<1, 128, 200 * 1024>>>(n, mode, 2048, out); cudaDeviceSynchronize(); }
            printf("%-28s N=%3d: issue %6.1f cyc/MMA, complete %6.1f cyc/MMA  (ideal %5.1f)  %s\n", names[mode], n, out[0] / 2048.0, out[1] / 2048.0,
                   128.0 * n / 256.0, cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
