// Microbenchmark: TMEM -> register bandwidth of tcgen05.ld.32x32b.{x16,x32} per SM as a function of the number
// of reading warps. One CTA per SM, 512 TMEM columns allocated, every warp re-reads its lane quarter.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int X>
__device__ __forceinline__ uint32_t ld(uint32_t taddr) {
    uint32_t acc = 0;
    if constexpr (X == 16) {
        uint32_t v[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(v[0]),"=r"(v[1]),"=r"(v[2]),"=r"(v[3]),"=r"(v[4]),"=r"(v[5]),"=r"(v[6]),"=r"(v[7]),"=r"(v[8]),"=r"(v[9]),"=r"(v[10]),"=r"(v[11]),"=r"(v[12]),"=r"(v[13]),"=r"(v[14]),"=r"(v[15]) : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 16; ++i) acc ^= v[i];
    } else {
        uint32_t v[32];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(v[0]),"=r"(v[1]),"=r"(v[2]),"=r"(v[3]),"=r"(v[4]),"=r"(v[5]),"=r"(v[6]),"=r"(v[7]),"=r"(v[8]),"=r"(v[9]),"=r"(v[10]),"=r"(v[11]),"=r"(v[12]),"=r"(v[13]),"=r"(v[14]),"=r"(v[15]),
              "=r"(v[16]),"=r"(v[17]),"=r"(v[18]),"=r"(v[19]),"=r"(v[20]),"=r"(v[21]),"=r"(v[22]),"=r"(v[23]),"=r"(v[24]),"=r"(v[25]),"=r"(v[26]),"=r"(v[27]),"=r"(v[28]),"=r"(v[29]),"=r"(v[30]),"=r"(v[31]) : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 32; ++i) acc ^= v[i];
    }
    return acc;
}
__device__ __forceinline__ uint32_t ld16x4(uint32_t taddr) {  // 4 loads in flight before one wait
    uint32_t v[4][16];
#pragma unroll
    for (int j = 0; j < 4; ++j)
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(v[j][0]),"=r"(v[j][1]),"=r"(v[j][2]),"=r"(v[j][3]),"=r"(v[j][4]),"=r"(v[j][5]),"=r"(v[j][6]),"=r"(v[j][7]),"=r"(v[j][8]),"=r"(v[j][9]),"=r"(v[j][10]),"=r"(v[j][11]),"=r"(v[j][12]),"=r"(v[j][13]),"=r"(v[j][14]),"=r"(v[j][15]) : "r"(taddr + j * 16) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    uint32_t acc = 0;
    for (int j = 0; j < 4; ++j) for (int i = 0; i < 16; ++i) acc ^= v[j][i];
    return acc;
}
template <int MODE>
__global__ void k(uint32_t* out, long long* cyc, int iters) {
    __shared__ uint32_t tptr;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tptr)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = tptr + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const uint32_t col = (it * 64) & 255;
        if (MODE == 0) acc ^= ld<16>(base + col);
        else if (MODE == 1) acc ^= ld<32>(base + col);
        else acc ^= ld16x4(base + col);
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    if (acc == 0x12345u) out[0] = acc;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tptr), "r"(512) : "memory");
}
template <int MODE>
void run(int warps, const char* name, int bytes_per_ld) {
    uint32_t* d; long long* c; cudaMalloc(&d, 4); cudaMalloc(&c, 8 * 148);
    const int iters = 20000;
    k<MODE><<<148, warps * 32>>>(d, c, 100);
    k<MODE><<<148, warps * 32>>>(d, c, iters);
    long long h[148]; cudaMemcpy(h, c, sizeof(h), cudaMemcpyDeviceToHost);
    cudaError_t e = cudaDeviceSynchronize();
    double cyc = (double)h[0];
    printf("%-10s warps=%2d: %.1f B/clk/SM (%.0f cycles per ld round) %s\n", name, warps, (double)bytes_per_ld * warps * iters / cyc, cyc / iters,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(d); cudaFree(c);
}
int main() {
    for (int w : {4, 8, 16}) { run<0>(w, "x16", 2048); run<1>(w, "x32", 4096); run<2>(w, "x16 x4pipe", 8192); }
    return 0;
}
