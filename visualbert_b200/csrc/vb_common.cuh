// vb_common.cuh — sm_100a device primitives shared by every kernel in libvbert_b200.
//
// Thin inline-PTX wrappers (mbarrier, TMA, tcgen05/TMEM) plus small math helpers.
// No CUTLASS/CuTe dependency: the descriptors are built by hand (see vb_gemm.cu).
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

namespace vb {

typedef __nv_bfloat16 bf16;

// host-side shared state / helpers (defined in vb_gemm.cu)
extern std::atomic<long long> g_launches;  // kernel launches issued by this library
int num_sms();      // SM count of the CALLER'S CURRENT device (cached per device)
constexpr int kMaxDevices = 64;
int current_device();  // cudaGetDevice, clamped to [0, kMaxDevices)
// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device property of a kernel: `cache` is the caller's static
// per-device record of the size already configured (DataParallel threads / several devices in one process).
template <typename K>
inline cudaError_t ensure_dyn_smem(K kern, int bytes, int (&cache)[kMaxDevices]) {
    const int d = current_device();
    if (cache[d] >= bytes) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) cache[d] = bytes;
    return e;
}

// Optional live profiling (vb_profile_enable): every launcher brackets its kernels with CUDA events on
// the launch stream and tags them with a category and the algorithmic work (FLOPs or bytes) of the call.
enum { PROF_GEMM_FWD = 0, PROF_GEMM_DGRAD = 1, PROF_GEMM_WGRAD = 2, PROF_ATTN_FWD = 3, PROF_ATTN_DQ = 4, PROF_ATTN_DKV = 5,
       PROF_LN_FWD = 6, PROF_LN_BWD = 7, PROF_COLSUM = 8, PROF_EMBED = 9, PROF_OTHER = 10, PROF_NCAT = 11 };
// ---- programmatic dependent launch (PDL) -------------------------------------------------------------------
// Hot-path kernels are launched with cudaLaunchAttributeProgrammaticStreamSerialization: the next kernel's CTAs may be
// scheduled, and run their on-chip prologue (barrier init, TMEM allocation, tensor-map prefetch), while the previous
// kernel's last CTAs drain. Every such kernel calls pdl_trigger() on entry and pdl_wait() BEFORE its first global
// memory access (griddepcontrol.wait returns once the preceding grid has completed and its writes are visible).
// VB_PDL=0 in the environment turns the attribute off (plain stream order).
bool pdl_enabled();
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
// same, for a kernel launched as clusters of `cluster` CTAs along x (cluster size chosen at launch, not at compile time)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster,
                                      Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

struct ProfScope {
    ProfScope(cudaStream_t st, int cat, double work, int launches);
    ~ProfScope();
    int slot;
    cudaStream_t st;
};

// ---------------------------------------------------------------------------------------------
// error plumbing (host)
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
#define VB_CHECK_CUDA(expr)                                                                   \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            vb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return 1;                                                                         \
        }                                                                                     \
    } while (0)
#define VB_REQUIRE(cond, ...)                                                                 \
    do {                                                                                      \
        if (!(cond)) {                                                                        \
            vb::set_error(__VA_ARGS__);                                                       \
            return 2;                                                                         \
        }                                                                                     \
    } while (0)

// ---------------------------------------------------------------------------------------------
// generic helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t v) {
    __nv_bfloat162 h = *reinterpret_cast<__nv_bfloat162*>(&v);
    return __bfloat1622float2(h);
}

// raw SFU approximations (1 MUFU each, no range fix-up branches); ex2(-inf) = +0
__device__ __forceinline__ float fast_ex2(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float fast_rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// gelu(x) = x * 0.5 * (1 + erf(x / sqrt(2)))   — reference modeling.py:56-61 (exact-erf form).
// erf via Abramowitz-Stegun 7.1.26 (|abs err| < 1.5e-7 + SFU approximation error ~1e-6, far below the
// bf16 output rounding): erf(|z|) = 1 - (a1 t + ... + a5 t^5) exp(-z^2), t = 1 / (1 + p |z|).
// Returns q = poly(t) * t * exp(-x^2/2) so that erf(|x|/sqrt2) = 1 - q; also hands back exp(-x^2/2).
__device__ __forceinline__ float erfc_abs_sqrt2(float x, float& e) {
    const float ax = fabsf(x);
    e = fast_ex2(x * x * -0.72134752044448170f);              // exp(-x^2/2)
    const float t = fast_rcp(fmaf(0.23164189f, ax, 1.0f));    // p/sqrt(2) = 0.3275911 * 0.70710678
    float poly = fmaf(1.061405429f, t, -1.453152027f);
    poly = fmaf(poly, t, 1.421413741f);
    poly = fmaf(poly, t, -0.284496736f);
    poly = fmaf(poly, t, 0.254829592f);
    return poly * t * e;
}
__device__ __forceinline__ float gelu_fwd(float x) {
    float e;
    const float q = erfc_abs_sqrt2(x, e);
    const float hx = 0.5f * x, ha = fabsf(hx);
    return fmaf(-ha, q, hx + ha);  // 0.5 x + 0.5 |x| (1 - q)
}
// gelu(x) and d/dx gelu(x) = Phi(x) + x phi(x) from the same exp / rcp (the FFN-up epilogue stores both, so the
// backward epilogue is a single multiply)
__device__ __forceinline__ void gelu_fwd_bwd(float x, float& g, float& gp) {
    float e;
    const float q = erfc_abs_sqrt2(x, e);
    const float cdf = 0.5f + copysignf(fmaf(-0.5f, q, 0.5f), x);
    g = x * cdf;
    gp = fmaf(x * e, 0.3989422804014327f, cdf);
}
// d/dx gelu(x) = 0.5 (1 + erf(x/√2)) + x φ(x),  φ(x) = exp(-x²/2)/√(2π)
__device__ __forceinline__ float gelu_bwd(float x) {
    float e;
    const float q = erfc_abs_sqrt2(x, e);
    const float cdf = 0.5f + copysignf(fmaf(-0.5f, q, 0.5f), x);
    return fmaf(x * e, 0.3989422804014327f, cdf);
}

// ---------------------------------------------------------------------------------------------
// counter-based dropout RNG: a pure function of (seed, stream id, element index), so forward and
// backward regenerate the same keep-mask without storing it. One 32-bit avalanche hash (lowbias32
// finaliser, ~8 integer ops) yields two 16-bit uniforms, i.e. 4 ops per element — the epilogues that
// apply dropout are ALU-bound, a 10-round Philox (13 ops per element) measurably slowed them down.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du;
    x ^= x >> 15; x *= 0x846ca68bu;
    x ^= x >> 16;
    return x;
}
__device__ __forceinline__ uint32_t dropout_key(uint64_t seed, uint32_t stream) {
    return mix32(static_cast<uint32_t>(seed) ^ mix32(static_cast<uint32_t>(seed >> 32) + 0x9E3779B9u * (stream + 1u)));
}
// Hidden-state dropout quantises the drop probability to n/256 (0.1 -> 26/256 = 0.1016) and scales survivors by
// 256/(256-n), so E[dropout(x)] = x exactly while one hash serves 4 elements and the threshold test is a single
// SIMD byte compare. Host side: dropout_quantise().
struct DropQ { unsigned thr8; float scale; };
inline DropQ dropout_quantise(float p) {
    DropQ q;
    unsigned n = static_cast<unsigned>(p * 256.0f + 0.5f);
    if (n > 255u) n = 255u;
    q.thr8 = n;
    q.scale = p > 0.f ? 256.0f / (256.0f - static_cast<float>(n)) : 0.f;
    return q;
}
// Keep-mask bits for the 8 consecutive elements starting at flat element index `elem8 * 8`.
// Each element consumes 8 random bits; kept iff bits >= thr8.
__device__ __forceinline__ uint32_t dropout_keep8(uint64_t seed, uint32_t stream, uint64_t elem8, uint32_t thr8) {
    const uint32_t key = dropout_key(seed, stream) ^ (static_cast<uint32_t>(elem8 >> 31) * 0x27d4eb2fu);
    const uint32_t base = static_cast<uint32_t>(elem8) << 1;
    const uint32_t t4 = thr8 * 0x01010101u;
    // __vcmpgeu4: per-byte (a >= b) ? 0xff : 0x00; the multiply gathers the 4 byte LSBs into one nibble
    const uint32_t m0 = __vcmpgeu4(mix32(base ^ key), t4) & 0x01010101u;
    const uint32_t m1 = __vcmpgeu4(mix32((base + 1u) ^ key), t4) & 0x01010101u;
    return ((m0 * 0x01020408u) >> 24) | (((m1 * 0x01020408u) >> 24) << 4);
}

// ---------------------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug traps (the kernel fails with an error the host reports) instead of hanging the GPU:
// ~4e9 cycles ≈ 2 s at 1.9 GHz, far beyond any legitimate wait in these kernels. Kept tiny and fully inline — the
// hot kernels wait at dozens of sites (code size = instruction-cache misses on every role switch), and a shared
// out-of-line helper would make ptxas apply the SMALLEST setmaxnreg budget of its callers to all of them.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 0x3ffu) == 0 && clock64() - t0 > 4000000000ll) __trap();
    }
}

// UMMA shared-memory descriptor split in two 32-bit halves: only the 14-bit start-address field (bits 0-13, in 16-byte
// units) changes between the MMAs of a tile, so stepping through k is ONE 32-bit add on the low half — the MMA-issuing
// thread is a single thread whose instruction latency is exposed, every instruction saved there shortens the pipeline.
struct UmmaDesc {
    uint32_t lo, hi;
    __device__ __forceinline__ uint64_t at(uint32_t byte_off) const {
        return (static_cast<uint64_t>(hi) << 32) | static_cast<uint64_t>(lo + (byte_off >> 4));
    }
};
__device__ __forceinline__ UmmaDesc make_umma_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    UmmaDesc d;
    d.lo = ((saddr & 0x3FFFFu) >> 4) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
    d.hi = ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);   // version 1 (bit 46), SWIZZLE_128B (bits 61-63)
    return d;
}

// ---------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor) — 2-D tiled load, completion on an mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* m, uint32_t bar,
                                            int c_inner, int c_outer) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c_inner), "r"(c_outer)
        : "memory");
}

// ---------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tcgen05_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// whole warp; writes the TMEM base address (lane<<16 | column) to *smem_dst
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
                 : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; issued by ONE thread on behalf of the CTA.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                          uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// mbarrier arrives when all previously issued tcgen05.mma of this thread have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
                 : "memory");
}
// ---- cta_group::2 (CTA pair) variants -------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t cta) {
    asm volatile(
        "{\n"
        ".reg .b32 ra;\n"
        "mapa.shared::cluster.u32 ra, %0, %1;\n"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n"
        "}\n" ::"r"(bar), "r"(cta)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// TMA load issued by either CTA of the pair; the transaction bytes are credited to the LEADER CTA's mbarrier
// (peer bit of the barrier address cleared, as CUTLASS' SM100_TMA_2SM_LOAD does)
__device__ __forceinline__ void tma_load_2d_2cta(uint32_t smem_dst, const CUtensorMap* m, uint32_t bar, int c_inner, int c_outer) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar & 0xFEFFFFFFu), "r"(c_inner), "r"(c_outer)
        : "memory");
}
// D[tmem of both CTAs] (+)= A * B with M = 256 split over the CTA pair; issued by ONE thread of the leader CTA
__device__ __forceinline__ void umma_bf16_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// commit: when the MMAs issued so far retire, arrive on the mbarrier at this offset in every CTA of `cta_mask`
// (bit i = CTA rank i of the cluster; 3 = both CTAs of the first pair)
__device__ __forceinline__ void umma_commit_2cta(uint32_t bar, uint32_t cta_mask = 3u) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"(static_cast<uint16_t>(cta_mask))
                 : "memory");
}
// TMA store of one box from (swizzled) shared memory; completion is tracked per thread in bulk async-groups
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t smem_src, int c_inner, int c_outer) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_src), "r"(c_inner), "r"(c_outer)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// the issuing thread's bulk groups have finished READING shared memory (the source may be overwritten)
__device__ __forceinline__ void bulk_wait_group_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_group0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void sts_v4(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// TMA load multicast to the CTAs of `cta_mask`: the box lands at the same CTA-relative offset in each of them and the
// bytes are credited to the barrier at this offset in the LEADER of each destination CTA's pair (peer bit cleared)
__device__ __forceinline__ void tma_load_2d_2cta_mc(uint32_t smem_dst, const CUtensorMap* m, uint32_t bar, int c_inner, int c_outer,
                                                    uint32_t cta_mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
        " [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar & 0xFEFFFFFFu), "r"(c_inner), "r"(c_outer),
          "h"(static_cast<uint16_t>(cta_mask))
        : "memory");
}

// 32 lanes x 32 columns of 32-bit: thread i of the warp receives row (lane base + i), 32 columns.
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
          "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
          "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
          "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
// A operand from TMEM (TS form): D[tmem] (+)= A[tmem, K-major: lane = row, 2 bf16 per 32-bit column] * B[smem desc]
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 32 lanes x 8 columns of 32-bit: thread i of the warp writes row (lane base + i), 8 consecutive columns
__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
                 "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// One lane of a CONVERGED warp (elect.sync). The MMA-issuing role runs its loop with all 32 lanes converged and predicates
// only the tcgen05 instructions on this: the compiler then keeps descriptors and loop state in uniform registers. Guarding
// the whole role with `if (lane == 0)` instead makes the code divergent, and every tcgen05.mma gets wrapped in an
// ELECT / BRA.U.ANY serialisation loop (~100+ cycles of issue latency per MMA — measured in the attention kernels).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "elect.sync _|P1, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}
// Register re-distribution between warp roles (whole warpgroups of 4 warps): the control warpgroup shrinks its
// allocation, the element-wise warpgroups grow theirs; ptxas compiles the code that follows against the new limit.
template <int N>
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------
// vectorised global access
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 ldg_v4(const void* p) {
    return __ldg(reinterpret_cast<const uint4*>(p));
}
__device__ __forceinline__ void stg_v4(void* p, uint4 v) { *reinterpret_cast<uint4*>(p) = v; }
// 256-bit global access (sm_100): one full 32-byte sector per thread per instruction
__device__ __forceinline__ void ldg_v8(const void* p, uint32_t (&r)[8]) {
    asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "l"(p));
}
__device__ __forceinline__ void stg_v8(void* p, const uint32_t (&r)[8]) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void red_add_v4_f32(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c),
                 "f"(d)
                 : "memory");
}

}  // namespace vb
