// vb_attention.cu — fused multi-head self-attention over the [text ; visual] sequence, fwd + bwd.
//
// Replaces BertSelfAttention.forward (reference modeling.py:231-261): scores = QK^T / sqrt(d), then
// + additive mask ((1-mask) * -10000, modeling.py:1293-1294), softmax over keys, dropout on the
// probabilities (modeling.py:251), context = P V, heads merged back to [B*S, H] — without ever
// materialising the [B, A, S, S] score tensor the reference reads/writes five times per layer.
//
// Layout: Q, K, V are column slices of the fused projection output qkv[B*S, 3H] (Q | K | V, head h at
// columns h*64 .. h*64+63 of each third); head_dim is fixed at 64 (BERT-base and -large).
// One CTA = (batch b, head h, 64 query rows); 4 warps x 16 rows; K/V streamed in 64-key blocks through
// a double-buffered, XOR-swizzled shared-memory ring (cp.async); scores, probabilities and the running
// max / sum live in registers (flash-style online softmax in the exp2 domain).
// Backward recomputes P from the saved log-sum-exp: kernel A (per query block) produces dQ and the
// row term D = rowsum(dO * O); kernel B (per key block) produces dK and dV. No atomics, deterministic.
//
// Tensor-core path here is warp-level mma.sync (m16n8k16, bf16 -> fp32); the tcgen05 budget of the
// layer is spent in vb_gemm.cu where > 96 % of the FLOPs are.
#include "../../include/vbert_b200.h"
#include "vb_common.cuh"

namespace vb {

constexpr int kHd = 64;             // head dim
constexpr int kBlk = 64;            // rows per tile (queries or keys)
constexpr int kTileBytes = kBlk * kHd * 2;
constexpr float kLog2e = 1.4426950408889634f;

struct AttnParams {
    const bf16* qkv;   // [B*S, 3H]
    const float* mask_bias;  // [B, S] additive key bias, natural-log domain ((1-mask) * -10000)
    bf16* ctx;         // [B*S, H]        (fwd out / bwd: O)
    float* lse;        // [B, A, S]       natural-log domain
    const bf16* dctx;  // [B*S, H]        (bwd)
    bf16* dqkv;        // [B*S, 3H]       (bwd out)
    float* drow;       // [B, A, S]       (bwd scratch: rowsum(dO * O))
    int B, S, A, H;
    float scale;       // 1/sqrt(head_dim)
    float drop_scale;  // 1/(1-p) or 0
    unsigned drop_thresh16;
    unsigned drop_seed;
};

// ------------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t swz(int row, int chunk) {  // byte offset inside a 64x64 bf16 tile
    return static_cast<uint32_t>(row * 128 + ((chunk ^ (row & 7)) << 4));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    const int sz = valid ? 16 : 0;  // src-size 0 => zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// 64 x 64 bf16 tile: rows row0..row0+63 of a [*, ld] matrix starting at column col0; rows >= nrows -> 0
__device__ __forceinline__ void load_tile(uint32_t tile, const bf16* base, long long ld, int row0, int nrows,
                                          int tid) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int idx = tid + i * 128;
        const int r = idx >> 3, c = idx & 7;
        const bool ok = (row0 + r) < nrows;
        const bf16* src = base + static_cast<long long>(ok ? row0 + r : 0) * ld + c * 8;
        cp_async16(tile + swz(r, c), src, ok);
    }
}

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
        "{%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// A fragments (16 rows x 64 k) of rows r0..r0+15 of a swizzled [64][64] tile
__device__ __forceinline__ void load_afrag(uint32_t (&a)[4][4], uint32_t tile, int r0, int lane) {
    const int row = r0 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        const int chunk = ks * 2 + (lane >> 4);
        ldsm_x4(tile + swz(row, chunk), a[ks][0], a[ks][1], a[ks][2], a[ks][3]);
    }
}
// acc(16 x 64 n) += A(16 x 64 k) * T^T, T = tile [n=64][k=64] row-major
__device__ __forceinline__ void gemm_nt(float (&acc)[8][4], const uint32_t (&a)[4][4], uint32_t tile, int lane) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
        for (int np = 0; np < 4; ++np) {
            const int row = np * 16 + (lane & 7) + (lane >> 4) * 8;
            const int chunk = ks * 2 + ((lane >> 3) & 1);
            uint32_t b0, b1, b2, b3;
            ldsm_x4(tile + swz(row, chunk), b0, b1, b2, b3);
            mma16816(acc[2 * np], a[ks], b0, b1);
            mma16816(acc[2 * np + 1], a[ks], b2, b3);
        }
    }
}
// acc(16 x 64 n) += A(16 x 64 k) * T, T = tile [k=64][n=64] row-major
__device__ __forceinline__ void gemm_nn(float (&acc)[8][4], const uint32_t (&a)[4][4], uint32_t tile, int lane) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
        for (int np = 0; np < 4; ++np) {
            const int row = ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
            const int chunk = np * 2 + (lane >> 4);
            uint32_t b0, b1, b2, b3;
            ldsm_x4_t(tile + swz(row, chunk), b0, b1, b2, b3);
            mma16816(acc[2 * np], a[ks], b0, b1);
            mma16816(acc[2 * np + 1], a[ks], b2, b3);
        }
    }
}
// accumulator tile (16 x 64, fp32) -> A fragments (bf16) for the next GEMM
__device__ __forceinline__ void acc_to_afrag(uint32_t (&a)[4][4], const float (&p)[8][4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        a[j][0] = pack_bf16x2(p[2 * j][0], p[2 * j][1]);
        a[j][1] = pack_bf16x2(p[2 * j][2], p[2 * j][3]);
        a[j][2] = pack_bf16x2(p[2 * j + 1][0], p[2 * j + 1][1]);
        a[j][3] = pack_bf16x2(p[2 * j + 1][2], p[2 * j + 1][3]);
    }
}
__device__ __forceinline__ void zero_acc(float (&c)[8][4]) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
}
// keep decision for attention-probability dropout, a pure function of (head-row index, key)
__device__ __forceinline__ bool attn_keep(unsigned seed, unsigned bh, int q, int key, int S, unsigned thresh16) {
    unsigned x = (bh * static_cast<unsigned>(S) + static_cast<unsigned>(q)) * static_cast<unsigned>(S) +
                 static_cast<unsigned>(key);
    x ^= seed;
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return (x >> 16) >= thresh16;
}
// store a 16 x 64 accumulator tile as bf16 rows of a [*, ld] matrix (rows >= nrows skipped)
__device__ __forceinline__ void store_acc(bf16* base, long long ld, int row0, int nrows, const float (&c)[8][4],
                                          int lane, float mul0, float mul1) {
    const int g = lane >> 2, t = lane & 3;
    const int ra = row0 + g, rb = row0 + g + 8;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        const int col = nt * 8 + 2 * t;
        if (ra < nrows)
            *reinterpret_cast<uint32_t*>(base + static_cast<long long>(ra) * ld + col) =
                pack_bf16x2(c[nt][0] * mul0, c[nt][1] * mul0);
        if (rb < nrows)
            *reinterpret_cast<uint32_t*>(base + static_cast<long long>(rb) * ld + col) =
                pack_bf16x2(c[nt][2] * mul1, c[nt][3] * mul1);
    }
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
attn_fwd_kernel(const AttnParams p) {
    __shared__ __align__(128) uint8_t smem[5 * kTileBytes];
    __shared__ float sbias[kBlk * 2];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int S = p.S;
    const long long ld = 3LL * p.H;
    const bf16* qbase = p.qkv + static_cast<long long>(b) * S * ld + h * kHd;
    const bf16* kbase = qbase + p.H;
    const bf16* vbase = qbase + 2 * p.H;
    const uint32_t sQ = smem_u32(smem), sK0 = sQ + kTileBytes, sV0 = sQ + 3 * kTileBytes;
    const int nkb = (S + kBlk - 1) / kBlk;

    load_tile(sQ, qbase, ld, qb * kBlk, S, tid);
    load_tile(sK0, kbase, ld, 0, S, tid);
    load_tile(sV0, vbase, ld, 0, S, tid);
    cp_async_commit();

    const float sc2 = p.scale * kLog2e;
    float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};
    float o[8][4];
    zero_acc(o);
    uint32_t qf[4][4];
    const unsigned bh = static_cast<unsigned>(b * p.A + h);
    const int qrow0 = qb * kBlk + warp * 16;

    for (int kb = 0; kb < nkb; ++kb) {
        const int buf = kb & 1;
        if (tid < kBlk) {
            const int key = kb * kBlk + tid;
            sbias[buf * kBlk + tid] = key < S ? p.mask_bias[static_cast<long long>(b) * S + key] * kLog2e : -INFINITY;
        }
        if (kb + 1 < nkb) {
            load_tile(sK0 + (buf ^ 1) * kTileBytes, kbase, ld, (kb + 1) * kBlk, S, tid);
            load_tile(sV0 + (buf ^ 1) * kTileBytes, vbase, ld, (kb + 1) * kBlk, S, tid);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (kb == 0) load_afrag(qf, sQ, warp * 16, lane);

        float s[8][4];
        zero_acc(s);
        gemm_nt(s, qf, sK0 + buf * kTileBytes, lane);
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const float b0 = sbias[buf * kBlk + nt * 8 + 2 * t], b1 = sbias[buf * kBlk + nt * 8 + 2 * t + 1];
            s[nt][0] = fmaf(s[nt][0], sc2, b0); s[nt][1] = fmaf(s[nt][1], sc2, b1);
            s[nt][2] = fmaf(s[nt][2], sc2, b0); s[nt][3] = fmaf(s[nt][3], sc2, b1);
            mx[0] = fmaxf(mx[0], fmaxf(s[nt][0], s[nt][1]));
            mx[1] = fmaxf(mx[1], fmaxf(s[nt][2], s[nt][3]));
        }
        float alpha[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
            const float mn = fmaxf(m[r], mx[r]);
            alpha[r] = fast_ex2(m[r] - mn);
            m[r] = mn;
        }
        float rs[2] = {0.f, 0.f};
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            s[nt][0] = fast_ex2(s[nt][0] - m[0]); s[nt][1] = fast_ex2(s[nt][1] - m[0]);
            s[nt][2] = fast_ex2(s[nt][2] - m[1]); s[nt][3] = fast_ex2(s[nt][3] - m[1]);
            rs[0] += s[nt][0] + s[nt][1];
            rs[1] += s[nt][2] + s[nt][3];
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], 1);
            rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], 2);
            l[r] = l[r] * alpha[r] + rs[r];
        }
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            o[nt][0] *= alpha[0]; o[nt][1] *= alpha[0];
            o[nt][2] *= alpha[1]; o[nt][3] *= alpha[1];
        }
        if (p.drop_scale != 0.f) {
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                const int key = kb * kBlk + nt * 8 + 2 * t;
                const int qa = qrow0 + g, qc = qrow0 + g + 8;
                s[nt][0] = attn_keep(p.drop_seed, bh, qa, key, S, p.drop_thresh16) ? s[nt][0] * p.drop_scale : 0.f;
                s[nt][1] = attn_keep(p.drop_seed, bh, qa, key + 1, S, p.drop_thresh16) ? s[nt][1] * p.drop_scale : 0.f;
                s[nt][2] = attn_keep(p.drop_seed, bh, qc, key, S, p.drop_thresh16) ? s[nt][2] * p.drop_scale : 0.f;
                s[nt][3] = attn_keep(p.drop_seed, bh, qc, key + 1, S, p.drop_thresh16) ? s[nt][3] * p.drop_scale : 0.f;
            }
        }
        uint32_t pf[4][4];
        acc_to_afrag(pf, s);
        gemm_nn(o, pf, sV0 + buf * kTileBytes, lane);
        __syncthreads();  // all warps done with this K/V buffer before it is refilled
    }
    const float inv0 = 1.f / l[0], inv1 = 1.f / l[1];
    store_acc(p.ctx + static_cast<long long>(b) * S * p.H + h * kHd, p.H, qrow0, S, o, lane, inv0, inv1);
    if (t == 0 && p.lse != nullptr) {
        float* lse = p.lse + (static_cast<long long>(b) * p.A + h) * S;
        if (qrow0 + g < S) lse[qrow0 + g] = (m[0] + log2f(l[0])) * 0.6931471805599453f;
        if (qrow0 + g + 8 < S) lse[qrow0 + g + 8] = (m[1] + log2f(l[1])) * 0.6931471805599453f;
    }
}

// ------------------------------------------------------------------------------------------------
// backward A: per query block — D = rowsum(dO * O), dQ = scale * sum_k dS K
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
attn_bwd_dq_kernel(const AttnParams p) {
    extern __shared__ __align__(128) uint8_t dsmem[];
    __shared__ float sbias[kBlk * 2];
    __shared__ float sD[kBlk];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int S = p.S;
    const long long ld = 3LL * p.H;
    const bf16* qbase = p.qkv + static_cast<long long>(b) * S * ld + h * kHd;
    const bf16* kbase = qbase + p.H;
    const bf16* vbase = qbase + 2 * p.H;
    const bf16* obase = p.ctx + static_cast<long long>(b) * S * p.H + h * kHd;
    const bf16* dobase = p.dctx + static_cast<long long>(b) * S * p.H + h * kHd;
    // tiles: Q, dO, O, K[2], V[2]
    const uint32_t sQ = smem_u32(dsmem), sdO = sQ + kTileBytes, sO = sQ + 2 * kTileBytes;
    const uint32_t sK0 = sQ + 3 * kTileBytes, sV0 = sQ + 5 * kTileBytes;
    const int nkb = (S + kBlk - 1) / kBlk;

    load_tile(sQ, qbase, ld, qb * kBlk, S, tid);
    load_tile(sdO, dobase, p.H, qb * kBlk, S, tid);
    load_tile(sO, obase, p.H, qb * kBlk, S, tid);
    load_tile(sK0, kbase, ld, 0, S, tid);
    load_tile(sV0, vbase, ld, 0, S, tid);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    {   // D[row] = sum_d dO * O : two threads per row
        const int r = tid >> 1, half = tid & 1;
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int chunk = half * 4 + c;
            const uint4 a = *reinterpret_cast<const uint4*>(dsmem + kTileBytes + swz(r, chunk));
            const uint4 o = *reinterpret_cast<const uint4*>(dsmem + 2 * kTileBytes + swz(r, chunk));
            const uint32_t av[4] = {a.x, a.y, a.z, a.w}, ov[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 x = unpack_bf16x2(av[i]), y = unpack_bf16x2(ov[i]);
                acc += x.x * y.x + x.y * y.y;
            }
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        if (half == 0) {
            sD[r] = acc;
            const int q = qb * kBlk + r;
            if (q < S) p.drow[(static_cast<long long>(b) * p.A + h) * S + q] = acc;
        }
    }
    uint32_t qf[4][4], dof[4][4];
    load_afrag(qf, sQ, warp * 16, lane);
    load_afrag(dof, sdO, warp * 16, lane);
    const int qrow0 = qb * kBlk + warp * 16;
    const float* lsep = p.lse + (static_cast<long long>(b) * p.A + h) * S;
    const float lse0 = (qrow0 + g < S) ? lsep[qrow0 + g] * kLog2e : 0.f;
    const float lse1 = (qrow0 + g + 8 < S) ? lsep[qrow0 + g + 8] * kLog2e : 0.f;
    __syncthreads();
    const float d0 = sD[warp * 16 + g], d1 = sD[warp * 16 + g + 8];
    const float sc2 = p.scale * kLog2e;
    const unsigned bh = static_cast<unsigned>(b * p.A + h);
    float dq[8][4];
    zero_acc(dq);

    for (int kb = 0; kb < nkb; ++kb) {
        const int buf = kb & 1;
        if (tid < kBlk) {
            const int key = kb * kBlk + tid;
            sbias[buf * kBlk + tid] = key < S ? p.mask_bias[static_cast<long long>(b) * S + key] * kLog2e : -INFINITY;
        }
        if (kb + 1 < nkb) {
            load_tile(sK0 + (buf ^ 1) * kTileBytes, kbase, ld, (kb + 1) * kBlk, S, tid);
            load_tile(sV0 + (buf ^ 1) * kTileBytes, vbase, ld, (kb + 1) * kBlk, S, tid);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        float s[8][4], dp[8][4];
        zero_acc(s);
        zero_acc(dp);
        gemm_nt(s, qf, sK0 + buf * kTileBytes, lane);
        gemm_nt(dp, dof, sV0 + buf * kTileBytes, lane);
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const float b0 = sbias[buf * kBlk + nt * 8 + 2 * t], b1 = sbias[buf * kBlk + nt * 8 + 2 * t + 1];
            const float p0 = fast_ex2(fmaf(s[nt][0], sc2, b0) - lse0), p1 = fast_ex2(fmaf(s[nt][1], sc2, b1) - lse0);
            const float p2 = fast_ex2(fmaf(s[nt][2], sc2, b0) - lse1), p3 = fast_ex2(fmaf(s[nt][3], sc2, b1) - lse1);
            float e0 = dp[nt][0], e1 = dp[nt][1], e2 = dp[nt][2], e3 = dp[nt][3];
            if (p.drop_scale != 0.f) {
                const int key = kb * kBlk + nt * 8 + 2 * t;
                const int qa = qrow0 + g, qc = qrow0 + g + 8;
                e0 = attn_keep(p.drop_seed, bh, qa, key, S, p.drop_thresh16) ? e0 * p.drop_scale : 0.f;
                e1 = attn_keep(p.drop_seed, bh, qa, key + 1, S, p.drop_thresh16) ? e1 * p.drop_scale : 0.f;
                e2 = attn_keep(p.drop_seed, bh, qc, key, S, p.drop_thresh16) ? e2 * p.drop_scale : 0.f;
                e3 = attn_keep(p.drop_seed, bh, qc, key + 1, S, p.drop_thresh16) ? e3 * p.drop_scale : 0.f;
            }
            s[nt][0] = p0 * (e0 - d0); s[nt][1] = p1 * (e1 - d0);
            s[nt][2] = p2 * (e2 - d1); s[nt][3] = p3 * (e3 - d1);
        }
        uint32_t dsf[4][4];
        acc_to_afrag(dsf, s);
        gemm_nn(dq, dsf, sK0 + buf * kTileBytes, lane);
        __syncthreads();
    }
    store_acc(p.dqkv + static_cast<long long>(b) * S * ld + h * kHd, ld, qrow0, S, dq, lane, p.scale, p.scale);
}

// ------------------------------------------------------------------------------------------------
// backward B: per key block — dV = P_drop^T dO, dK = scale * dS^T Q
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
attn_bwd_dkv_kernel(const AttnParams p) {
    extern __shared__ __align__(128) uint8_t dsmem[];
    __shared__ float slse[kBlk * 2];
    __shared__ float sD[kBlk * 2];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int kbk = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int S = p.S;
    const long long ld = 3LL * p.H;
    const bf16* qbase = p.qkv + static_cast<long long>(b) * S * ld + h * kHd;
    const bf16* kbase = qbase + p.H;
    const bf16* vbase = qbase + 2 * p.H;
    const bf16* dobase = p.dctx + static_cast<long long>(b) * S * p.H + h * kHd;
    // tiles: K, V, Q[2], dO[2]
    const uint32_t sK = smem_u32(dsmem), sV = sK + kTileBytes, sQ0 = sK + 2 * kTileBytes, sdO0 = sK + 4 * kTileBytes;
    const int nqb = (S + kBlk - 1) / kBlk;
    const float* lsep = p.lse + (static_cast<long long>(b) * p.A + h) * S;
    const float* drp = p.drow + (static_cast<long long>(b) * p.A + h) * S;

    load_tile(sK, kbase, ld, kbk * kBlk, S, tid);
    load_tile(sV, vbase, ld, kbk * kBlk, S, tid);
    load_tile(sQ0, qbase, ld, 0, S, tid);
    load_tile(sdO0, dobase, p.H, 0, S, tid);
    cp_async_commit();

    const int krow0 = kbk * kBlk + warp * 16;
    const int ka = krow0 + g, kc = krow0 + g + 8;
    const float bias0 = ka < S ? p.mask_bias[static_cast<long long>(b) * S + ka] * kLog2e : -INFINITY;
    const float bias1 = kc < S ? p.mask_bias[static_cast<long long>(b) * S + kc] * kLog2e : -INFINITY;
    const float sc2 = p.scale * kLog2e;
    const unsigned bh = static_cast<unsigned>(b * p.A + h);
    uint32_t kf[4][4], vf[4][4];
    float dk[8][4], dv[8][4];
    zero_acc(dk);
    zero_acc(dv);

    for (int qb = 0; qb < nqb; ++qb) {
        const int buf = qb & 1;
        if (tid < kBlk) {
            const int q = qb * kBlk + tid;
            slse[buf * kBlk + tid] = q < S ? lsep[q] * kLog2e : INFINITY;  // +inf => p = 0 for padded queries
            sD[buf * kBlk + tid] = q < S ? drp[q] : 0.f;
        }
        if (qb + 1 < nqb) {
            load_tile(sQ0 + (buf ^ 1) * kTileBytes, qbase, ld, (qb + 1) * kBlk, S, tid);
            load_tile(sdO0 + (buf ^ 1) * kTileBytes, dobase, p.H, (qb + 1) * kBlk, S, tid);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (qb == 0) {
            load_afrag(kf, sK, warp * 16, lane);
            load_afrag(vf, sV, warp * 16, lane);
        }
        float st[8][4], dpt[8][4];
        zero_acc(st);
        zero_acc(dpt);
        gemm_nt(st, kf, sQ0 + buf * kTileBytes, lane);    // S^T  = K Q^T   (16 keys x 64 queries)
        gemm_nt(dpt, vf, sdO0 + buf * kTileBytes, lane);  // dP^T = V dO^T
        float pt[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const int qi = nt * 8 + 2 * t;
            const float l0 = slse[buf * kBlk + qi], l1 = slse[buf * kBlk + qi + 1];
            const float dd0 = sD[buf * kBlk + qi], dd1 = sD[buf * kBlk + qi + 1];
            const float p0 = fast_ex2(fmaf(st[nt][0], sc2, bias0) - l0), p1 = fast_ex2(fmaf(st[nt][1], sc2, bias0) - l1);
            const float p2 = fast_ex2(fmaf(st[nt][2], sc2, bias1) - l0), p3 = fast_ex2(fmaf(st[nt][3], sc2, bias1) - l1);
            float e0 = dpt[nt][0], e1 = dpt[nt][1], e2 = dpt[nt][2], e3 = dpt[nt][3];
            float w0 = p0, w1 = p1, w2 = p2, w3 = p3;  // dropped probabilities feeding dV
            if (p.drop_scale != 0.f) {
                const int q = qb * kBlk + qi;
                const bool k0 = attn_keep(p.drop_seed, bh, q, ka, S, p.drop_thresh16);
                const bool k1 = attn_keep(p.drop_seed, bh, q + 1, ka, S, p.drop_thresh16);
                const bool k2 = attn_keep(p.drop_seed, bh, q, kc, S, p.drop_thresh16);
                const bool k3 = attn_keep(p.drop_seed, bh, q + 1, kc, S, p.drop_thresh16);
                e0 = k0 ? e0 * p.drop_scale : 0.f; w0 = k0 ? w0 * p.drop_scale : 0.f;
                e1 = k1 ? e1 * p.drop_scale : 0.f; w1 = k1 ? w1 * p.drop_scale : 0.f;
                e2 = k2 ? e2 * p.drop_scale : 0.f; w2 = k2 ? w2 * p.drop_scale : 0.f;
                e3 = k3 ? e3 * p.drop_scale : 0.f; w3 = k3 ? w3 * p.drop_scale : 0.f;
            }
            pt[nt][0] = w0; pt[nt][1] = w1; pt[nt][2] = w2; pt[nt][3] = w3;
            st[nt][0] = p0 * (e0 - dd0); st[nt][1] = p1 * (e1 - dd1);
            st[nt][2] = p2 * (e2 - dd0); st[nt][3] = p3 * (e3 - dd1);
        }
        uint32_t af[4][4];
        acc_to_afrag(af, pt);
        gemm_nn(dv, af, sdO0 + buf * kTileBytes, lane);  // dV += P^T dO
        acc_to_afrag(af, st);
        gemm_nn(dk, af, sQ0 + buf * kTileBytes, lane);   // dK += dS^T Q
        __syncthreads();
    }
    bf16* dbase = p.dqkv + static_cast<long long>(b) * S * ld + h * kHd;
    store_acc(dbase + p.H, ld, krow0, S, dk, lane, p.scale, p.scale);
    store_acc(dbase + 2 * p.H, ld, krow0, S, dv, lane, 1.f, 1.f);
}

// ------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------
static int fill_params(AttnParams& p, const void* qkv, const float* mask_bias, void* ctx, float* lse,
                       const void* dctx, void* dqkv, float* drow, int B, int S, int A, int H, float dropout_p,
                       unsigned long long seed, unsigned stream_id) {
    VB_REQUIRE(B > 0 && S > 0 && A > 0, "attention: empty problem");
    VB_REQUIRE(H == A * kHd, "attention: head_dim must be 64 (hidden=%d heads=%d)", H, A);
    VB_REQUIRE(A <= 65535 && B <= 65535, "attention: grid too large");
    VB_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, "attention: dropout_p out of range");
    p.qkv = static_cast<const bf16*>(qkv);
    p.mask_bias = mask_bias;
    p.ctx = static_cast<bf16*>(ctx);
    p.lse = lse;
    p.dctx = static_cast<const bf16*>(dctx);
    p.dqkv = static_cast<bf16*>(dqkv);
    p.drow = drow;
    p.B = B; p.S = S; p.A = A; p.H = H;
    p.scale = 0.125f;
    p.drop_scale = dropout_p > 0.f ? 1.f / (1.f - dropout_p) : 0.f;
    p.drop_thresh16 = static_cast<unsigned>(dropout_p * 65536.f + 0.5f);
    // fold the per-layer stream id into the 32-bit seed of the element hash
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (stream_id + 1ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    p.drop_seed = static_cast<unsigned>(z ^ (z >> 31));
    return 0;
}

int attn_fwd(const void* qkv, const float* mask_bias, void* ctx, float* lse, int B, int S, int A, int H,
             float dropout_p, unsigned long long seed, unsigned stream_id, cudaStream_t st) {
    AttnParams p;
    int rc = fill_params(p, qkv, mask_bias, ctx, lse, nullptr, nullptr, nullptr, B, S, A, H, dropout_p, seed, stream_id);
    if (rc) return rc;
    dim3 grid((S + kBlk - 1) / kBlk, A, B);
    {
        ProfScope ps(st, PROF_ATTN, 4.0 * B * A * S * S * kHd, 1);
        attn_fwd_kernel<<<grid, 128, 0, st>>>(p);
    }
    VB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int attn_bwd(const void* qkv, const float* mask_bias, const void* ctx, const float* lse, const void* dctx,
             void* dqkv, float* drow, int B, int S, int A, int H, float dropout_p, unsigned long long seed,
             unsigned stream_id, cudaStream_t st) {
    AttnParams p;
    int rc = fill_params(p, qkv, mask_bias, const_cast<void*>(ctx), const_cast<float*>(lse), dctx, dqkv, drow, B, S,
                         A, H, dropout_p, seed, stream_id);
    if (rc) return rc;
    static bool configured = false;
    if (!configured) {
        VB_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 7 * kTileBytes));
        VB_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * kTileBytes));
        configured = true;
    }
    dim3 grid((S + kBlk - 1) / kBlk, A, B);
    {
        ProfScope ps(st, PROF_ATTN, 8.0 * B * A * S * S * kHd, 2);  // algorithmic: 2x forward
        attn_bwd_dq_kernel<<<grid, 128, 7 * kTileBytes, st>>>(p);
        attn_bwd_dkv_kernel<<<grid, 128, 6 * kTileBytes, st>>>(p);
    }
    VB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace vb

extern "C" {
int vb_attention_fwd(const void* qkv, const float* mask_bias, void* ctx, float* lse, int32_t batch, int32_t seq,
                     int32_t heads, int32_t hidden, float dropout_p, uint64_t dropout_seed, uint32_t dropout_stream,
                     void* stream) {
    return vb::attn_fwd(qkv, mask_bias, ctx, lse, batch, seq, heads, hidden, dropout_p, dropout_seed, dropout_stream,
                        static_cast<cudaStream_t>(stream));
}
int vb_attention_bwd(const void* qkv, const float* mask_bias, const void* ctx, const float* lse, const void* dctx,
                     void* dqkv, float* drow, int32_t batch, int32_t seq, int32_t heads, int32_t hidden,
                     float dropout_p, uint64_t dropout_seed, uint32_t dropout_stream, void* stream) {
    return vb::attn_bwd(qkv, mask_bias, ctx, lse, dctx, dqkv, drow, batch, seq, heads, hidden, dropout_p, dropout_seed,
                        dropout_stream, static_cast<cudaStream_t>(stream));
}
}
