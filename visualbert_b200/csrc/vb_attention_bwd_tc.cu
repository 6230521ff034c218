// vb_attention_bwd_tc.cu — attention backward on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), seq <= 192.
//
// Adjoint of reference modeling.py:241-256 (QK^T / sqrt(d) + mask -> softmax -> dropout -> P V). One persistent CTA per SM
// walks over (batch, head) items. All five contractions run on tcgen05.mma with TMEM accumulators; the element-wise
// work runs TRANSPOSED — TMEM lane = key, column = query — so nothing needs a cross-thread reduction (the row
// statistics lse[q] and D[q] = sum_d dO O are per COLUMN vectors staged in shared memory):
//
//   per key tile kt (128 lanes; the second tile of seq = 128 + r sits at a rotating lane offset, see vb_attention_tc.cu):
//     S^T  = K_kt Q^T        M=128 N=npq K=64    A = K tile (smem), B = Q (smem)               -> TMEM
//     dP^T = V_kt dO^T       M=128 N=npq K=64    A = V tile,        B = dO                     -> TMEM
//     element-wise, 8 warps (two warp-groups split the query columns), one thread per (key, half row):
//         p   = exp2(s * scale*log2e + mask[key] - lse[q])
//         Pd  = keep ? p / (1-pd) : 0              -> bf16 -> TMEM, in place over the consumed S^T columns
//         dS  = p * (keep ? dP / (1-pd) : 0 - D[q]) -> bf16 -> shared memory [key][q], 128B-swizzled
//     dV_kt = Pd^T dO        M=128 N=64 K=npq    A = Pd^T FROM TMEM,  B = dO (MN-major)        -> TMEM -> global
//     dK_kt = dS^T Q         M=128 N=64 K=npq    A = dS^T (smem, K-major), B = Q (MN-major)    -> TMEM -> global (* scale)
//   per item, once both key tiles have left their dS^T in shared memory:
//     dQ    = dS K           M=128 (queries) N=64 K=keys   A = the SAME dS^T bytes read MN-major, B = K (MN-major)
//
//   warp 0  TMA producer (Q, dO, K double-buffered per item; V single-buffered: it is dead after the two dP^T MMAs)
//   warp 1  MMA issuer (one thread)      warp 2  TMEM allocator      warp 3  stages lse * log2e and D of the item
//   warps 4-11  element-wise + output drains (dV / dQ tile 0 by warp-group 0, dK / dQ tile 1 by warp-group 1)
#include "vb_attention.cuh"

namespace vb {

namespace {

constexpr int kRows = 128;            // UMMA M (keys per tile / queries per dQ tile)
constexpr int kWgT = 128;             // threads per warp-group
constexpr int kThreadsBw = 128 + 2 * kWgT;
constexpr int kMaxNpq = 192;
constexpr uint32_t kTmSt = 0, kTmDpt = 192, kTmAcc = 384;   // TMEM columns: S^T | dP^T | dV,dK (or dQ tile 0, 1)

__device__ __forceinline__ void tma_load_3d(uint32_t smem_dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= 1ull << 46;
    d |= 2ull << 61;
    return d;
}
__device__ __forceinline__ uint32_t idesc_bw(int m, int n, bool a_mn, bool b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(a_mn) << 15) | (static_cast<uint32_t>(b_mn) << 16) |
           (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

struct BwLayout {   // byte offsets from the 1 KB-aligned base
    int qbytes;      // npq * 128 (Q and dO tiles)
    int k1bytes;     // r2pad * 128 (compact window of the second key tile)
    int stage_bytes; // Q | dO | K0
    int nstage;
    int k1_off;      // [nstage] K windows, then the V window, then dS^T of the second tile (natoms x k1bytes)
    int v1_off, ds1_off;
    int ds0_off;     // natoms x 16 KB
    int v0_off;
    int vec_off;     // fp32 [nstage][2][kMaxNpq]: lse * log2e | D
    int bar_off, tmem_ptr_off, total;
};
__host__ __device__ inline BwLayout bw_layout(int npq, int r2pad, int nstage) {
    BwLayout L;
    const int natoms = (npq + 63) / 64;
    L.qbytes = npq * 128;
    L.k1bytes = r2pad * 128;
    L.stage_bytes = 2 * L.qbytes + kRows * 128;
    L.nstage = nstage;
    L.k1_off = nstage * L.stage_bytes;
    L.v1_off = L.k1_off + nstage * L.k1bytes;
    L.ds1_off = L.v1_off + L.k1bytes;
    L.ds0_off = L.ds1_off + natoms * L.k1bytes;
    L.v0_off = L.ds0_off + natoms * kRows * 128;
    L.vec_off = L.v0_off + kRows * 128;
    L.bar_off = L.vec_off + nstage * 2 * kMaxNpq * 4;
    L.tmem_ptr_off = L.bar_off + 16 * 8;
    L.total = L.tmem_ptr_off + 16 + 1024;
    return L;
}

struct BwParams {
    AttnParams a;
    int npq;     // queries padded to 16 (MMA N of S^T / dP^T, K of dV / dK)
    int n0;      // query columns of warp-group 0 (multiple of 16); warp-group 1 owns [n0, npq)
    int nkt;     // key tiles (1 or 2)
    int r2;      // keys of the second tile, r2pad = r2 rounded up to 16
    int r2pad;
    int nkb;     // ceil(S / 64)
    int nstage;
};

__device__ __forceinline__ int kt1_offset(int li, int r2pad) {
    const int lim = kRows - r2pad;   // multiple of 16
    const int o = (li & 3) * 32;
    return o < lim ? o : lim;
}

__global__ void __launch_bounds__(kThreadsBw, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmDO,
                   const __grid_constant__ CUtensorMap tmK0, const __grid_constant__ CUtensorMap tmK1, const BwParams bp) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - raw);
    const AttnParams& p = bp.a;
    const BwLayout L = bw_layout(bp.npq, bp.r2pad, bp.nstage);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = p.S, npq = bp.npq, nkt = bp.nkt, nstage = bp.nstage;
    const int natoms = (npq + 63) / 64;

    auto q_tile = [&](int s) { return base + s * L.stage_bytes; };
    auto do_tile = [&](int s) { return base + s * L.stage_bytes + L.qbytes; };
    auto k0_tile = [&](int s) { return base + s * L.stage_bytes + 2 * L.qbytes; };
    auto k1_win = [&](int s) { return base + L.k1_off + s * L.k1bytes; };
    const uint32_t v0_tile = base + L.v0_off, v1_win = base + L.v1_off;
    const uint32_t ds0 = base + L.ds0_off, ds1 = base + L.ds1_off;
    auto bar = [&](int i) { return base + L.bar_off + 8 * i; };
    enum { IN_FULL0 = 0, IN_EMPTY0 = 2, V_FULL = 4, V_EMPTY = 5, SD_FULL = 6, EW_DONE = 7, ACC_FULL = 8, ACC_EMPTY = 9 };
    volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(smem + L.tmem_ptr_off);
    float* svec_all = reinterpret_cast<float*>(smem + L.vec_off);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmDO);
        tma_prefetch_desc(&tmK0);
        tma_prefetch_desc(&tmK1);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(bar(IN_FULL0 + s), 2);   // TMA producer (expect_tx) + the vector stager
            mbar_init(bar(IN_EMPTY0 + s), 1);
        }
        mbar_init(bar(V_FULL), 1);
        mbar_init(bar(V_EMPTY), 1);
        mbar_init(bar(SD_FULL), 1);
        mbar_init(bar(EW_DONE), 2 * kWgT);
        mbar_init(bar(ACC_FULL), 1);
        mbar_init(bar(ACC_EMPTY), 2 * kWgT);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(base + L.tmem_ptr_off, 512);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_trigger();
    pdl_wait();

    const int total = p.B * p.A;
    const int n_local = (total - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

    // register split: 128 x 88 (control warpgroup) + 256 x 208 (element-wise warpgroups) = the 384 x 168 of the launch
    if (warp < 4) {
        reg_dec<88>();
        const int n_units = n_local * nkt;   // (item, key tile) units of this CTA, in order
        if (warp == 0) {
            if (lane == 0) {
                // ---------------- TMA producer ----------------
                const uint32_t in_tx = static_cast<uint32_t>(2 * npq * 128 + kRows * 128 + (nkt == 2 ? bp.r2pad * 128 : 0));
                const uint32_t v_tx = static_cast<uint32_t>(kRows * 128 + (nkt == 2 ? bp.r2pad * 128 : 0));
                for (int li = 0; li < n_local; ++li) {
                    const int item = blockIdx.x + li * gridDim.x;
                    const int b = item / p.A, h = item % p.A;
                    const int s = li % nstage;
                    mbar_wait(bar(IN_EMPTY0 + s), ((li / nstage) & 1) ^ 1u);
                    mbar_arrive_expect_tx(bar(IN_FULL0 + s), in_tx);
                    tma_load_3d(q_tile(s), &tmQ, bar(IN_FULL0 + s), h * kHd, 0, b);
                    tma_load_3d(do_tile(s), &tmDO, bar(IN_FULL0 + s), h * kHd, 0, b);
                    tma_load_3d(k0_tile(s), &tmK0, bar(IN_FULL0 + s), p.H + h * kHd, 0, b);
                    if (nkt == 2) tma_load_3d(k1_win(s), &tmK1, bar(IN_FULL0 + s), p.H + h * kHd, kRows, b);
                    mbar_wait(bar(V_EMPTY), (li & 1) ^ 1u);
                    mbar_arrive_expect_tx(bar(V_FULL), v_tx);
                    tma_load_3d(v0_tile, &tmK0, bar(V_FULL), 2 * p.H + h * kHd, 0, b);
                    if (nkt == 2) tma_load_3d(v1_win, &tmK1, bar(V_FULL), 2 * p.H + h * kHd, kRows, b);
                }
            }
        } else if (warp == 1) {
            if (lane == 0) {
                // ---------------- MMA issuer ----------------
                // One loop over the units, ONE copy of every MMA sequence (the issuing thread's instruction stream is the
                // critical path of the pipeline: per MMA it is a 32-bit add per descriptor and the tcgen05.mma itself).
                const uint32_t idesc_sd = idesc_bw(kRows, npq, false, false);
                const uint32_t idesc_vk = idesc_bw(kRows, kHd, false, true);
                const uint32_t idesc_dq = idesc_bw(kRows, kHd, true, true);
                const uint32_t tm_st = tmem_base + kTmSt, tm_dpt = tmem_base + kTmDpt, tm_acc = tmem_base + kTmAcc;
                const int ksteps = npq / 16;
                const int k0steps = bp.n0 / 16;       // k-steps whose Pd^T chunk belongs to warp-group 0
                const int nqt = npq > kRows ? 2 : 1;
                uint32_t acc_n = 0;                     // uses of the ACC barriers so far
                for (int u = -1; u < n_units; ++u) {
                    if (u >= 0) {
                        // ---- dV_kt = Pd^T dO (A from TMEM), dK_kt = dS^T Q (A from shared memory) ----
                        const int li = u / nkt, kt = u - li * nkt, s = li % nstage;
                        const uint32_t off = kt == 0 ? 0u : static_cast<uint32_t>(kt1_offset(li, bp.r2pad) * 128);
                        mbar_wait(bar(EW_DONE), u & 1);
                        mbar_wait(bar(ACC_EMPTY), (acc_n & 1) ^ 1u);
                        tcgen05_fence_after();
                        const UmmaDesc d_do = make_umma_desc_sw128(do_tile(s), 0, 1024);
                        const UmmaDesc d_q = make_umma_desc_sw128(q_tile(s), 0, 1024);
                        // Pd^T: warp-group 0's half starts at column 0, warp-group 1's at column n0 (8 columns per k-step)
                        uint32_t a_col = tm_st;
                        for (int k = 0; k < ksteps; ++k) {
                            if (k == k0steps) a_col = tm_st + bp.n0;
                            umma_bf16_ts(tm_acc, a_col, d_do.at(k * 2048), idesc_vk, k > 0 ? 1u : 0u);
                            a_col += 8;
                        }
                        // dS^T tile: K-major, 64-query atoms (tile 1: compact window addressed `off` bytes early)
                        const uint32_t ds_base = kt == 0 ? ds0 : ds1 - off;
                        const uint32_t ds_atom = kt == 0 ? kRows * 128 : static_cast<uint32_t>(L.k1bytes);
                        const UmmaDesc d_ds = make_umma_desc_sw128(ds_base, 0, 1024);
                        for (int k = 0; k < ksteps; ++k)
                            umma_bf16(tm_acc + kHd, d_ds.at((k >> 2) * ds_atom + (k & 3) * 32), d_q.at(k * 2048), idesc_vk, k > 0 ? 1u : 0u);
                        umma_commit(bar(ACC_FULL));
                        ++acc_n;
                    }
                    const bool item_end = u >= 0 && (u + 1) % nkt == 0;
                    // after an item's last tile: the next item's first S^T / dP^T run under this item's drains when its
                    // inputs live in the other stage (nstage == 2); with one stage dQ must release the inputs first
                    for (int ph = 0; ph < 2; ++ph) {
                        const bool do_sd = item_end ? ph == (nstage == 2 ? 0 : 1) : ph == 0;
                        if (do_sd) {
                            if (u + 1 < n_units) {
                                // ---- S^T = K_kt Q^T and dP^T = V_kt dO^T of unit u + 1 ----
                                const int li = (u + 1) / nkt, kt = (u + 1) - li * nkt, s = li % nstage;
                                if (kt == 0) {
                                    mbar_wait(bar(IN_FULL0 + s), (li / nstage) & 1);
                                    mbar_wait(bar(V_FULL), li & 1);
                                    tcgen05_fence_after();
                                }
                                const uint32_t off = kt == 0 ? 0u : static_cast<uint32_t>(kt1_offset(li, bp.r2pad) * 128);
                                const UmmaDesc d_k = make_umma_desc_sw128(kt == 0 ? k0_tile(s) : k1_win(s) - off, 0, 1024);
                                const UmmaDesc d_v = make_umma_desc_sw128(kt == 0 ? v0_tile : v1_win - off, 0, 1024);
                                const UmmaDesc d_q = make_umma_desc_sw128(q_tile(s), 0, 1024);
                                const UmmaDesc d_do = make_umma_desc_sw128(do_tile(s), 0, 1024);
#pragma unroll
                                for (int k = 0; k < kHd / 16; ++k) umma_bf16(tm_st, d_k.at(k * 32), d_q.at(k * 32), idesc_sd, k > 0 ? 1u : 0u);
#pragma unroll
                                for (int k = 0; k < kHd / 16; ++k) umma_bf16(tm_dpt, d_v.at(k * 32), d_do.at(k * 32), idesc_sd, k > 0 ? 1u : 0u);
                                umma_commit(bar(SD_FULL));
                                if (kt == nkt - 1) umma_commit(bar(V_EMPTY));   // V is dead once the item's last dP^T has retired
                            }
                        } else if (item_end) {
                            // ---- dQ = dS K over all keys of the item: A = the dS^T bytes read MN-major (M = queries) ----
                            const int li = u / nkt, s = li % nstage;
                            mbar_wait(bar(ACC_EMPTY), (acc_n & 1) ^ 1u);
                            tcgen05_fence_after();
                            const UmmaDesc d_k0 = make_umma_desc_sw128(k0_tile(s), 0, 1024);
                            const UmmaDesc d_k1 = make_umma_desc_sw128(k1_win(s), 0, 1024);
                            for (int m = 0; m < nqt; ++m) {
                                const UmmaDesc a0 = make_umma_desc_sw128(ds0 + (2 * m) * (kRows * 128), kRows * 128, 1024);
                                const UmmaDesc a1 = make_umma_desc_sw128(ds1 + (2 * m) * L.k1bytes, L.k1bytes, 1024);
                                for (int j = 0; j < kRows / 16; ++j)      // keys of tile 0: 16 key rows = 2048 B per step
                                    umma_bf16(tm_acc + m * kHd, a0.at(j * 2048), d_k0.at(j * 2048), idesc_dq, j > 0 ? 1u : 0u);
                                if (nkt == 2)
                                    for (int j = 0; j < bp.r2pad / 16; ++j)   // keys of tile 1: the compact window rows
                                        umma_bf16(tm_acc + m * kHd, a1.at(j * 2048), d_k1.at(j * 2048), idesc_dq, 1u);
                            }
                            umma_commit(bar(ACC_FULL));
                            umma_commit(bar(IN_EMPTY0 + s));   // every MMA of the item has retired: Q / dO / K are free
                            ++acc_n;
                        }
                    }
                }
            }
        } else if (warp == 3) {
            // ---------------- per-item column vectors: lse * log2e (+inf on padding: p = 0) and D ----------------
            for (int li = 0; li < n_local; ++li) {
                const int item = blockIdx.x + li * gridDim.x;
                const int s = li % nstage;
                mbar_wait(bar(IN_EMPTY0 + s), ((li / nstage) & 1) ^ 1u);
                float* sv = svec_all + s * 2 * kMaxNpq;
                const float* lsep = p.lse + static_cast<long long>(item) * S;
                const float* drp = p.drow + static_cast<long long>(item) * S;
                for (int i = lane; i < npq; i += 32) {
                    sv[i] = i < S ? __ldg(lsep + i) * kLog2e : INFINITY;
                    sv[kMaxNpq + i] = i < S ? __ldg(drp + i) : 0.f;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(IN_FULL0 + s));
            }
        }
    } else {
        reg_inc<208>();
        const int n_units = n_local * nkt;
        // ---------------- element-wise + drains: thread = (key row r, query half g) ----------------
        const int g = (warp - 4) >> 2;
        const int q4 = warp & 3;
        const int r = q4 * 32 + lane;
        const uint32_t lane_sel = static_cast<uint32_t>(q4 * 32) << 16;
        const float sc2 = p.scale * kLog2e;
        const bool drop = p.drop_scale != 0.f;
        const float ds = drop ? p.drop_scale : 1.f;
        const int np64 = bp.nkb * kBlk;
        const int q0 = g ? bp.n0 : 0;                       // first query column of this warp-group
        const int np8 = (g ? npq - bp.n0 : bp.n0) / 8;      // 8-column pieces of this warp-group (even)
        const uint32_t tm_st = tmem_base + kTmSt + lane_sel + q0, tm_dpt = tmem_base + kTmDpt + lane_sel + q0;
        const uint32_t tm_acc = tmem_base + kTmAcc + lane_sel + g * kHd;
        const unsigned long long* keepT = p.keep + static_cast<long long>(p.B) * p.A * np64 * bp.nkb;
        uint32_t acc_n = 0;
        const long long ld3 = 3LL * p.H;
        for (int u = 0; u < n_units; ++u) {
            const int li = u / nkt, kt = u - li * nkt;
            const int item = blockIdx.x + li * gridDim.x;
            const int b = item / p.A, h = item % p.A;
            const int s = li % nstage;
            const float* slse = svec_all + s * 2 * kMaxNpq;
            const float* sD = slse + kMaxNpq;
            bf16* dbase = p.dqkv + static_cast<long long>(b) * S * ld3 + h * kHd;
            const int off = kt == 0 ? 0 : kt1_offset(li, bp.r2pad);
            const int wrows = kt == 0 ? kRows : bp.r2pad;            // rows of this tile that carry (possibly zero) data
            const bool inwin = r >= off && r < off + wrows;
            const bool wany = __any_sync(0xffffffffu, inwin);
            const int j = kt * kRows + r - off;                      // key index of this lane
            float bias2 = -INFINITY;
            unsigned long long kw0 = ~0ull, kw1 = ~0ull, kw2 = ~0ull;
            if (inwin && j < S) {
                bias2 = __ldg(p.mask_bias + static_cast<long long>(b) * S + j) * kLog2e;
                if (drop) {
                    const unsigned long long* kp = keepT + (static_cast<long long>(item) * np64 + j) * bp.nkb;
                    kw0 = kp[0];
                    if (bp.nkb > 1) kw1 = kp[1];
                    if (bp.nkb > 2) kw2 = kp[2];
                }
            }
            if (kt == 0) mbar_wait(bar(IN_FULL0 + s), (li / nstage) & 1);   // lse / D of this item staged and visible
            mbar_wait(bar(SD_FULL), u & 1);
            tcgen05_fence_after();
            if (wany) {
                // dS^T row of this lane: [atom = q / 64][row][128 B], 16-byte chunks XOR-swizzled with (row & 7)
                const uint32_t ds_row = kt == 0 ? ds0 + r * 128 : ds1 + (r - off) * 128;
                const uint32_t ds_atom = kt == 0 ? kRows * 128 : static_cast<uint32_t>(L.k1bytes);
                const int sw = r & 7;
                // one 8-query piece: i-th piece of this warp-group's column range
                auto ew8 = [&](const uint32_t (&vs)[8], const uint32_t (&vd)[8], int i) {
                    const int c8 = (q0 >> 3) + i;        // 8-query piece index inside the row
                    const int cg = c8 >> 1;               // its 16-query chunk
                    const float4* l4 = reinterpret_cast<const float4*>(slse + c8 * 8);
                    const float4* d4 = reinterpret_cast<const float4*>(sD + c8 * 8);
                    const unsigned long long w = cg < 4 ? kw0 : (cg < 8 ? kw1 : kw2);
                    const uint32_t bits = drop ? static_cast<uint32_t>(w >> ((c8 & 7) * 8)) : 0xffu;
                    float pd[8], dsv[8];
#pragma unroll
                    for (int jj = 0; jj < 2; ++jj) {
                        const float4 ll = l4[jj], dd = d4[jj];
                        const float lv[4] = {ll.x, ll.y, ll.z, ll.w}, dv[4] = {dd.x, dd.y, dd.z, dd.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int k = 4 * jj + e;
                            const float pr = fast_ex2(fmaf(__uint_as_float(vs[k]), sc2, bias2) - lv[e]);
                            const float t = ((bits >> k) & 1u) ? ds : 0.f;
                            pd[k] = pr * t;
                            dsv[k] = pr * fmaf(__uint_as_float(vd[k]), t, -dv[e]);
                        }
                    }
                    // Pd^T: 4 columns inside this thread's consumed S^T columns (8 bf16)
                    tmem_st_32x32b_x4(tm_st + i * 4, pack_bf16x2(pd[0], pd[1]), pack_bf16x2(pd[2], pd[3]), pack_bf16x2(pd[4], pd[5]),
                                      pack_bf16x2(pd[6], pd[7]));
                    if (inwin)
                        st_shared_v4(ds_row + (cg >> 2) * ds_atom + (((c8 & 7) ^ sw) << 4), pack_bf16x2(dsv[0], dsv[1]),
                                     pack_bf16x2(dsv[2], dsv[3]), pack_bf16x2(dsv[4], dsv[5]), pack_bf16x2(dsv[6], dsv[7]));
                };
                uint32_t sa[8], da[8], sb[8], db[8];
                if (np8 > 0) {
                    tmem_ld_32x32b_x8(tm_st, sa);
                    tmem_ld_32x32b_x8(tm_dpt, da);
                }
                for (int i = 0; i < np8; i += 2) {
                    tmem_ld_wait();
                    tmem_ld_32x32b_x8(tm_st + (i + 1) * 8, sb);      // np8 is even
                    tmem_ld_32x32b_x8(tm_dpt + (i + 1) * 8, db);
                    ew8(sa, da, i);
                    tmem_ld_wait();
                    if (i + 2 < np8) {
                        tmem_ld_32x32b_x8(tm_st + (i + 2) * 8, sa);
                        tmem_ld_32x32b_x8(tm_dpt + (i + 2) * 8, da);
                    }
                    ew8(sb, db, i + 1);
                }
                tmem_st_wait();
            }
            fence_proxy_async_smem();   // dS^T (generic-proxy stores) -> visible to the tensor core
            tcgen05_fence_before();
            mbar_arrive(bar(EW_DONE));
            // ---- drains (one copy of the code): dV_kt (warp-group 0) / dK_kt (warp-group 1), lane = key; after the item's
            //      last tile additionally dQ tile 0 (warp-group 0) / tile 1 (warp-group 1), lane = query ----
            const int ndrain = kt == nkt - 1 ? 2 : 1;
            for (int dr = 0; dr < ndrain; ++dr) {
                const int row = dr == 0 ? j : g * kRows + r;                       // key (dV / dK) or query (dQ) index
                const bool valid = (dr == 0 ? inwin : true) && row < S;
                const bool wload = __any_sync(0xffffffffu, valid);
                const float mul = (dr == 0 && g == 0) ? 1.f : p.scale;             // dV is unscaled, dK and dQ carry 1/sqrt(d)
                bf16* dst = dbase + static_cast<long long>(row) * ld3 + (dr == 0 ? (g == 0 ? 2 * p.H : p.H) : 0);
                mbar_wait(bar(ACC_FULL), acc_n & 1);
                ++acc_n;
                tcgen05_fence_after();
#pragma unroll 1
                for (int half = 0; half < 2; ++half) {   // 2 x 32 columns: short live ranges (the EW loop owns the registers)
                    uint32_t o[2][16];
                    if (wload) {
                        tmem_ld_32x32b_x16(tm_acc + half * 32, o[0]);
                        tmem_ld_32x32b_x16(tm_acc + half * 32 + 16, o[1]);
                        tmem_ld_wait();
                    }
                    if (half == 1) {
                        tcgen05_fence_before();
                        mbar_arrive(bar(ACC_EMPTY));
                    }
                    if (valid) {
#pragma unroll
                        for (int jj = 0; jj < 2; ++jj) {
                            uint32_t w[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                w[i] = pack_bf16x2(__uint_as_float(o[jj][2 * i]) * mul, __uint_as_float(o[jj][2 * i + 1]) * mul);
                            stg_v8(dst + half * 32 + jj * 16, w);
                        }
                    }
                }
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace

static bool bwd_tc_config(const AttnParams& p, BwParams& bp) {
    if (p.S < 1 || p.S > kMaxNpq || (p.H % 8) != 0) return false;
    bp.a = p;
    bp.npq = (p.S + 15) / 16 * 16;
    bp.n0 = ((bp.npq / 16 + 1) / 2) * 16;
    bp.nkt = p.S > kRows ? 2 : 1;
    bp.r2 = bp.nkt == 2 ? p.S - kRows : 0;
    bp.r2pad = (bp.r2 + 15) / 16 * 16;
    bp.nkb = (p.S + kBlk - 1) / kBlk;
    bp.nstage = bw_layout(bp.npq, bp.r2pad, 2).total <= 227 * 1024 ? 2 : 1;
    return bw_layout(bp.npq, bp.r2pad, bp.nstage).total <= 227 * 1024;
}

bool attn_bwd_tc_supported(const AttnParams& p) {
    BwParams bp;
    return bwd_tc_config(p, bp) && (reinterpret_cast<uintptr_t>(p.qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.dctx) & 15) == 0 &&
           (reinterpret_cast<uintptr_t>(p.dqkv) & 31) == 0;
}

// dQ | dK | dV of one layer call; p.drow must already hold D = rowsum(dO * O) (attn_delta_kernel)
int attn_bwd_tc(const AttnParams& p, cudaStream_t st) {
    BwParams bp;
    VB_REQUIRE(bwd_tc_config(p, bp), "attention backward (tcgen05): unsupported shape S=%d", p.S);
    const BwLayout L = bw_layout(bp.npq, bp.r2pad, bp.nstage);
    CUtensorMap tq, tdo, tk0, tk1;
    int rc = make_tmap_3d(&tq, p.qkv, p.S, p.B, 3 * p.H, bp.npq);
    if (rc) return rc;
    rc = make_tmap_3d(&tdo, p.dctx, p.S, p.B, p.H, bp.npq);
    if (rc) return rc;
    rc = make_tmap_3d(&tk0, p.qkv, p.S, p.B, 3 * p.H, kRows);
    if (rc) return rc;
    rc = make_tmap_3d(&tk1, p.qkv, p.S, p.B, 3 * p.H, bp.nkt == 2 ? bp.r2pad : 16);
    if (rc) return rc;
    static int configured[kMaxDevices] = {0};
    VB_CHECK_CUDA(ensure_dyn_smem(attn_bwd_tc_kernel, L.total, configured));
    const int total = p.B * p.A;
    const int grid = total < num_sms() ? total : num_sms();
    {
        ProfScope ps(st, PROF_ATTN_DKV, 8.0 * p.B * p.A * p.S * p.S * kHd, 1);
        VB_CHECK_CUDA(launch_pdl(attn_bwd_tc_kernel, dim3(grid), dim3(kThreadsBw), static_cast<size_t>(L.total), st, tq, tdo, tk0, tk1, bp));
    }
    VB_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace vb
