// vb_api.cu — layer-level entry points of the C ABI: one call launches every kernel of a BertLayer
// forward (8 launches with attention dropout) or backward (13 launches), or of the visual+text embedding block.
//
// Kernel sequence of vb_layer_fwd (reference modeling.py:331-341):
//   1 GEMM   qkv  = x Wqkv^T + b                       (M.py:232-234, three Linears fused into N = 3H)
//   2 ATTN   ctx  = softmax(QK^T/8 + mask) V           (M.py:241-256)
//   3 GEMM   pre1 = dropout(ctx Wo^T + b) + x          (M.py:271-273, bias/dropout/residual in the epilogue)
//   4 LN     x1   = LayerNorm(pre1)                    (M.py:273)
//   5 GEMM   u, g = x1 W1^T + b, gelu(u)               (M.py:303-304, GELU in the epilogue)
//   6 GEMM   pre2 = dropout(g W2^T + b) + x1           (M.py:316-318)
//   7 LN     out  = LayerNorm(pre2)
// vb_layer_bwd is the exact adjoint (autograd of the above), weight gradients accumulated in fp32.
#include <string.h>

#include "vb_internal.h"

namespace vb {

constexpr float kLnEps = 1e-12f;
constexpr unsigned kEmbedDropStream = 0xE0000001u;
enum { kSiteAttnProbs = 0, kSiteAttnOut = 1, kSiteFfnOut = 2 };
static inline unsigned drop_stream(unsigned layer, unsigned site) { return layer * 8u + site; }

#define VB_TRY(expr)            \
    do {                        \
        int _rc = (expr);       \
        if (_rc) return _rc;    \
    } while (0)

// y[M,N] (bf16) = epi(A[M,K] W[N,K]^T)
static vb_gemm_args fwd_args(const void* A, const void* W, void* D, int M, int N, int K) {
    vb_gemm_args a;
    memset(&a, 0, sizeof(a));
    a.A = A; a.lda = K; a.B = W; a.ldb = K; a.M = M; a.N = N; a.K = K; a.D = D; a.ldd = N;
    return a;
}
// dX[M,K] (bf16) = dY[M,N] W[N,K]
static vb_gemm_args dgrad_args(const void* dY, const void* W, void* dX, int M, int N, int K) {
    vb_gemm_args a;
    memset(&a, 0, sizeof(a));
    a.A = dY; a.lda = N; a.B = W; a.ldb = K; a.b_mn_major = 1; a.M = M; a.N = K; a.K = N; a.D = dX; a.ldd = K;
    return a;
}
// dW[N,K] (fp32, +=) = dY[M,N]^T X[M,K]
static vb_gemm_args wgrad_args(const void* dY, const void* X, float* dW, int M, int N, int K) {
    vb_gemm_args a;
    memset(&a, 0, sizeof(a));
    a.A = dY; a.lda = N; a.a_mn_major = 1; a.B = X; a.ldb = K; a.b_mn_major = 1;
    a.M = N; a.N = K; a.K = M; a.D = dW; a.ldd = K; a.d_fp32 = 1;
    const int tiles = ((N + 127) / 128) * ((K + 255) / 256);
    int splits = (2 * num_sms()) / tiles;
    a.splits = splits < 1 ? 1 : splits;
    return a;
}

static int check_layer(const vb_layer_desc* d) {
    VB_REQUIRE(d != nullptr, "layer: null descriptor");
    VB_REQUIRE(d->batch > 0 && d->seq > 0, "layer: empty batch");
    VB_REQUIRE(d->hidden == d->heads * 64, "layer: hidden (%d) must equal heads (%d) * 64", d->hidden, d->heads);
    VB_REQUIRE(d->hidden % 16 == 0 && d->inter % 16 == 0, "layer: hidden/intermediate must be multiples of 16");
    VB_REQUIRE(d->w_qkv && d->w_attn_out && d->w_inter && d->w_out && d->mask_bias, "layer: null weight pointer");
    return 0;
}

int layer_fwd(const vb_layer_desc* d, const void* x_in, void* x_out, const vb_layer_acts* s, cudaStream_t st) {
    VB_TRY(check_layer(d));
    VB_REQUIRE(x_in && x_out && s, "layer_fwd: null pointer");
    const int M = d->batch * d->seq, H = d->hidden, I = d->inter;
    vb_gemm_args a = fwd_args(x_in, d->w_qkv, s->qkv, M, 3 * H, H);
    a.bias = d->b_qkv;
    // the attention-dropout bits depend on (seed, layer) only: they are drawn on a side stream UNDER the QKV GEMM (attn_mask_async)
    static cudaEvent_t before_qkv[kMaxDevices] = {nullptr};
    const int dev = current_device();
    const bool want_mask = d->attn_dropout > 0.f && s->keep_mask != nullptr;
    if (want_mask) {
        if (before_qkv[dev] == nullptr) VB_CHECK_CUDA(cudaEventCreateWithFlags(&before_qkv[dev], cudaEventDisableTiming));
        VB_CHECK_CUDA(cudaEventRecord(before_qkv[dev], st));
    }
    VB_TRY(gemm(a, st));
    int mask_ready = 0;
    if (want_mask) {
        mask_ready = attn_mask_async(s->keep_mask, d->batch, d->seq, d->heads, H, d->attn_dropout, d->seed,
                                     drop_stream(d->layer_index, kSiteAttnProbs), before_qkv[dev], st);
        if (mask_ready < 0) return 2;
    }
    VB_TRY(attn_fwd(s->qkv, d->mask_bias, s->ctx, s->lse, s->keep_mask, d->batch, d->seq, d->heads, H, d->attn_dropout, d->seed,
                    drop_stream(d->layer_index, kSiteAttnProbs), st, mask_ready == 1));
    a = fwd_args(s->ctx, d->w_attn_out, s->pre1, M, H, H);
    a.bias = d->b_attn_out; a.addend = x_in; a.ld_add = H;
    a.dropout_p = d->hidden_dropout; a.dropout_seed = d->seed; a.dropout_stream = drop_stream(d->layer_index, kSiteAttnOut);
    VB_TRY(gemm(a, st));
    VB_TRY(ln_fwd(s->pre1, H, d->ln1_gamma, d->ln1_beta, s->x1, H, s->mean1, s->rstd1, M, H, kLnEps, st));
    a = fwd_args(s->x1, d->w_inter, s->u, M, I, H);
    a.bias = d->b_inter; a.epilogue = VB_EPI_GELU; a.aux_out = s->g; a.ld_aux = I;
    a.gp_tiled = gemm_gp_tiled_ok(M, I) ? 1 : 0;   // acts.u is private to the library: tile-native whenever the shape allows
    VB_TRY(gemm(a, st));
    a = fwd_args(s->g, d->w_out, s->pre2, M, H, I);
    a.bias = d->b_out; a.addend = s->x1; a.ld_add = H;
    a.dropout_p = d->hidden_dropout; a.dropout_seed = d->seed; a.dropout_stream = drop_stream(d->layer_index, kSiteFfnOut);
    VB_TRY(gemm(a, st));
    VB_TRY(ln_fwd(s->pre2, H, d->ln2_gamma, d->ln2_beta, x_out, H, s->mean2, s->rstd2, M, H, kLnEps, st));
    return 0;
}

int layer_bwd(const vb_layer_desc* d, const void* x_in, const vb_layer_acts* s, const void* dy, void* dx,
              const vb_layer_grads* g, const vb_layer_scratch* w, cudaStream_t st) {
    VB_TRY(check_layer(d));
    VB_REQUIRE(x_in && s && dy && dx && g && w, "layer_bwd: null pointer");
    const int M = d->batch * d->seq, H = d->hidden, I = d->inter;
    const bool hd = d->hidden_dropout > 0.f;
    VB_REQUIRE(!hd || w->d_pre_drop, "layer_bwd: d_pre_drop scratch required when hidden_dropout > 0");
    void* dpm = hd ? w->d_pre_drop : w->d_pre;  // gradient entering the Linear in front of each LayerNorm

    // ---- BertOutput: LN2, output.dense ----
    VB_TRY(ln_bwd(dy, s->pre2, s->mean2, s->rstd2, d->ln2_gamma, w->d_pre, hd ? w->d_pre_drop : nullptr, g->dln2_gamma,
                  g->dln2_beta, g->db_out, M, H, d->hidden_dropout, d->seed, drop_stream(d->layer_index, kSiteFfnOut),
                  0.f, 0, st));
    VB_TRY(gemm(wgrad_args(dpm, s->g, g->dw_out, M, H, I), st));
    vb_gemm_args a = dgrad_args(dpm, d->w_out, w->d_big, M, H, I);  // d_g, then * gelu'(u) -> d_u
    a.epilogue = VB_EPI_DGELU; a.aux_in = s->u; a.ld_aux = I;
    a.gp_tiled = gemm_gp_tiled_ok(M, I) ? 1 : 0;
    VB_TRY(gemm(a, st));
    // ---- BertIntermediate ----
    VB_TRY(colsum(w->d_big, I, g->db_inter, M, I, st));
    VB_TRY(gemm(wgrad_args(w->d_big, s->x1, g->dw_inter, M, I, H), st));
    a = dgrad_args(w->d_big, d->w_inter, w->d_x1, M, I, H);
    a.addend = w->d_pre; a.ld_add = H;  // + residual branch of BertOutput
    VB_TRY(gemm(a, st));
    // ---- BertSelfOutput: LN1, attention.output.dense ----
    VB_TRY(ln_bwd(w->d_x1, s->pre1, s->mean1, s->rstd1, d->ln1_gamma, w->d_pre, hd ? w->d_pre_drop : nullptr,
                  g->dln1_gamma, g->dln1_beta, g->db_attn_out, M, H, d->hidden_dropout, d->seed,
                  drop_stream(d->layer_index, kSiteAttnOut), 0.f, 0, st));
    VB_TRY(gemm(wgrad_args(dpm, s->ctx, g->dw_attn_out, M, H, H), st));
    a = dgrad_args(dpm, d->w_attn_out, w->d_ctx, M, H, H);
    // D = rowsum(dO * O) of the attention backward falls out of this GEMM's epilogue (a thread holds two whole heads of a row)
    const bool fused_delta = gemm_delta_ok(M, H) && w->drow != nullptr &&
                             attn_bwd_takes_delta(s->qkv, w->d_ctx, w->d_big, d->batch, d->seq, d->heads, H);
    if (fused_delta) { a.delta_ctx = s->ctx; a.delta_out = w->drow; a.delta_seq = d->seq; }
    VB_TRY(gemm(a, st));
    // ---- BertSelfAttention ----
    VB_TRY(attn_bwd(s->qkv, d->mask_bias, s->ctx, s->lse, s->keep_mask, w->d_ctx, w->d_big, w->drow, d->batch, d->seq, d->heads, H,
                    d->attn_dropout, d->seed, drop_stream(d->layer_index, kSiteAttnProbs), st, fused_delta));
    VB_TRY(colsum(w->d_big, 3 * H, g->db_qkv, M, 3 * H, st));
    VB_TRY(gemm(wgrad_args(w->d_big, x_in, g->dw_qkv, M, 3 * H, H), st));
    a = dgrad_args(w->d_big, d->w_qkv, dx, M, 3 * H, H);
    a.addend = w->d_pre; a.ld_add = H;  // + residual branch of BertSelfOutput
    VB_TRY(gemm(a, st));
    return 0;
}

// ---- whole-encoder entry points: one arena, one call (see include/vbert_b200.h) ----
static long long align256(long long x) { return (x + 255) / 256 * 256; }

long long encoder_arena_layout(int B, int S, int H, int A, int I, int attn_drop, long long* off) {
    const long long M = static_cast<long long>(B) * S;
    const long long sizes[VB_ENCODER_ARENA_BUFFERS] = {
        M * 3 * H * 2, M * H * 2, static_cast<long long>(B) * A * S * 4, M * H * 2, M * 4, M * 4, M * H * 2, M * I * 2, M * I * 2,
        M * H * 2, M * 4, M * 4, attn_drop ? attn_keep_bytes(B, S, A) : 0, M * H * 2};
    long long o = 0;
    for (int i = 0; i < VB_ENCODER_ARENA_BUFFERS; ++i) {
        if (off) off[i] = o;
        o += align256(sizes[i]);
    }
    return o;
}

static void arena_acts(const vb_layer_desc* d, void* arena, int l, vb_layer_acts* a, void** y) {
    long long off[VB_ENCODER_ARENA_BUFFERS];
    const long long stride = encoder_arena_layout(d->batch, d->seq, d->hidden, d->heads, d->inter, d->attn_dropout > 0.f, off);
    char* base = static_cast<char*>(arena) + l * stride;
    a->qkv = base + off[0]; a->ctx = base + off[1]; a->lse = reinterpret_cast<float*>(base + off[2]);
    a->pre1 = base + off[3]; a->mean1 = reinterpret_cast<float*>(base + off[4]); a->rstd1 = reinterpret_cast<float*>(base + off[5]);
    a->x1 = base + off[6]; a->u = base + off[7]; a->g = base + off[8]; a->pre2 = base + off[9];
    a->mean2 = reinterpret_cast<float*>(base + off[10]); a->rstd2 = reinterpret_cast<float*>(base + off[11]);
    a->keep_mask = d->attn_dropout > 0.f ? base + off[12] : nullptr;
    *y = base + off[13];
}

int encoder_fwd(const vb_layer_desc* descs, int n, const void* x_in, void* arena, cudaStream_t st) {
    VB_REQUIRE(descs && n > 0 && x_in && arena, "encoder_fwd: null pointer / no layers");
    const void* x = x_in;
    for (int l = 0; l < n; ++l) {
        VB_REQUIRE(descs[l].batch == descs[0].batch && descs[l].seq == descs[0].seq && descs[l].hidden == descs[0].hidden &&
                   descs[l].heads == descs[0].heads && descs[l].inter == descs[0].inter &&
                   (descs[l].attn_dropout > 0.f) == (descs[0].attn_dropout > 0.f), "encoder_fwd: layers differ in shape");
        vb_layer_acts a;
        void* y;
        arena_acts(&descs[l], arena, l, &a, &y);
        VB_TRY(layer_fwd(&descs[l], x, y, &a, st));
        x = y;
    }
    return 0;
}

int encoder_bwd(const vb_layer_desc* descs, int n, const void* x_in, void* arena, const void* dy, void* dx,
                const vb_layer_grads* grads, const vb_layer_scratch* w, cudaStream_t st) {
    VB_REQUIRE(descs && n > 0 && x_in && arena && dy && dx && grads && w, "encoder_bwd: null pointer / no layers");
    const void* g_in = dy;
    for (int l = n - 1; l >= 0; --l) {
        vb_layer_acts a;
        void* y;
        arena_acts(&descs[l], arena, l, &a, &y);
        const void* xl = x_in;
        if (l > 0) {
            vb_layer_acts ap;
            void* yp;
            arena_acts(&descs[l - 1], arena, l - 1, &ap, &yp);
            xl = yp;
        }
        // the gradient buffer ping-pongs inside `dx` (vb_layer_bwd allows dx to alias dy)
        VB_TRY(layer_bwd(&descs[l], xl, &a, g_in, dx, &grads[l], w, st));
        g_in = dx;
    }
    return 0;
}

static int check_embed(const vb_embed_desc* d) {
    VB_REQUIRE(d != nullptr, "embed: null descriptor");
    VB_REQUIRE(d->batch > 0 && d->text_len > 0 && d->num_regions >= 0, "embed: bad shape");
    VB_REQUIRE(d->hidden % 16 == 0, "embed: hidden must be a multiple of 16");
    VB_REQUIRE(d->num_regions == 0 || (d->visual_dim % 8 == 0 && d->visual_feats && d->w_proj && d->visual_type),
               "embed: visual inputs missing or visual_dim not a multiple of 8");
    return 0;
}

int embed_fwd_api(const vb_embed_desc* d, void* y, const vb_embed_acts* s, cudaStream_t st) {
    VB_TRY(check_embed(d));
    VB_REQUIRE(y && s && s->pre && s->mean && s->rstd, "embed_fwd: null pointer");
    const int BV = d->batch * d->num_regions;
    if (BV > 0) {
        vb_gemm_args a = fwd_args(d->visual_feats, d->w_proj, s->vis_proj, BV, d->hidden, d->visual_dim);
        a.bias = d->b_proj;
        if (d->visual_addend != nullptr) {  // aligned position embeddings ride the projection GEMM's residual input
            a.addend = d->visual_addend;
            a.ld_add = d->hidden;
        }
        VB_TRY(gemm(a, st));
    }
    EmbedParams p;
    memset(&p, 0, sizeof(p));
    p.ids = reinterpret_cast<const long long*>(d->input_ids);
    p.tt = reinterpret_cast<const long long*>(d->token_type_ids);
    p.vt = reinterpret_cast<const long long*>(d->visual_type);
    p.vis_proj = static_cast<const bf16*>(s->vis_proj);
    p.word = d->word; p.pos = d->pos; p.type = d->type; p.pos_vis = d->pos_vis; p.type_vis = d->type_vis;
    p.gamma = d->gamma; p.beta = d->beta;
    p.pre = static_cast<bf16*>(s->pre); p.y = static_cast<bf16*>(y); p.mean = s->mean; p.rstd = s->rstd;
    p.B = d->batch; p.T = d->text_len; p.V = d->num_regions; p.H = d->hidden;
    p.vocab = d->vocab; p.max_pos = d->max_pos; p.n_types = d->n_types;
    p.eps = d->eps;
    if (d->dropout > 0.f) {
        const DropQ q = dropout_quantise(d->dropout);
        p.drop_scale = q.scale;
        p.drop_thresh16 = q.thr8;
        p.drop_seed = d->seed;
        p.drop_stream = kEmbedDropStream;
    }
    return embed_fwd(p, st);
}

int embed_bwd_api(const vb_embed_desc* d, const vb_embed_acts* s, const void* dy, const vb_embed_grads* g,
                  cudaStream_t st) {
    VB_TRY(check_embed(d));
    VB_REQUIRE(s && dy && g && g->d_pre, "embed_bwd: null pointer");
    const int M = d->batch * (d->text_len + d->num_regions), H = d->hidden, BV = d->batch * d->num_regions;
    VB_TRY(ln_bwd(dy, s->pre, s->mean, s->rstd, d->gamma, g->d_pre, nullptr, g->dgamma, g->dbeta, nullptr, M, H, 0.f,
                  d->seed, 0, d->dropout, kEmbedDropStream, st));
    EmbedBwdParams p;
    memset(&p, 0, sizeof(p));
    p.de = static_cast<const bf16*>(g->d_pre);
    p.ids = reinterpret_cast<const long long*>(d->input_ids);
    p.tt = reinterpret_cast<const long long*>(d->token_type_ids);
    p.vt = reinterpret_cast<const long long*>(d->visual_type);
    p.dword = g->dword; p.dpos = g->dpos; p.dtype = g->dtype; p.dpos_vis = g->dpos_vis; p.dtype_vis = g->dtype_vis;
    p.dvis = static_cast<bf16*>(g->d_vis);
    p.B = d->batch; p.T = d->text_len; p.V = d->num_regions; p.H = H;
    p.vocab = d->vocab; p.max_pos = d->max_pos; p.n_types = d->n_types;
    VB_TRY(embed_bwd(p, st));
    if (BV > 0) {
        VB_TRY(colsum(g->d_vis, H, g->db_proj, BV, H, st));
        VB_TRY(gemm(wgrad_args(g->d_vis, d->visual_feats, g->dw_proj, BV, H, d->visual_dim), st));
        if (g->d_feats) VB_TRY(gemm(dgrad_args(g->d_vis, d->w_proj, g->d_feats, BV, H, d->visual_dim), st));
    }
    return 0;
}

}  // namespace vb

extern "C" {
int vb_layer_fwd(const vb_layer_desc* d, const void* x_in, void* x_out, const vb_layer_acts* acts, void* stream) {
    return vb::layer_fwd(d, x_in, x_out, acts, static_cast<cudaStream_t>(stream));
}
int vb_layer_bwd(const vb_layer_desc* d, const void* x_in, const vb_layer_acts* acts, const void* dy, void* dx,
                 const vb_layer_grads* grads, const vb_layer_scratch* scratch, void* stream) {
    return vb::layer_bwd(d, x_in, acts, dy, dx, grads, scratch, static_cast<cudaStream_t>(stream));
}
int64_t vb_encoder_arena_layout(int32_t batch, int32_t seq, int32_t hidden, int32_t heads, int32_t inter, int32_t attn_dropout_on,
                                int64_t* offsets) {
    long long off[VB_ENCODER_ARENA_BUFFERS];
    const long long stride = vb::encoder_arena_layout(batch, seq, hidden, heads, inter, attn_dropout_on, off);
    if (offsets) for (int i = 0; i < VB_ENCODER_ARENA_BUFFERS; ++i) offsets[i] = off[i];
    return stride;
}
int vb_encoder_fwd(const vb_layer_desc* descs, int32_t n_layers, const void* x_in, void* arena, void* stream) {
    return vb::encoder_fwd(descs, n_layers, x_in, arena, static_cast<cudaStream_t>(stream));
}
int vb_encoder_bwd(const vb_layer_desc* descs, int32_t n_layers, const void* x_in, void* arena, const void* dy, void* dx,
                   const vb_layer_grads* grads, const vb_layer_scratch* scratch, void* stream) {
    return vb::encoder_bwd(descs, n_layers, x_in, arena, dy, dx, grads, scratch, static_cast<cudaStream_t>(stream));
}
int vb_embed_fwd(const vb_embed_desc* d, void* y, const vb_embed_acts* acts, void* stream) {
    return vb::embed_fwd_api(d, y, acts, static_cast<cudaStream_t>(stream));
}
int vb_embed_bwd(const vb_embed_desc* d, const vb_embed_acts* acts, const void* dy, const vb_embed_grads* g, void* stream) {
    return vb::embed_bwd_api(d, acts, dy, g, static_cast<cudaStream_t>(stream));
}
int vb_mask_bias(const int64_t* input_mask, const int64_t* image_mask, float* out, int32_t batch, int32_t text_len,
                 int32_t num_regions, void* stream) {
    return vb::mask_bias(reinterpret_cast<const long long*>(input_mask), reinterpret_cast<const long long*>(image_mask),
                         out, batch, text_len, num_regions, static_cast<cudaStream_t>(stream));
}
int vb_cast_f32_to_bf16(const float* src, void* dst, int64_t n, void* stream) {
    return vb::cast_f32_bf16(src, dst, n, static_cast<cudaStream_t>(stream));
}
int vb_cast_bf16_to_f32(const void* src, float* dst, int64_t n, void* stream) {
    return vb::cast_bf16_f32(src, dst, n, static_cast<cudaStream_t>(stream));
}
int vb_colsum_bf16(const void* x, int64_t ld, float* out, int32_t rows, int32_t cols, void* stream) {
    return vb::colsum(x, ld, out, rows, cols, static_cast<cudaStream_t>(stream));
}
}
