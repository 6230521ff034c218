// vb_internal.h — declarations shared between the translation units of libvbert_b200 (not part of the ABI).
#pragma once
#include "../../include/vbert_b200.h"
#include "vb_common.cuh"

namespace vb {

int gemm(const vb_gemm_args& a, cudaStream_t st);
bool gemm_gp_tiled_ok(int M, int N);
bool gemm_delta_ok(int M, int N);
int ln_fwd(const void* x, long long ldx, const float* gamma, const float* beta, void* y, long long ldy, float* mean,
           float* rstd, int rows, int H, float eps, cudaStream_t st);
int ln_bwd(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma, void* dx,
           void* dx_drop, float* dgamma, float* dbeta, float* dbias, int rows, int H, float dropout_p,
           unsigned long long seed, unsigned stream_id, float in_dropout_p, unsigned in_stream_id, cudaStream_t st);
int attn_fwd(const void* qkv, const float* mask_bias, void* ctx, float* lse, void* keep, int B, int S, int A, int H,
             float dropout_p, unsigned long long seed, unsigned stream_id, cudaStream_t st, bool mask_ready = false);
int attn_mask_async(void* keep, int B, int S, int A, int H, float dropout_p, unsigned long long seed, unsigned stream_id,
                    cudaEvent_t before_gemm, cudaStream_t main);
int attn_bwd(const void* qkv, const float* mask_bias, const void* ctx, const float* lse, const void* keep,
             const void* dctx, void* dqkv, float* drow, int B, int S, int A, int H, float dropout_p,
             unsigned long long seed, unsigned stream_id, cudaStream_t st, bool delta_ready = false);
// true when attn_bwd for this shape runs the kernel that takes D = rowsum(dO * O) from `drow` (so a caller may provide it)
bool attn_bwd_takes_delta(const void* qkv, const void* dctx, void* dqkv, int B, int S, int A, int H);
long long attn_keep_bytes(int B, int S, int A);
int colsum(const void* x, long long ld, float* out, int M, int N, cudaStream_t st);
int cast_f32_bf16(const float* src, void* dst, long long n, cudaStream_t st);
int cast_bf16_f32(const void* src, float* dst, long long n, cudaStream_t st);
int mask_bias(const long long* input_mask, const long long* image_mask, float* out, int B, int T, int V, cudaStream_t st);

struct EmbedParams {
    const long long* ids; const long long* tt; const long long* vt;
    const bf16* vis_proj;
    const float* word; const float* pos; const float* type; const float* pos_vis; const float* type_vis;
    const float* gamma; const float* beta;
    bf16* pre; bf16* y; float* mean; float* rstd;
    int B, T, V, H, vocab, max_pos, n_types;
    float eps;
    float drop_scale; unsigned drop_thresh16; unsigned long long drop_seed; unsigned drop_stream;
};
struct EmbedBwdParams {
    const bf16* de;
    const long long* ids; const long long* tt; const long long* vt;
    float* dword; float* dpos; float* dtype; float* dpos_vis; float* dtype_vis;
    bf16* dvis;
    int B, T, V, H, vocab, max_pos, n_types;
};
int embed_fwd(const EmbedParams& p, cudaStream_t st);
int embed_bwd(const EmbedBwdParams& p, cudaStream_t st);


}  // namespace vb
