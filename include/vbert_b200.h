/*
 * vbert_b200.h — C ABI of libvbert_b200.so: the VisualBERT encoder hot path as sm_100a kernels.
 *
 * Drop-in boundary: every entry point takes plain device pointers, sizes and a
 * cudaStream_t (passed as void*); no torch types, no allocation of persistent state, re-entrant.
 * All functions return 0 on success, non-zero on error; vb_last_error() returns the message of the
 * last failure on the calling thread. The library never throws and never calls exit().
 *
 * Each entry point names the reference code it replaces
 * (paths under uclanlp/visualbert: visualbert/pytorch_pretrained_bert/modeling.py = "M.py").
 *
 * Layout conventions: activations are row-major [rows = batch*seq, features], bf16 (2 bytes);
 * parameters handed to the library are bf16 copies ("compute weights") of the fp32 master
 * parameters in nn.Linear layout [out, in]; statistics, biases, LayerNorm affine and all
 * parameter gradients are fp32.
 */
#ifndef VBERT_B200_H
#define VBERT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VB_ABI_VERSION 1

/* ---- library ---------------------------------------------------------------------------- */
int vb_abi_version(void);
const char* vb_last_error(void);
/* number of kernel launches issued by this library on the calling process since load */
int64_t vb_launch_count(void);

/* ---- GEMM core (tcgen05.mma + TMA + TMEM) ------------------------------------------------ */
/* epilogue selectors */
#define VB_EPI_NONE 0
#define VB_EPI_GELU 1  /* D = u (pre-activation), aux_out = gelu(u)      — M.py:56-61, 302-305 */
#define VB_EPI_DGELU 2 /* D = acc * gelu'(aux_in)                        — backward of M.py:304 */

typedef struct {
    /* D[M,N] = epilogue( sum_k A(m,k) * B(n,k) )
     * a_mn_major = 0: A stored [M,K] row-major (K contiguous), lda = row stride (elements)
     * a_mn_major = 1: A stored [K,M] row-major (M contiguous)  — used for weight gradients
     * b_mn_major = 0: B stored [N,K] row-major (nn.Linear weight layout for y = x W^T)
     * b_mn_major = 1: B stored [K,N] row-major                  — used for input gradients */
    const void* A; int64_t lda; int32_t a_mn_major;
    const void* B; int64_t ldb; int32_t b_mn_major;
    int32_t M, N, K;
    void* D; int64_t ldd;
    int32_t d_fp32;   /* 0: D is bf16; 1: D is fp32 and the result is ACCUMULATED into D (red.add) */
    int32_t splits;   /* split-K factor (d_fp32 only; 0/1 = no split) */
    const float* bias;            /* fp32 [N] or NULL */
    const void* addend; int64_t ld_add; /* bf16 [M,N] added after bias/dropout (residual) or NULL */
    int32_t epilogue;             /* VB_EPI_* */
    const void* aux_in;           /* VB_EPI_DGELU: u, bf16 [M,N] */
    void* aux_out;                /* VB_EPI_GELU: gelu(u), bf16 [M,N] */
    int64_t ld_aux;
    /* inverted dropout on (acc + bias) before the addend — M.py:272, 317 (nn.Dropout) */
    float dropout_p; uint64_t dropout_seed; uint32_t dropout_stream;
} vb_gemm_args;

int vb_gemm(const vb_gemm_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VBERT_B200_H */
