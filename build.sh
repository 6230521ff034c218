#!/bin/bash
# Build libvbert_b200.so for sm_100a (in-tree; the .so is git-ignored but travels with gpurun).
set -e
cd "$(dirname "$0")/visualbert_b200/csrc"
mkdir -p ../lib
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared \
     -o ../lib/libvbert_b200.so vb_gemm.cu vb_layernorm.cu vb_attention.cu vb_attention_head.cu vb_attention_tc.cu vb_embed.cu vb_heads.cu vb_optim.cu vb_api.cu "$@"
