#!/bin/bash
# Build libvbert_b200.so for sm_100a (in-tree; the .so is git-ignored but travels with gpurun).
set -e
make -C "$(dirname "$0")/visualbert_b200/csrc" -j"$(nproc)" "$@"
