#!/bin/bash
# Build the instrumented attention kernel beside the engine's other objects (python ml-depth-pro-video_b200/build.py first).
set -e
cd "$(dirname "$0")/../.."
C=ml-depth-pro-video_b200/csrc; B=ml-depth-pro-video_b200/build
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17"
nvcc $F -DATTN_PROFILE -I $C -c $C/attention_tc.cu -o /tmp/attention_tc_prof.o
nvcc $F -I $C -c scripts/ubench/attn_prof.cu -o /tmp/attn_prof_main.o
nvcc -gencode arch=compute_100a,code=sm_100a -o scripts/ubench/attn_prof /tmp/attn_prof_main.o /tmp/attention_tc_prof.o \
  $B/gemm_tc.o $B/kernels.o $B/engine.o $B/gemm_simt.o $B/attention.o $B/ground.o -lcuda
