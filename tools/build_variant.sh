#!/bin/bash
# usage: tools/build_variant.sh NAME [-DFLAG ...]  -> kaldi-decoder_b200/lib/variants/libkd_NAME.so (160-thread lanes only)
set -e
name=$1; shift
mkdir -p kaldi-decoder_b200/lib/variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC,-fvisibility=hidden \
  -Iinclude -Ikaldi-decoder_b200/csrc -DKD_ONLY_160 "$@" kaldi-decoder_b200/csrc/kd_capi.cu \
  -o kaldi-decoder_b200/lib/variants/libkd_$name.so
echo built $name
