#!/bin/bash
# dev compile of one csrc file with ptxas -v summary: tools/exp/cc.sh decode [-DMB_DEC_DEV]
cd /root/repo/deep-learning-based-sequence-models-for-music-generation_b200
f=$1; shift
nvcc "$@" -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -Xptxas=-v -c csrc/$f.cu -o /tmp/$f.o 2>&1 | grep -E "error|Compiling|spill|Used" | paste - - - | sed -E 's/.*Compiling entry function .([^ ]*). for .sm_100a.(.*)/\1 \2/' | sed -E 's/_ZN2mb[0-9]*_GLOBAL__N__[0-9a-f_]*cu_[0-9a-f]*//; s/ptxas info    : //g; s/[0-9]+ bytes stack frame, //'
