#!/bin/bash
# Measurement build (not part of the product): libyolo3_b200.so with the clock64 traces of the fused stem kernel
# (-DY3_STEM_TRACE: per-step time stamps of the MMA threads, two stem warps and one epilogue group of CTA 0) and of
# k_conv_tc2h (-DY3_CONV_TRACE: cycles CTA 0's MMA thread / producer / first epilogue warp spend waiting), written to
# variants/tr/.  Run on the GPU box with
#   cp variants/tr/libyolo3_b200.so object-detection-yolov3_b200/libyolo3_b200.so
#   python tools/profile_layers.py 512 512 1 1 128 2>&1 | grep -E "TRACE|CONVTRACE"
# (profiles/r2_trace_*.txt are outputs of exactly that).
set -e
cd "$(dirname "$0")/../object-detection-yolov3_b200"
make -s -j16 all
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr -Xcompiler -fPIC"
mkdir -p ../variants/tr
$NV -DY3_STEM_TRACE -c csrc/conv_stem1.cu -o ../variants/tr/conv_stem1.o
$NV -DY3_CONV_TRACE -c csrc/conv_tc2h.cu -o ../variants/tr/conv_tc2h.o
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o ../variants/tr/libyolo3_b200.so \
    $(ls build/*.o | grep -v -E "umma_probe|conv_stem1|conv_tc2h") ../variants/tr/conv_stem1.o ../variants/tr/conv_tc2h.o -ldl
echo "built variants/tr/libyolo3_b200.so"
