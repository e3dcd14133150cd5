#!/bin/bash
# Build libvtc variants with the attention ablation switches (attention_cs.cu: VTC_ACS_ABLATE) for tools/ab_attn_lib.py:
#   tools/build_ablate.sh 1 2 3 4 8 15   ->  tools/ab/libvtc_ablate<k>.so   (objects of the normal build are reused)
#   k = 10000000 r + 1000000 s + 100000 p + 10000 n + 1000 t + 100 e + a:  r = 1: with the %clock64 timeline stamps (tools/attn_trace_cs.py needs this build: VTC_LIB_PATH=tools/ab/libvtc_ablate10000000.so);  s = 1: the scalar form of the polynomial;  p = VTC_ACS_POLY (every p-th pair of exponentials on the FMA pipe);  n = 1 compiles the producer without the precomputed-mask-operand path;  a = VTC_ACS_ABLATE, e = VTC_ACS_EARLY_QK, t = 1 turns the TMA output store OFF (registers -> global)
set -e
cd "$(dirname "$0")/.."
python -m vision_transformer_cam_b200.build > /dev/null
mkdir -p tools/ab
OBJS=$(ls vision_transformer_cam_b200/build/*.o | grep -v attention_cs.o)
for k in "$@"; do
  if [[ $k == g* ]]; then      # gN: GEMM ablation N (gemm.cu: VTC_GEMM_ABLATE), timed with VTC_LIB_PATH=... python tools/prof_gemm.py
    n=${k#g}
    GOBJS=$(ls vision_transformer_cam_b200/build/*.o | grep -v "/gemm.o")
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
        -DVTC_GEMM_ABLATE=$n -c vision_transformer_cam_b200/csrc/gemm.cu -o /tmp/gemm_ablate$n.o
    /usr/local/cuda/bin/nvcc -shared -o tools/ab/libvtc_gemm_ablate$n.so $GOBJS /tmp/gemm_ablate$n.o -gencode arch=compute_100a,code=sm_100a -cudart static
    echo built tools/ab/libvtc_gemm_ablate$n.so
    continue
  fi
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
      -DVTC_ACS_ABLATE=$((k % 100)) -DVTC_ACS_EARLY_QK=$(((k / 100) % 10)) -DVTC_ACS_TMA_OUT=$((1 - (k / 1000) % 10)) -DVTC_ACS_NO_PREAUG=$(((k / 10000) % 10)) -DVTC_ACS_POLY=$(((k / 100000) % 10)) -DVTC_ACS_POLY_SCALAR=$(((k / 1000000) % 10)) -DVTC_ACS_TRACE=$((k / 10000000)) -Xptxas -v -c vision_transformer_cam_b200/csrc/attention_cs.cu -o /tmp/acs_ablate$k.o
  /usr/local/cuda/bin/nvcc -shared -o tools/ab/libvtc_ablate$k.so $OBJS /tmp/acs_ablate$k.o -gencode arch=compute_100a,code=sm_100a -cudart static
  echo built tools/ab/libvtc_ablate$k.so
done
