#!/bin/bash
# Build libvtc variants with the attention ablation switches (attention_cs.cu: VTC_ACS_ABLATE) for tools/ab_attn_lib.py:
#   tools/build_ablate.sh 1 2 3 4 8 15   ->  tools/ab/libvtc_ablate<k>.so   (objects of the normal build are reused)
#   k = 10000 n + 1000 t + 100 e + a:  n = 1 compiles the producer without the precomputed-mask-operand path;  a = VTC_ACS_ABLATE, e = VTC_ACS_EARLY_QK, t = 1 turns the TMA output store OFF (registers -> global)
set -e
cd "$(dirname "$0")/.."
python -m vision_transformer_cam_b200.build > /dev/null
mkdir -p tools/ab
OBJS=$(ls vision_transformer_cam_b200/build/*.o | grep -v attention_cs.o)
for k in "$@"; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
      -DVTC_ACS_ABLATE=$((k % 100)) -DVTC_ACS_EARLY_QK=$(((k / 100) % 10)) -DVTC_ACS_TMA_OUT=$((1 - (k / 1000) % 10)) -DVTC_ACS_NO_PREAUG=$((k / 10000)) -c vision_transformer_cam_b200/csrc/attention_cs.cu -o /tmp/acs_ablate$k.o
  /usr/local/cuda/bin/nvcc -shared -o tools/ab/libvtc_ablate$k.so $OBJS /tmp/acs_ablate$k.o -gencode arch=compute_100a,code=sm_100a -cudart static
  echo built tools/ab/libvtc_ablate$k.so
done
