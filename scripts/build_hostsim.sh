#!/bin/bash
# Developer helper: build the host-simulated engine (tests/cpp/hostsim, test infrastructure) into /tmp/hs for quick CPU iterations:
#   bash scripts/build_hostsim.sh && FMA_B200_LIB=/tmp/hs/libfma_b200_hostsim.so FMA_HOSTSIM=1 HOSTSIM_DEVICES=2 python -m pytest tests/test_gpu_parity.py -m gpu -k <expr>
set -e
cd "$(dirname "$0")/.."
mkdir -p /tmp/hs
C=llm-d-fast-model-actuation_b200/csrc
g++ -std=c++17 -O1 -g -shared -fPIC -fvisibility=hidden -I/usr/local/cuda/include -I$C -Iinclude -x c++ \
  $C/fma_engine.cu $C/fma_sleep.cu $C/fma_wake.cu $C/fma_load.cu $C/fma_image.cu $C/fma_gate.cu $C/fma_pull.cu \
  tests/cpp/hostsim/hostsim_cuda.cpp tests/cpp/hostsim/hostsim_kernels.cpp -o /tmp/hs/libfma_b200_hostsim.so -lpthread
echo /tmp/hs/libfma_b200_hostsim.so
