#!/bin/bash
# compute-sanitizer over the REWRITTEN pack kernels (sub-page work items, global exception counter, last-part header) and K1/K2.
set -u
out=gpurun_out/r2san
mkdir -p "$out"
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pack_kernels_match or page_gather_scatter" > "$out/memcheck.log" 2>&1; echo "memcheck rc=$?" | tee "$out/status.txt"; tail -3 "$out/memcheck.log"
timeout 400 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pack_kernels_match" > "$out/racecheck.log" 2>&1; echo "racecheck rc=$?" | tee -a "$out/status.txt"; tail -3 "$out/racecheck.log"
timeout 200 compute-sanitizer --tool memcheck tests/cpp/cuda_emu/pack_kernels_gpu_test > "$out/memcheck_kernel_binary.log" 2>&1; echo "memcheck binary rc=$?" | tee -a "$out/status.txt"; tail -3 "$out/memcheck_kernel_binary.log"
