#!/bin/bash
# Round 2, GPU call 1 (1 GPU): every test that round 1 left gated, against the oracle on a B200; sanitizers over the
# pack kernels; isolated K4p/K4/K5 throughput; ncu --set full of K5 / K4 / K4p; a default bench line.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts/r2_call1.sh'
set -u
out=gpurun_out/r2c1
mkdir -p "$out"
export FMA_TEST_PACK_ON_GPU=1 FMA_TEST_IMAGE_ON_GPU=1 FMA_TEST_NEW_ON_GPU=1
nvidia-smi -L > "$out/gpus.txt" 2>&1; nproc >> "$out/gpus.txt"; free -g >> "$out/gpus.txt"

timeout 120 tests/cpp/cuda_emu/pack_kernels_gpu_test > "$out/pack_kernels_gpu_test.log" 2>&1; echo "kernel test rc=$?" | tee "$out/status.txt"
timeout 900 python -m pytest tests -m gpu -q -rfEs --timeout 240 > "$out/pytest_gpu_all.log" 2>&1; echo "pytest gpu (all gates open) rc=$?" | tee -a "$out/status.txt"
tail -40 "$out/pytest_gpu_all.log"

timeout 400 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pack_kernels_match" > "$out/sanitizer_memcheck_pack.log" 2>&1; echo "memcheck rc=$?" | tee -a "$out/status.txt"
timeout 400 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pack_kernels_match" > "$out/sanitizer_racecheck_pack.log" 2>&1; echo "racecheck rc=$?" | tee -a "$out/status.txt"

timeout 300 python scripts/pack_sweep.py > "$out/pack_sweep.log" 2>&1; echo "pack sweep rc=$?" | tee -a "$out/status.txt"
cp gpurun_out/sweep/pack_sweep.json "$out/" 2>/dev/null

timeout 500 python bench.py --steps 10 --warmup 3 > "$out/bench_default.json" 2> "$out/bench_default.err"; echo "bench default rc=$?" | tee -a "$out/status.txt"
timeout 300 env FMA_MAP_PIECE_MIB=2048 python bench.py --steps 16 --warmup 3 --no-cpu-baseline --packed-extra 0 > "$out/bench_pieces_2g.json" 2>> "$out/bench_default.err"; echo "bench pieces rc=$?" | tee -a "$out/status.txt"

for k in "fma_k_unpack$:k5_ldg:0" "fma_k_pack$:k4_ldg:0" "fma_k_pack_probe:k4p:0" "fma_k_unpack_tma:k5_tma:1" "fma_k_pack_tma:k4_tma:1"; do
  IFS=: read -r rx name variant <<< "$k"
  timeout 400 env FMA_PACK_KERNEL=$variant ncu --set full --clock-control none --import-source on -k "regex:$rx" -s 2 -c 1 -f -o "$out/${name}_full" \
      python bench.py --steps 1 --warmup 3 --contents bf16 --pack 1 --no-cpu-baseline --packed-extra 0 > "$out/ncu_$name.log" 2>&1
  echo "ncu $name rc=$?" | tee -a "$out/status.txt"
done
cat "$out/status.txt"
