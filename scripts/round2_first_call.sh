#!/bin/bash
# First GPU call of the next round (run under gpurun from the repo root):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts/round2_first_call.sh'
# Validates on a B200 what was written after round 1's GPU minutes were spent, cheapest and most informative first.
# Every step is bounded by `timeout`; outputs land in gpurun_out/round2/.
set -u
out=gpurun_out/round2
mkdir -p "$out"
export FMA_TEST_PACK_ON_GPU=1

# 0. the kernels alone (managed memory, no engine, no Python): pinpoints a device-side mismatch by page
timeout 120 tests/cpp/cuda_emu/pack_kernels_gpu_test > "$out/pack_kernels_gpu_test.log" 2>&1; echo "kernel test rc=$?" | tee "$out/status0.txt"

# 1. parity of the PACKED image kernels and engine path against the oracle, then the whole GPU suite
timeout 300 env FMA_TEST_PACK_KERNELS=0 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pack" > "$out/pytest_pack.log" 2>&1; echo "pytest pack (LDG/STG kernels) rc=$?" | tee "$out/status.txt"
timeout 300 env FMA_TEST_PACK_KERNELS=1 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pack and not binary" > "$out/pytest_pack_tma.log" 2>&1; echo "pytest pack (TMA kernels) rc=$?" | tee -a "$out/status.txt"
timeout 600 env FMA_TEST_PACK_KERNELS=0 python -m pytest tests -m gpu -x -q > "$out/pytest_gpu.log" 2>&1; echo "pytest gpu rc=$?" | tee -a "$out/status.txt"

# 2. memcheck + racecheck over the pack tests (small tables)
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pack_kernels_match" > "$out/sanitizer_memcheck_pack.log" 2>&1; echo "memcheck rc=$?" | tee -a "$out/status.txt"
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pack_kernels_match" > "$out/sanitizer_racecheck_pack.log" 2>&1; echo "racecheck rc=$?" | tee -a "$out/status.txt"

# 3. image hand-over (memfd store + descriptor) on the GPU: the tests are gated on FMA_TEST_IMAGE_ON_GPU today
timeout 300 env FMA_TEST_IMAGE_ON_GPU=1 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "image" > "$out/pytest_image.log" 2>&1; echo "image rc=$?" | tee -a "$out/status.txt"

# 4. bench: default line (with the packed_image extra), then bf16 contents plain vs packed
timeout 600 python bench.py --steps 10 --warmup 3 > "$out/bench_default.json" 2> "$out/bench_default.err"; echo "bench default rc=$?" | tee -a "$out/status.txt"
timeout 400 python bench.py --steps 10 --warmup 3 --contents bf16 --pack 0 --no-cpu-baseline --packed-extra 0 > "$out/bench_bf16_plain.json" 2>> "$out/bench_default.err"
timeout 400 python bench.py --steps 10 --warmup 3 --contents bf16 --pack 1 --no-cpu-baseline --packed-extra 0 > "$out/bench_bf16_packed.json" 2>> "$out/bench_default.err"; echo "bench packed rc=$?" | tee -a "$out/status.txt"
timeout 400 env FMA_PACK_KERNEL=1 python bench.py --steps 10 --warmup 3 --contents bf16 --pack 1 --no-cpu-baseline --packed-extra 0 > "$out/bench_bf16_packed_tma.json" 2>> "$out/bench_default.err"; echo "bench packed (TMA kernels) rc=$?" | tee -a "$out/status.txt"

# 4a. isolated K4p / K4 / K5 throughput, both variants (under timeout: the TMA variant has never run on a GPU)
timeout 300 python scripts/pack_sweep.py > "$out/pack_sweep.log" 2>&1; echo "pack sweep rc=$?" | tee -a "$out/status.txt"

# 4b. VMM granularity / VA alignment probe (seconds)
timeout 120 python scripts/gran_probe.py > "$out/gran_probe.log" 2>&1

# 5. ncu: launch list of the packed bench, then one full capture of K5 and K4
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$out/launches_packed.csv" \
    python bench.py --steps 2 --warmup 3 --contents bf16 --pack 1 --no-cpu-baseline --packed-extra 0 > "$out/ncu_launches.log" 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fma_k_unpack -c 1 -o "$out/k5_full" \
    python bench.py --steps 1 --warmup 3 --contents bf16 --pack 1 --no-cpu-baseline --packed-extra 0 > "$out/ncu_k5.log" 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fma_k_pack$ -c 1 -o "$out/k4_full" \
    python bench.py --steps 1 --warmup 3 --contents bf16 --pack 1 --no-cpu-baseline --packed-extra 0 > "$out/ncu_k4.log" 2>&1
# 6. end to end through the unmodified reference launcher + real vLLM: packed image over HTTP, and --load-format fma
timeout 900 env E2E_ARMS=fma_b200,fma_b200_packed python scripts/e2e_launcher_vllm.py llama-3-8b > "$out/e2e_packed.log" 2>&1; echo "e2e packed rc=$?" | tee -a "$out/status.txt"
timeout 900 env E2E_ARMS=ckpt_default,ckpt_fma python scripts/e2e_launcher_vllm.py llama-1b > "$out/e2e_ckpt.log" 2>&1; echo "e2e ckpt rc=$?" | tee -a "$out/status.txt"
# 7. the compiled host side over HTTP (no vLLM, no Python in the serving process)
timeout 300 python scripts/native_server_e2e.py llama-3-8b 1 0 > "$out/native_server.log" 2>&1; echo "native server rc=$?" | tee -a "$out/status.txt"
cat "$out/status.txt"
