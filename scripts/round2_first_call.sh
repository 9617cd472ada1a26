#!/bin/bash
# First GPU call of the next round (run under gpurun from the repo root):
#   /usr/local/graft/bin/gpurun --timeout 1200 -- 'bash scripts/round2_first_call.sh'      (≈10-12 min when everything passes)
# The slower evidence runs (sweeps, ncu, end-to-end) are in scripts/round2_profiles.sh.
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

# 5. INCREMENTAL sleep (gated test + bench: sleeps after the first move nothing)
timeout 300 env FMA_TEST_NEW_ON_GPU=1 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "incremental or random_alloc" > "$out/pytest_new.log" 2>&1; echo "pytest incremental + random histories rc=$?" | tee -a "$out/status.txt"
timeout 400 python bench.py --steps 10 --warmup 3 --incremental 1 --no-cpu-baseline --packed-extra 0 > "$out/bench_incremental.json" 2>> "$out/bench_default.err"; echo "bench incremental rc=$?" | tee -a "$out/status.txt"

cat "$out/status.txt"
