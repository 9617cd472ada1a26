#!/bin/bash
# Second GPU call of the next round, once scripts/round2_first_call.sh is green:
#   /usr/local/graft/bin/gpurun --timeout 1800 -- 'bash scripts/round2_profiles.sh'
# Sweeps, ncu captures and end-to-end runs of the code written after round 1's GPU minutes were spent.
set -u
out=gpurun_out/round2
mkdir -p "$out"
: > "$out/status_profiles.txt"

# 1. isolated K4p / K4 / K5 throughput, both variants
timeout 300 python scripts/pack_sweep.py > "$out/pack_sweep.log" 2>&1; echo "pack sweep rc=$?" | tee -a "$out/status_profiles.txt"

# 1b. piecewise mapping of the weights run: per-step wake latencies with and without (24 steps each)
timeout 400 python bench.py --steps 24 --warmup 3 --no-cpu-baseline --packed-extra 0 > "$out/bench_24_whole_runs.json" 2> "$out/bench_pieces.err"
timeout 400 env FMA_MAP_PIECE_MIB=2048 python bench.py --steps 24 --warmup 3 --no-cpu-baseline --packed-extra 0 > "$out/bench_24_pieces_2g.json" 2>> "$out/bench_pieces.err"; echo "bench pieces rc=$?" | tee -a "$out/status_profiles.txt"

# (N>1, run separately, 8x charged:  gpurun --gpus 8 -- 'python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1
#   --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 --extras packed,incremental'  -> one run carries plain + packed + incremental)

# 2. VMM granularity / VA alignment probe (seconds)
timeout 120 python scripts/gran_probe.py > "$out/gran_probe.log" 2>&1

# 3. ncu: launch list of the packed bench, then one full capture of K5 and K4
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$out/launches_packed.csv" \
    python bench.py --steps 2 --warmup 3 --contents bf16 --pack 1 --no-cpu-baseline --packed-extra 0 > "$out/ncu_launches.log" 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fma_k_unpack -c 1 -o "$out/k5_full" \
    python bench.py --steps 1 --warmup 3 --contents bf16 --pack 1 --no-cpu-baseline --packed-extra 0 > "$out/ncu_k5.log" 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fma_k_pack$ -c 1 -o "$out/k4_full" \
    python bench.py --steps 1 --warmup 3 --contents bf16 --pack 1 --no-cpu-baseline --packed-extra 0 > "$out/ncu_k4.log" 2>&1
# 4. end to end through the unmodified reference launcher + real vLLM: packed image over HTTP, and --load-format fma
timeout 900 env E2E_ARMS=fma_b200,fma_b200_packed python scripts/e2e_launcher_vllm.py llama-3-8b > "$out/e2e_packed.log" 2>&1; echo "e2e packed rc=$?" | tee -a "$out/status_profiles.txt"
timeout 900 env E2E_ARMS=ckpt_default,ckpt_fma python scripts/e2e_launcher_vllm.py llama-1b > "$out/e2e_ckpt.log" 2>&1; echo "e2e ckpt rc=$?" | tee -a "$out/status_profiles.txt"
# 5. the compiled host side over HTTP (no vLLM, no Python in the serving process)
timeout 300 python scripts/native_server_e2e.py llama-3-8b 1 0 > "$out/native_server.log" 2>&1; echo "native server rc=$?" | tee -a "$out/status_profiles.txt"
cat "$out/status_profiles.txt"
