#!/bin/bash
# Round 2: the ncu launch list of bench.py (main line only) and a fresh --set full capture of K2 (refreshes profiles/k2_traffic.json).
set -u
out=gpurun_out/r2ncu
mkdir -p "$out"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file "$out/launches_r2.csv" \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --packed-extra 0 --extras none --kv-gib 4 > "$out/ncu_launches.log" 2>&1; echo "launch list rc=$?" | tee "$out/status.txt"
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:fma_k_page_copy_tma" -s 12 -c 3 -f -o "$out/k2_full_r2" \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --packed-extra 0 --extras none --kv-gib 4 > "$out/ncu_k2.log" 2>&1; echo "k2 full rc=$?" | tee -a "$out/status.txt"
cat "$out/status.txt"
