#!/bin/bash
# Round 2, GPU call 4 (4 GPUs, short): multi-path wake with the NUMA-striped store; new K4/K5 (sub-page work items) sweep.
set -u
out=gpurun_out/r2c4
mkdir -p "$out"
nvidia-smi topo -m > "$out/topo.txt" 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_http.py -m gpu -q -rfEs --timeout 200 -k "multipath or pack or http or compiled" > "$out/pytest.log" 2>&1; echo "pytest rc=$?" | tee "$out/status.txt"
tail -6 "$out/pytest.log"
timeout 120 tests/cpp/cuda_emu/pack_kernels_gpu_test > "$out/pack_kernels_gpu_test.log" 2>&1; echo "kernel test rc=$?" | tee -a "$out/status.txt"
timeout 300 python scripts/pack_sweep.py > "$out/pack_sweep.log" 2>&1; echo "pack sweep rc=$?" | tee -a "$out/status.txt"; tail -7 "$out/pack_sweep.log"
cp gpurun_out/sweep/pack_sweep.json "$out/" 2>/dev/null
timeout 400 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --packed-extra 1 --extras multipath > "$out/bench_n1.json" 2> "$out/bench_n1.err"; echo "bench rc=$?" | tee -a "$out/status.txt"
python - "$out/bench_n1.json" <<'PY' | tee -a "$out/status.txt"
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('n1 e2e',d['e2e']['value'],'wake',d['wake_latency_s'],d['wake_latency_s_min_max'])
for r in (d.get('multipath_wake') or {}).get('rows',[]): print(json.dumps(r))
print('mp', (d.get('multipath_wake') or {}).get('bit_exact'), (d.get('multipath_wake') or {}).get('error'))
print('packed', json.dumps(d.get('packed_image'))[:700])
PY
for k in "fma_k_unpack$:k5" "fma_k_pack$:k4"; do
  IFS=: read -r rx name <<< "$k"
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$rx" -s 2 -c 1 -f -o "$out/${name}_full" \
      python bench.py --steps 1 --warmup 3 --contents bf16 --pack 1 --no-cpu-baseline --packed-extra 0 --extras none > "$out/ncu_$name.log" 2>&1
  echo "ncu $name rc=$?" | tee -a "$out/status.txt"
done
cat "$out/status.txt"
