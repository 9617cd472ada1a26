#!/bin/bash
# Round 2, GPU call 8 (4 GPUs, short): multi-path SLEEP (new) + wake parity tests and numbers.
set -u
out=gpurun_out/r2c8
mkdir -p "$out"
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -rfEs --timeout 200 -k "multipath or piecewise or phase_timeline or retried_wake or parking or peer" > "$out/pytest.log" 2>&1; echo "pytest rc=$?" | tee "$out/status.txt"; tail -5 "$out/pytest.log"
timeout 400 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --packed-extra 0 --extras multipath > "$out/bench_n1.json" 2> "$out/bench_n1.err"; echo "bench rc=$?" | tee -a "$out/status.txt"
python - "$out/bench_n1.json" <<'PY' | tee -a "$out/status.txt"
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('n1 e2e',d['e2e']['value'],'wake',d['wake_latency_s'],'sleep',d['sleep_latency_s'])
for r in (d.get('multipath_wake') or {}).get('rows',[]): print(json.dumps(r))
print('mp', (d.get('multipath_wake') or {}).get('bit_exact'), (d.get('multipath_wake') or {}).get('error'))
PY
cat "$out/status.txt"
