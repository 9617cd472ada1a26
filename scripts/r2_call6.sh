#!/bin/bash
# Round 2, GPU call 6 (1 GPU): the whole GPU suite as the driver runs it, smoke(), pack sweep (K5 at 32 registers), hot swap with and
# without the VMM gate, the default bench line.   gpurun --timeout 900 -- 'bash scripts/r2_call6.sh'
set -u
out=gpurun_out/r2c6
mkdir -p "$out"
timeout 600 python -m pytest tests -m gpu -q -rfEs --timeout 240 > "$out/pytest_gpu.log" 2>&1; echo "pytest gpu rc=$?" | tee "$out/status.txt"; tail -12 "$out/pytest_gpu.log"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > "$out/smoke.log" 2>&1; echo "smoke rc=$?" | tee -a "$out/status.txt"; tail -2 "$out/smoke.log"
timeout 300 python scripts/pack_sweep.py > "$out/pack_sweep.log" 2>&1; echo "pack sweep rc=$?" | tee -a "$out/status.txt"; tail -6 "$out/pack_sweep.log"; cp gpurun_out/sweep/pack_sweep.json "$out/" 2>/dev/null
for g in 0 1; do
  timeout 300 env FMA_VMM_GATE=$g python bench.py --steps 3 --warmup 3 --no-cpu-baseline --packed-extra 0 --extras swap > "$out/bench_swap_gate$g.json" 2> "$out/bench_swap_gate$g.err"
  echo "swap gate=$g rc=$? $(python -c "
import json,sys
d=json.loads(open('$out/bench_swap_gate$g.json').read().strip().splitlines()[-1]); print(json.dumps(d.get('swap_config4')))")" | tee -a "$out/status.txt"
done
timeout 600 python bench.py --steps 10 --warmup 3 > "$out/bench_default.json" 2> "$out/bench_default.err"; echo "bench default rc=$?" | tee -a "$out/status.txt"
python - "$out/bench_default.json" <<'PY' | tee -a "$out/status.txt"
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('n1 value',d['value'],'e2e',d['e2e']['value'],'mean',d['e2e']['mean_gbs'],'wake',d['wake_latency_s'],d['wake_latency_s_min_max'],'ms_per_step',d['ms_per_step'])
print('cpu_baseline',json.dumps(d.get('cpu_baseline'))[:300])
print('packed',json.dumps(d.get('packed_image'))[:400])
print('swap',json.dumps(d.get('swap_config4')))
PY
cat "$out/status.txt"
