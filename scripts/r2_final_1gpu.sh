#!/bin/bash
# Round 2, final 1-GPU check of HEAD exactly as the driver runs it: pytest -m gpu, smoke(), bench.py with default flags, and the
# reference arm.   gpurun --timeout 900 -- 'bash scripts/r2_final_1gpu.sh'
set -u
out=gpurun_out/r2final1
mkdir -p "$out"
timeout 600 python -m pytest tests -x -q -m gpu > "$out/pytest_gpu.log" 2>&1; echo "pytest gpu rc=$?" | tee "$out/status.txt"; tail -4 "$out/pytest_gpu.log"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > "$out/smoke.log" 2>&1; echo "smoke rc=$?" | tee -a "$out/status.txt"; tail -1 "$out/smoke.log"
( time timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 > "$out/bench.json" 2> "$out/bench.err" ) 2> "$out/bench.time"; echo "bench rc=$? $(grep real $out/bench.time)" | tee -a "$out/status.txt"
python - "$out/bench.json" <<'PY' | tee -a "$out/status.txt"
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'mean',d['e2e']['mean_gbs'],'wake',d['wake_latency_s'],d['wake_latency_s_min_max'],'ms_per_step',d['ms_per_step'],'W',d['config']['weights_gib_per_rank'],'segs',d['config']['segments_per_rank'],'naive',d['pcie']['naive_pinned_h2d_per_gpu'],d['pcie']['vs_naive_pinned_h2d'])
print('roofline',json.dumps(d['roofline'])[:500])
print('cpu_baseline',json.dumps(d.get('cpu_baseline'))[:260])
print('swap',json.dumps(d.get('swap_config4'))[:400])
print('scaling_base',json.dumps(d.get('n1_on_scaling_workload'))[:300])
print('packed',json.dumps(d.get('packed_image'))[:300])
print('clocks',json.dumps(d.get('clocks'))[:200])
PY
( time timeout 600 python bench.py --impl reference --gpus 1 --steps 5 --warmup 3 > "$out/bench_ref.json" 2> "$out/bench_ref.err" ) 2> "$out/bench_ref.time"; echo "ref rc=$? $(grep real $out/bench_ref.time)" | tee -a "$out/status.txt"
tail -c 700 "$out/bench_ref.json"
cat "$out/status.txt"
