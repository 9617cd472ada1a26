#!/bin/bash
# Round 2, GPU call 3 (4 GPUs): new parity tests (multi-path wake, node-level parking buffer across processes, peer tier), the
# default N=1 bench line with its extras (swap, scaling base, multi-path over 1 and 3 helpers), then N=2 and N=4 with the
# executor-style phase barriers and timelines.   gpurun --gpus 4 --timeout 1200 -- 'bash scripts/r2_call3.sh'
set -u
out=gpurun_out/r2c3
mkdir -p "$out"
nvidia-smi -L > "$out/gpus.txt" 2>&1; nvidia-smi topo -m >> "$out/gpus.txt" 2>&1
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -rfEs --timeout 200 -k "multipath or parking or peer" > "$out/pytest_new.log" 2>&1; echo "pytest new rc=$?" | tee "$out/status.txt"
tail -8 "$out/pytest_new.log"
timeout 600 python bench.py --steps 10 --warmup 3 --timeline "$out/tl_n1" > "$out/bench_n1.json" 2> "$out/bench_n1.err"; echo "bench n1 rc=$?" | tee -a "$out/status.txt"
python - "$out/bench_n1.json" <<'PY' | tee -a "$out/status.txt"
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print('n1 value',d['value'],'e2e',d['e2e'],'wake',d['wake_latency_s'],d['wake_latency_s_min_max'])
    print('swap',json.dumps(d.get('swap_config4')))
    print('scaling_base',json.dumps(d.get('n1_on_scaling_workload')))
    print('multipath',json.dumps(d.get('multipath_wake')))
    print('cpu_baseline',json.dumps(d.get('cpu_baseline'))[:400])
    print('packed', json.dumps(d.get('packed_image'))[:500])
except Exception as e: print('parse error',e)
PY
for n in 2 4; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n \
      bench.py --gpus $n --steps 10 --warmup 3 --timeline "$out/tl_n$n" > "$out/bench_n$n.json" 2> "$out/bench_n$n.err"
  echo "bench n$n rc=$? $(python - "$out/bench_n$n.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print('value',d['value'],'e2e',d['e2e']['value'],'mean',d['e2e']['mean_gbs'],'wake',d['wake_latency_s'],d['wake_latency_s_min_max'],'sleep',d['sleep_latency_s'],'vs_naive',d['pcie']['vs_naive_pinned_h2d'],'peer',json.dumps(d.get('peer_tier')),'rr',json.dumps(d.get('roundrobin_config5')))
except Exception as e: print('parse error',e)
PY
)" | tee -a "$out/status.txt"
done
cat "$out/status.txt"
