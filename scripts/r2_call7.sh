#!/bin/bash
# Round 2, GPU call 7 (8 GPUs): A/B of the cross-process VMM gate with piece-index priority classes at N=8 (host tier + peer tier),
# and multi-path wake at N=1 with the remap head start scaled to the number of paths.
set -u
out=gpurun_out/r2c7
mkdir -p "$out"
run8() {
  label=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  timeout 300 env "${envs[@]}" python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29589 \
      bench.py --gpus 8 --steps 10 --warmup 3 --timeline "$out/tl_$label" "$@" > "$out/bench_$label.json" 2> "$out/bench_$label.err"
  echo "bench $label rc=$? $(python - "$out/bench_$label.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print('value',d['value'],'e2e',d['e2e']['value'],'mean',d['e2e']['mean_gbs'],'wake',d['wake_latency_s'],d['wake_latency_s_min_max'],'sleep',d['sleep_latency_s'],'vs_naive',d['pcie']['vs_naive_pinned_h2d'],'peer',json.dumps(d.get('peer_tier')),'rr',json.dumps(d.get('roundrobin_config5')))
except Exception as e: print('parse error',e)
PY
)" | tee -a "$out/status.txt"
}
: > "$out/status.txt"
run8 gate FMA_VMM_GATE=1 -- --extras roundrobin
run8 nogate FMA_VMM_GATE=0 -- --extras none
timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --packed-extra 0 --extras multipath > "$out/bench_n1_on8.json" 2> "$out/bench_n1_on8.err"; echo "bench n1 rc=$?" | tee -a "$out/status.txt"
python - "$out/bench_n1_on8.json" <<'PY' | tee -a "$out/status.txt"
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
for r in (d.get('multipath_wake') or {}).get('rows',[]): print(json.dumps(r))
print('mp', (d.get('multipath_wake') or {}).get('bit_exact'), (d.get('multipath_wake') or {}).get('error'))
PY
cat "$out/status.txt"
