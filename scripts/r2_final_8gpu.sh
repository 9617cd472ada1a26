#!/bin/bash
# Round 2, final N=8 run of HEAD with default flags, as the driver's scaling bench launches it.
set -u
out=gpurun_out/r2final8
mkdir -p "$out"
( time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29591 \
    bench.py --gpus 8 --steps 10 --warmup 3 --timeline "$out/tl" > "$out/bench_n8.json" 2> "$out/bench_n8.err" ) 2> "$out/time.txt"; echo "bench n8 rc=$? $(grep real $out/time.txt)" | tee "$out/status.txt"
python - "$out/bench_n8.json" <<'PY' | tee -a "$out/status.txt"
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'mean',d['e2e']['mean_gbs'],'wake',d['wake_latency_s'],d['wake_latency_s_min_max'],'sleep',d['sleep_latency_s'],'ms_per_step',d['ms_per_step'],'W',d['config']['weights_gib_per_rank'],'naive',d['pcie']['naive_pinned_h2d_per_gpu'],d['pcie']['vs_naive_pinned_h2d'])
print('peer',json.dumps(d.get('peer_tier')))
print('rr',json.dumps(d.get('roundrobin_config5')))
print('clocks',json.dumps(d.get('clocks'))[:200])
PY
