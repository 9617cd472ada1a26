#!/bin/bash
# Round 2, GPU call 2 (2 GPUs): peer-tier parity tests (need a second GPU), then the N=2 A/B of the wake's VMM ordering with
# per-rank phase timelines.   /usr/local/graft/bin/gpurun --gpus 2 --timeout 1200 -- 'bash scripts/r2_call2.sh'
set -u
out=gpurun_out/r2c2
mkdir -p "$out"
nvidia-smi -L > "$out/gpus.txt" 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -rfEs --timeout 200 -k "peer or parking" > "$out/pytest_peer.log" 2>&1; echo "pytest peer rc=$?" | tee "$out/status.txt"
tail -5 "$out/pytest_peer.log"
run() {   # label, env..., then "--" and extra bench args
  label=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  timeout 240 env "${envs[@]}" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus 2 --steps 8 --warmup 3 --no-cpu-baseline --timeline "$out/tl_$label" "$@" > "$out/bench_$label.json" 2> "$out/bench_$label.err"
  echo "bench $label rc=$? $(python - "$out/bench_$label.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print('e2e',d['e2e']['value'],'wake',d['wake_latency_s'],'median',d['wake_latency_s_median'],'map_s',d['wake_map_s'],'sleep',d['sleep_latency_s'],'naive',d['pcie']['naive_pinned_h2d_per_gpu'],'vs_naive',d['pcie']['vs_naive_pinned_h2d'], 'peer', (d.get('peer_tier') or {}).get('wake_latency_s'), (d.get('peer_tier') or {}).get('frac_of_nvlink_900'))
except Exception as e: print('parse error',e)
PY
)" | tee -a "$out/status.txt"
}
run A_whole   FMA_MAP_PIECE_MIB=0 -- --peer-extra 0
run B_pieces  FMA_X=0 -- --peer-extra 0
run C_gate    FMA_VMM_GATE=1 -- --peer-extra 0
run D_after   FMA_REMAP_AFTER_COPY=1 -- --peer-extra 0
run E_direct  FMA_MAP_PIECE_MIB=1024 -- --peer-extra 0 --mode direct
run F_gate_after FMA_VMM_GATE=1 FMA_REMAP_AFTER_COPY=1 -- --peer-extra 0
run G_pieces512 FMA_MAP_PIECE_MIB=512 FMA_VMM_GATE=1 -- --peer-extra 1
run H_peer_nogate FMA_X=0 -- --peer-extra 1 --steps 4
cat "$out/status.txt"
