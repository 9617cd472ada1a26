#!/bin/bash
# Last GPU call of the round: HEAD once more — pytest -m gpu, smoke(), a short default bench.
set -u
out=gpurun_out/r2final2
mkdir -p "$out"
timeout 600 python -m pytest tests -x -q -m gpu > "$out/pytest_gpu.log" 2>&1; echo "pytest gpu rc=$?" | tee "$out/status.txt"; tail -3 "$out/pytest_gpu.log"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > "$out/smoke.log" 2>&1; echo "smoke rc=$?" | tee -a "$out/status.txt"; tail -1 "$out/smoke.log"
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --packed-extra 0 --extras swap > "$out/bench.json" 2> "$out/bench.err"; echo "bench rc=$?" | tee -a "$out/status.txt"
python - "$out/bench.json" <<'PY' | tee -a "$out/status.txt"
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'wake',d['wake_latency_s'],d['wake_latency_s_min_max'],'traffic',d['roofline']['traffic'],'swap',json.dumps(d.get('swap_config4'))[:200])
PY
