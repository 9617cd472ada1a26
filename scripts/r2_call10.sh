#!/bin/bash
# Round 2, GPU call 10 (1 GPU): unmodified reference launcher + real vLLM (Llama-3-8B shapes, dummy weights) with the engine underneath:
# POST /sleep / wake_up wall times, tokens before == after, and the synthetic allocation table validated against the live weights pool.
set -u
out=gpurun_out/r2c10
mkdir -p "$out"
timeout 900 env E2E_ARMS=fma_b200 python scripts/e2e_launcher_vllm.py llama-3-8b > "$out/e2e.log" 2>&1; echo "e2e rc=$?" | tee "$out/status.txt"
tail -5 "$out/e2e.log" | cut -c1-3000
cp gpurun_out/e2e/e2e_launcher_vllm_llama-3-8b_tp1.json "$out/" 2>/dev/null
python - <<'PY' | tee -a "$out/status.txt"
import json
try:
    d=json.load(open('gpurun_out/e2e/e2e_launcher_vllm_llama-3-8b_tp1.json'))
    print(json.dumps(d['fma_b200'].get('table_validation'), indent=1)[:3000])
    print([ (round(r['sleep_s'],3), round(r['wake_s'],3)) for r in d['fma_b200']['rows']], d['fma_b200']['same_tokens'])
except Exception as e: print('parse error', e)
PY
