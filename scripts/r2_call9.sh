#!/bin/bash
# Round 2, GPU call 9 (2 GPUs): the node agent (launcher-compatible REST, owning the parking buffers) with REAL vLLM children:
# an instance restricted to GPU 0 sleeps to the peer tier through the agent's parking service, is deleted asleep, and a new instance
# with the same ID adopts the parked image at start-up.   gpurun --gpus 2 --timeout 1200 -- 'bash scripts/r2_call9.sh'
set -u
out=gpurun_out/r2c9
mkdir -p "$out"
timeout 1000 env E2E_LAUNCHER=node_agent E2E_ARMS=fma_b200_peer_parked python scripts/e2e_launcher_vllm.py llama-1b > "$out/e2e_node_agent.log" 2>&1; echo "e2e node agent rc=$?" | tee "$out/status.txt"
tail -30 "$out/e2e_node_agent.log" | cut -c1-1500
cp gpurun_out/e2e/*.json "$out/" 2>/dev/null; cp gpurun_out/e2e/launcher.log "$out/node_agent.log" 2>/dev/null
for f in gpurun_out/e2e/*_vllm.log; do tail -c 20000 "$f" > "$out/$(basename $f)"; done
cat "$out/status.txt"
