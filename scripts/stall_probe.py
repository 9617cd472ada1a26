"""Are the sporadic +10..90 ms wake stalls an artefact of waking IMMEDIATELY after a sleep (the driver still scrubbing /
returning the 48 GiB just released)?  Same loop with an idle gap between sleep and wake."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fma_b200
from fma_b200 import workloads as W
eng = fma_b200.Engine(0)
table = W.allocation_table("llama-3-8b", kv_cache_bytes=32 << 30)
for s in table: eng.alloc(s.bytes, s.tag)
eng.host_reserve(W.weight_bytes(table))
for _ in range(3): eng.sleep(["weights"]); eng.wake(None)
out = {}
for gap in (0.0, 0.25, 1.0, 0.0):
    ws = []
    for _ in range(16):
        eng.sleep(["weights"]); time.sleep(gap); eng.wake(None); ws.append(round(eng.stats()["wake_seconds"], 4))
    out[f"gap_{gap}"] = ws
    print("gap", gap, "s: outliers>0.3:", sum(1 for x in ws if x > 0.3), "mean", round(sum(ws) / len(ws), 4), ws, flush=True)
os.makedirs("gpurun_out/stall", exist_ok=True); json.dump(out, open("gpurun_out/stall/stall_probe.json", "w"), indent=1)
