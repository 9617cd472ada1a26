import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["FMA_DEBUG_VMM"] = "1"
import fma_b200
from fma_b200 import workloads as W, _lib as L
for kv in (32,):
    eng = fma_b200.Engine(0)
    table = W.allocation_table("llama-3-8b", kv_cache_bytes=kv << 30)
    ptrs = [eng.alloc(s.bytes, s.tag) for s in table]
    Wb = W.weight_bytes(table)
    eng.host_reserve(Wb)
    for mode in (L.FMA_MODE_DIRECT, L.FMA_MODE_STAGED):
        eng.set_option("mode", mode)
        for r in range(6):
            eng.sleep(["weights"]); s1 = eng.stats(); eng.wake(None); s2 = eng.stats()
            print(f"kv={kv} mode={mode} rep={r} sleep={s1['sleep_seconds']:.3f} copy={s1['sleep_copy_seconds']:.3f} unmap={s1['sleep_unmap_seconds']:.3f} wake={s2['wake_seconds']:.4f} map={s2['wake_map_seconds']:.3f}", flush=True)
    eng.close()
