"""Isolated K1/K2 launch-size efficiency: GB/s (read+write) vs pages per launch, TMA configs, on scattered pages."""
import os, sys, json, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fma_b200
from fma_b200 import workloads as W, _lib as L
eng = fma_b200.Engine(0)
table = [s for s in W.allocation_table("llama-3-8b") if s.tag == "weights"]
for s in table: eng.alloc(s.bytes, s.tag)
pages = [s.va + o for s in eng.segments() for o in range(0, s.bytes, L.FMA_PAGE_BYTES)]
dst = eng.scratch_alloc(2048 * L.FMA_PAGE_BYTES)
rows = []
cfgs = [(32, 3, 2, 1), (32, 3, 1, 1), (16, 3, 2, 1), (16, 4, 2, 1), (32, 2, 2, 1), (64, 3, 1, 1), (16, 3, 2, 2), (8, 4, 4, 1), (16, 6, 2, 1), (32, 3, 1, 2)]
for n in (16, 32, 64, 128, 256, 512, 1024, 2048):
    src = pages[1000:1000 + n]
    for tile, stages, pipes, cps in cfgs:
        eng.set_option("tma_tile_bytes", tile << 10); eng.set_option("tma_stages", stages)
        eng.set_option("tma_pipes", pipes); eng.set_option("tma_ctas_per_sm", cps)
        ms = sorted(eng.op_page_copy(n, src_pages=src, dst_base=dst, variant=L.FMA_KERNEL_TMA) for _ in range(7))[1]
        rows.append(dict(n_pages=n, mib=n * 2, variant="tma", tile=tile, stages=stages, pipes=pipes, cps=cps, us=ms * 1e3, gbs=2 * n * L.FMA_PAGE_BYTES / ms / 1e6))
    ms = sorted(eng.op_page_copy(n, src_pages=src, dst_base=dst, variant=L.FMA_KERNEL_LDG) for _ in range(7))[1]
    rows.append(dict(n_pages=n, mib=n * 2, variant="ldg", us=ms * 1e3, gbs=2 * n * L.FMA_PAGE_BYTES / ms / 1e6))
    best = max((r for r in rows if r["n_pages"] == n), key=lambda r: r["gbs"])
    dflt = [r for r in rows if r["n_pages"] == n and r.get("tile") == 32 and r.get("stages") == 3 and r.get("pipes") == 2 and r.get("cps") == 1][0]
    print(n * 2, "MiB  default", round(dflt["us"], 1), "us", round(dflt["gbs"]), "GB/s | best", {k: (round(v, 1) if isinstance(v, float) else v) for k, v in best.items()}, flush=True)
os.makedirs("gpurun_out/sweep", exist_ok=True); json.dump(rows, open("gpurun_out/sweep/k_size_sweep.json", "w"), indent=1)
eng.close()
