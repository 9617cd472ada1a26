"""Does a concurrent copy-engine H2D slow the K2 launch (fixed or proportional)?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fma_b200
from fma_b200 import _lib as L
eng = fma_b200.Engine(0)
eng.alloc(2048 * L.FMA_PAGE_BYTES, "default")
src = eng.segment(0).va
dst = eng.scratch_alloc(2048 * L.FMA_PAGE_BYTES)
h = torch.empty(8 << 30, dtype=torch.uint8, pin_memory=True)
d = torch.empty(8 << 30, dtype=torch.uint8, device="cuda")
s = torch.cuda.Stream()
def run(n, label, conc):
    ts = []
    for rep in range(3):
        if conc == "h2d":
            with torch.cuda.stream(s): d.copy_(h, non_blocking=True)
        elif conc == "d2h":
            with torch.cuda.stream(s): h.copy_(d, non_blocking=True)
        time.sleep(0.005)
        t = [eng.op_page_copy(n, src_base=src, dst_base=dst) for _ in range(20)]
        busy = not s.query()
        torch.cuda.synchronize()
        ts.append((sorted(t)[len(t) // 2], busy))
    ms, busy = min(ts)
    print(f"{label:10s} n={n*2:5d} MiB median {ms*1e3:7.1f} us  {2*n*L.FMA_PAGE_BYTES/ms/1e6:7.0f} GB/s copy_still_running={busy}", flush=True)
for n in (16, 64, 128, 256, 512):
    run(n, "alone", None); run(n, "with-h2d", "h2d"); run(n, "with-d2h", "d2h")
