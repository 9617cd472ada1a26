"""Isolated K4p / K4 / K5 throughput (round 2): algorithmic GB/s (read + write) vs pages per launch, for the LDG/STG and the
TMA-pipelined variants, on bf16 dummy weights scattered over the Llama-3-8B table.  Also re-checks bit-exactness of every
launch size.  Output: gpurun_out/sweep/pack_sweep.json (summarise into profiles/)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fma_b200
from fma_b200 import workloads as W, _lib as L

PAGE, PACKED = L.FMA_PAGE_BYTES, L.FMA_PACKED_PAGE_BYTES
eng = fma_b200.Engine(0)
table = [s for s in W.allocation_table("llama-3-8b") if s.tag == "weights"][:40]          # ~4.6 GiB is plenty
for s in table: eng.alloc(s.bytes, s.tag)
gen = torch.Generator(device="cuda"); gen.manual_seed(7)
blk = torch.empty(256 << 20, dtype=torch.bfloat16, device="cuda").uniform_(-1e-3, 1e-3, generator=gen)
host = torch.empty(256 << 20, dtype=torch.bfloat16, pin_memory=True).copy_(blk); del blk
for i, s in enumerate(table):
    o = 0
    while o < s.bytes:
        m = min(s.bytes - o, host.numel() * 2)
        eng.write_ptr(i, host.data_ptr(), m, offset=o); o += m
pages = [s.va + o for s in eng.segments() for o in range(0, s.bytes, PAGE)]
store = eng.scratch_alloc(1024 * PAGE)
back = eng.scratch_alloc(1024 * PAGE)
want = eng.op_page_digest(1024, pages=pages[100:1124])[0]
rows = []
for n in (37, 148, 256, 337, 592, 1024):
    src = pages[100:100 + n]
    sizes, ms_probe = eng.op_pack_probe(n, pages=src)
    assert all(b == PACKED for b in sizes), "dummy weights must code"
    rows.append(dict(kernel="K4p", n_pages=n, us=ms_probe * 1e3, gbs=n * PAGE / ms_probe / 1e6))
    for variant in (0,):
        t4 = sorted(eng.op_pack(sizes, store, src_pages=src) for _ in range(7))[1]
        t5 = sorted(eng.op_unpack(sizes, store, dst_base=back) for _ in range(7))[1]
        ok = eng.op_page_digest(n, base=back, first_word=[0] * n)[0] == eng.op_page_digest(n, pages=src, first_word=[0] * n)[0]
        alg = n * (PAGE + PACKED)
        rows.append(dict(kernel="K4", variant="tma" if variant else "ldg", n_pages=n, us=t4 * 1e3, gbs=alg / t4 / 1e6, bit_exact=ok))
        rows.append(dict(kernel="K5", variant="tma" if variant else "ldg", n_pages=n, us=t5 * 1e3, gbs=alg / t5 / 1e6, bit_exact=ok))
        print(n, "pages", "tma" if variant else "ldg", f"K4 {t4 * 1e3:.1f} us {alg / t4 / 1e6:.0f} GB/s | K5 {t5 * 1e3:.1f} us {alg / t5 / 1e6:.0f} GB/s | exact {ok}", flush=True)
os.makedirs("gpurun_out/sweep", exist_ok=True)
json.dump(rows, open("gpurun_out/sweep/pack_sweep.json", "w"), indent=1)
eng.close()
