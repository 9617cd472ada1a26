"""First GPU contact: correctness of K0/K1/K2/K3 + every sleep/wake mode vs the oracle, then timings.
Writes gpurun_out/first/*.json.  Exploratory (not the bench, not a test)."""
import ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))  # repo root
import numpy as np
import fma_b200
from fma_b200 import workloads as W, _lib as L
from oracle import oracle as O

OUT = "gpurun_out/first"; os.makedirs(OUT, exist_ok=True)
res = {}
def log(*a):
    print(*a, flush=True)

eng = fma_b200.Engine(0)
# ---- correctness on the tiny table ----------------------------------------------------------
table = W.allocation_table("tiny-llama-test", kv_cache_bytes=64 << 20, kv_tensors=2)
ptrs = [eng.alloc(s.bytes, s.tag) for s in table]
wseg = [i for i, s in enumerate(table) if s.tag == "weights"]
first = 0; ref = {}
for i in wseg:
    eng.fill(i, 1234, first)
    ref[i] = O.fill(table[i].bytes, 1234, first)
    first += table[i].bytes // 8
ok_fill = all(eng.read(i, table[i].bytes) == ref[i].tobytes() for i in wseg)
dg = eng.digest_all(["weights"])
ok_dig = all(dg[i] == O.digest(ref[i]) for i in wseg)
log("fill bit-exact", ok_fill, "digest bit-exact", ok_dig)
res["fill_ok"], res["digest_ok"] = ok_fill, ok_dig
image = O.packed_image([ref[i] for i in wseg])
for tier, tname in [(L.FMA_TIER_HOST, "host"), (L.FMA_TIER_LOCAL, "local")]:
    for mode, mname in [(L.FMA_MODE_DIRECT, "direct"), (L.FMA_MODE_STAGED, "staged"), (L.FMA_MODE_KERNEL, "kernel")]:
        for kern, kname in [(L.FMA_KERNEL_TMA, "tma"), (L.FMA_KERNEL_LDG, "ldg")]:
            if mode == L.FMA_MODE_DIRECT and kern == L.FMA_KERNEL_LDG: continue
            eng.set_option("mode", mode); eng.set_option("kernel", kern); eng.set_option("chunk_bytes", 8 << 20)
            eng.sleep(["weights"], tier=tier, flags=L.FMA_FLAG_VERIFY)
            assert eng.is_sleeping()
            img_ok = None
            if tier == L.FMA_TIER_HOST:
                base, n = eng.host_store_view()
                got = np.ctypeslib.as_array((C.c_uint8 * n).from_address(base))
                img_ok = bool(n == image.size and np.array_equal(got, image))
            eng.wake(None, flags=L.FMA_FLAG_VERIFY)
            same_va = [s.va for s in eng.segments()] == ptrs
            back = all(eng.read(i, table[i].bytes) == ref[i].tobytes() for i in wseg)
            log(tname, mname, kname, "image", img_ok, "roundtrip", back, "same_va", same_va, "sleeping", eng.is_sleeping())
            res[f"rt_{tname}_{mname}_{kname}"] = dict(image=img_ok, roundtrip=back, same_va=same_va)
for p in ptrs: eng.free(p)
json.dump(res, open(f"{OUT}/correctness.json", "w"), indent=1)

# ---- timings on the Llama-3-8B table ----------------------------------------------------------
eng.set_option("chunk_bytes", 32 << 20)
table = W.allocation_table("llama-3-8b", kv_cache_bytes=32 << 30)
t0 = time.time(); ptrs = [eng.alloc(s.bytes, s.tag) for s in table]; t_alloc = time.time() - t0
Wb = W.weight_bytes(table)
log("alloc s", t_alloc, "W GiB", Wb / 2**30)
first = 0
for i, s in enumerate(table):
    if s.tag == "weights":
        eng.fill(i, 1234, first); first += s.bytes // 8
before = eng.digest_all(["weights"])
t0 = time.time(); eng.host_reserve(Wb); log("host_reserve s", time.time() - t0, eng.stats()["host_store_numa_node"])
timing = {"alloc_s": t_alloc, "pin_s": eng.stats()["host_store_pin_seconds"], "numa": eng.stats()["host_store_numa_node"]}
def cycle(label, tier=L.FMA_TIER_HOST, reps=3):
    rows = []
    for r in range(reps):
        eng.sleep(["weights"], tier=tier); s1 = eng.stats()
        eng.wake(None); s2 = eng.stats()
        rows.append(dict(sleep_s=s1["sleep_seconds"], sleep_copy_s=s1["sleep_copy_seconds"], unmap_s=s1["sleep_unmap_seconds"],
                         wake_s=s2["wake_seconds"], wake_copy_s=s2["wake_copy_seconds"], map_s=s2["wake_map_seconds"],
                         first_copy=s2["wake_first_copy_delay"], k_s=s2["kernel_seconds"], k_n=s2["kernel_launches"],
                         d2h=Wb / s1["sleep_copy_seconds"] / 1e9, h2d=Wb / s2["wake_copy_seconds"] / 1e9, e2e_h2d=Wb / s2["wake_seconds"] / 1e9))
    after = eng.digest_all(["weights"])
    ok = after == before
    best = min(rows, key=lambda r: r["wake_s"])
    log(label, "ok", ok, {k: round(v, 4) for k, v in best.items()})
    timing[label] = dict(ok=ok, rows=rows)
for ns in (1, 2, 4):
    eng.set_option("mode", L.FMA_MODE_DIRECT); eng.set_option("copy_streams", ns)
    cycle(f"host_direct_s{ns}")
eng.set_option("copy_streams", 4)
for ch in (8, 128):
    eng.set_option("chunk_bytes", ch << 20); cycle(f"host_direct_chunk{ch}")
eng.set_option("chunk_bytes", 32 << 20)
for mt in (2, 4):
    eng.set_option("map_threads", mt); cycle(f"host_direct_map{mt}")
eng.set_option("map_threads", 1)
eng.set_option("mode", L.FMA_MODE_STAGED)
for kern, kname in [(L.FMA_KERNEL_TMA, "tma"), (L.FMA_KERNEL_LDG, "ldg")]:
    eng.set_option("kernel", kern); cycle(f"host_staged_{kname}")
eng.set_option("mode", L.FMA_MODE_KERNEL)
for kern, kname in [(L.FMA_KERNEL_TMA, "tma"), (L.FMA_KERNEL_LDG, "ldg")]:
    eng.set_option("kernel", kern); cycle(f"host_zerocopy_{kname}", reps=2)
for kern, kname in [(L.FMA_KERNEL_TMA, "tma"), (L.FMA_KERNEL_LDG, "ldg")]:
    eng.set_option("kernel", kern); cycle(f"local_kernel_{kname}", tier=L.FMA_TIER_LOCAL)
json.dump(timing, open(f"{OUT}/timing.json", "w"), indent=1)

# ---- K1 roofline sweep: whole-image gather, HBM -> HBM ------------------------------------------
wsegs = [s for s in eng.segments() if s.tag == "weights"]
pages = [s.va + o for s in wsegs for o in range(0, s.bytes, L.FMA_PAGE_BYTES)]
n = len(pages)
dst = eng.scratch_alloc(n * L.FMA_PAGE_BYTES)
sweep = []
for tile in (8, 16, 32, 64):
    for stages in (2, 3, 4, 6):
        for pipes in (1, 2, 4):
            for cps in (1, 2):
                smem = tile * 1024 * stages * pipes + 128
                if smem * cps > 227 * 1024: continue
                eng.set_option("tma_tile_bytes", tile << 10); eng.set_option("tma_stages", stages)
                eng.set_option("tma_pipes", pipes); eng.set_option("tma_ctas_per_sm", cps)
                ms = min(eng.op_page_copy(n, src_pages=pages, dst_base=dst, variant=L.FMA_KERNEL_TMA) for _ in range(3))
                gbs = 2 * n * L.FMA_PAGE_BYTES / ms / 1e6
                sweep.append(dict(tile_kib=tile, stages=stages, pipes=pipes, ctas_per_sm=cps, ms=ms, gbs=gbs))
ms = min(eng.op_page_copy(n, src_pages=pages, dst_base=dst, variant=L.FMA_KERNEL_LDG) for _ in range(3))
sweep.append(dict(variant="ldg", ms=ms, gbs=2 * n * L.FMA_PAGE_BYTES / ms / 1e6))
sweep.sort(key=lambda r: -r["gbs"])
for r in sweep[:8]: log(r)
log("ldg", [r for r in sweep if r.get("variant") == "ldg"])
dig, dms = eng.op_page_digest(n, pages=pages)
log("digest ms", dms, "GB/s", n * L.FMA_PAGE_BYTES / dms / 1e6)
json.dump(sweep, open(f"{OUT}/k1_sweep.json", "w"), indent=1)
eng.scratch_free(dst)
eng.close()
log("DONE")
