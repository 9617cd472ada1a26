"""Where does VMM time go on this box?  Times cuMemCreate / Map / SetAccess / Unmap / Release separately
(cuda-python driver bindings), for the segment sizes of the Llama-3-8B table + a 32 GiB kv region."""
import time, sys, json, os
from cuda.bindings import driver as cu
import numpy as np

def ck(r):
    if isinstance(r, tuple):
        err = r[0]; rest = r[1:]
    else:
        err = r; rest = ()
    if err != cu.CUresult.CUDA_SUCCESS:
        raise RuntimeError(str(err))
    return rest[0] if len(rest) == 1 else rest

ck(cu.cuInit(0)); dev = ck(cu.cuDeviceGet(0)); ctx = ck(cu.cuDevicePrimaryCtxRetain(dev)); ck(cu.cuCtxSetCurrent(ctx))
prop = cu.CUmemAllocationProp(); prop.type = cu.CUmemAllocationType.CU_MEM_ALLOCATION_TYPE_PINNED
prop.location.type = cu.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE; prop.location.id = 0
acc = cu.CUmemAccessDesc(); acc.location.type = cu.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE; acc.location.id = 0
acc.flags = cu.CUmemAccess_flags.CU_MEM_ACCESS_FLAGS_PROT_READWRITE
MiB = 1 << 20
def run(sizes, touch, label, reps=3):
    vas = [ck(cu.cuMemAddressReserve(s, 2 * MiB, 0, 0)) for s in sizes]
    out = []
    for r in range(reps):
        t = dict(create=0, map=0, access=0, unmap=0, release=0)
        hs = []
        for s, va in zip(sizes, vas):
            t0 = time.perf_counter(); h = ck(cu.cuMemCreate(s, prop, 0)); t1 = time.perf_counter()
            ck(cu.cuMemMap(va, s, 0, h, 0)); t2 = time.perf_counter()
            ck(cu.cuMemSetAccess(va, s, [acc], 1)); t3 = time.perf_counter()
            t["create"] += t1 - t0; t["map"] += t2 - t1; t["access"] += t3 - t2; hs.append(h)
        if touch:
            for s, va in zip(sizes, vas): ck(cu.cuMemsetD8(va, 0x5A, s))
            ck(cu.cuCtxSynchronize())
        for s, va, h in zip(sizes, vas, hs):
            t0 = time.perf_counter(); ck(cu.cuMemUnmap(va, s)); t1 = time.perf_counter()
            ck(cu.cuMemRelease(h)); t2 = time.perf_counter()
            t["unmap"] += t1 - t0; t["release"] += t2 - t1
        out.append({k: round(v * 1e3, 2) for k, v in t.items()})
    print(label, "n", len(sizes), "GiB", sum(sizes) / 2**30, "touch", touch, out, flush=True)
    for s, va in zip(sizes, vas): ck(cu.cuMemAddressFree(va, s))
    return out
w8b = [1002 * MiB] + sum([[48 * MiB, 32 * MiB, 224 * MiB, 112 * MiB] for _ in range(32)], []) + [2 * MiB, 1002 * MiB]
res = {}
for touch in (0, 1):
    res[f"w8b_touch{touch}"] = run(w8b, touch, "weights-8b", 4)
    res[f"kv32_touch{touch}"] = run([1024 * MiB] * 32, touch, "kv 32x1GiB", 4)
    res[f"big_touch{touch}"] = run([16 * 1024 * MiB], touch, "one 16 GiB", 3)
    res[f"small_touch{touch}"] = run([2 * MiB] * 1024, touch, "1024 x 2MiB", 3)
os.makedirs("gpurun_out/vmm", exist_ok=True); json.dump(res, open("gpurun_out/vmm/vmm_probe.json", "w"), indent=1)
