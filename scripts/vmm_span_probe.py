"""VMM questions behind the arena design: (1) may one cuMemUnmap span several adjacent mappings? (2) may the handle be
released right after cuMemMap (memory lives until unmap)? (3) cost of ONE 16 GiB create+map+access vs 131 pieces,
alone and with a second process hammering VMM calls on another GPU (RM lock contention)."""
import time, sys, os, json, multiprocessing as mp
from cuda.bindings import driver as cu

def ck(r):
    err, rest = (r[0], r[1:]) if isinstance(r, tuple) else (r, ())
    if err != cu.CUresult.CUDA_SUCCESS: raise RuntimeError(str(err))
    return rest[0] if len(rest) == 1 else rest
MiB = 1 << 20
def setup(dev_idx):
    ck(cu.cuInit(0)); dev = ck(cu.cuDeviceGet(dev_idx)); ctx = ck(cu.cuDevicePrimaryCtxRetain(dev)); ck(cu.cuCtxSetCurrent(ctx))
    prop = cu.CUmemAllocationProp(); prop.type = cu.CUmemAllocationType.CU_MEM_ALLOCATION_TYPE_PINNED
    prop.location.type = cu.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE; prop.location.id = dev_idx
    acc = cu.CUmemAccessDesc(); acc.location.type = cu.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE; acc.location.id = dev_idx
    acc.flags = cu.CUmemAccess_flags.CU_MEM_ACCESS_FLAGS_PROT_READWRITE
    return prop, acc
def hammer(dev_idx, stop):
    prop, acc = setup(dev_idx)
    va = int(ck(cu.cuMemAddressReserve(64 * MiB, 2 * MiB, 0, 0))); n = 0
    while not stop.is_set():
        h = ck(cu.cuMemCreate(64 * MiB, prop, 0)); ck(cu.cuMemMap(va, 64 * MiB, 0, h, 0)); ck(cu.cuMemSetAccess(va, 64 * MiB, [acc], 1))
        ck(cu.cuMemUnmap(va, 64 * MiB)); ck(cu.cuMemRelease(h)); n += 1
def main():
    prop, acc = setup(0)
    free0 = ck(cu.cuMemGetInfo())[0]
    # (1)+(2)
    sizes = [64 * MiB, 32 * MiB, 128 * MiB]
    va = int(ck(cu.cuMemAddressReserve(sum(sizes), 2 * MiB, 0, 0))); off = 0
    for s in sizes:
        h = ck(cu.cuMemCreate(s, prop, 0)); ck(cu.cuMemMap(va + off, s, 0, h, 0)); ck(cu.cuMemSetAccess(va + off, s, [acc], 1)); ck(cu.cuMemRelease(h)); off += s
    ck(cu.cuMemsetD8(va, 1, sum(sizes))); ck(cu.cuCtxSynchronize())
    used = free0 - ck(cu.cuMemGetInfo())[0]
    try:
        ck(cu.cuMemUnmap(va, sum(sizes))); span = True
    except RuntimeError as e:
        span = str(e); off = 0
        for s in sizes: ck(cu.cuMemUnmap(va + off, s)); off += s
    after = free0 - ck(cu.cuMemGetInfo())[0]
    print("release-after-map keeps memory alive: used MiB", used // MiB, "| spanning unmap:", span, "| still used after unmap MiB", after // MiB, flush=True)
    # sub-range unmap of a merged mapping (expected to fail)
    h = ck(cu.cuMemCreate(sum(sizes), prop, 0)); ck(cu.cuMemMap(va, sum(sizes), 0, h, 0)); ck(cu.cuMemSetAccess(va, sum(sizes), [acc], 1)); ck(cu.cuMemRelease(h))
    try:
        ck(cu.cuMemUnmap(va, sizes[0])); sub = True
    except RuntimeError as e:
        sub = str(e)
    print("sub-range unmap of one mapping:", sub, flush=True)
    try: ck(cu.cuMemUnmap(va, sum(sizes)))
    except RuntimeError: pass
    # (3) cost: pieces vs merged, alone and under contention
    w8b = [1002 * MiB] + sum([[48 * MiB, 32 * MiB, 224 * MiB, 112 * MiB] for _ in range(32)], []) + [2 * MiB, 1002 * MiB]
    tot = sum(w8b); big = int(ck(cu.cuMemAddressReserve(tot, 2 * MiB, 0, 0)))
    def pieces():
        off = 0; t0 = time.perf_counter()
        for s in w8b:
            h = ck(cu.cuMemCreate(s, prop, 0)); ck(cu.cuMemMap(big + off, s, 0, h, 0)); ck(cu.cuMemSetAccess(big + off, s, [acc], 1)); ck(cu.cuMemRelease(h)); off += s
        t1 = time.perf_counter(); off = 0
        for s in w8b: ck(cu.cuMemUnmap(big + off, s)); off += s
        return (t1 - t0) * 1e3, (time.perf_counter() - t1) * 1e3
    def merged():
        t0 = time.perf_counter(); h = ck(cu.cuMemCreate(tot, prop, 0)); ck(cu.cuMemMap(big, tot, 0, h, 0)); ck(cu.cuMemSetAccess(big, tot, [acc], 1)); ck(cu.cuMemRelease(h))
        t1 = time.perf_counter(); ck(cu.cuMemUnmap(big, tot)); return (t1 - t0) * 1e3, (time.perf_counter() - t1) * 1e3
    res = {}
    for label in ("alone", "contended"):
        procs = []; stop = None
        if label == "contended":
            ctx = mp.get_context("spawn"); stop = ctx.Event()
            ndev = ck(cu.cuDeviceGetCount())
            for d in range(1, max(2, ndev)):
                p = ctx.Process(target=hammer, args=(d if d < ndev else 0, stop)); p.start(); procs.append(p)
            time.sleep(3)
        res[label] = dict(pieces=[pieces() for _ in range(4)], merged=[merged() for _ in range(4)])
        print(label, "pieces(map ms, unmap ms)", [(round(a, 1), round(b, 1)) for a, b in res[label]["pieces"]],
              "merged", [(round(a, 2), round(b, 2)) for a, b in res[label]["merged"]], flush=True)
        if stop is not None:
            stop.set(); [p.join() for p in procs]
    os.makedirs("gpurun_out/vmm", exist_ok=True); json.dump(res, open("gpurun_out/vmm/vmm_span_probe.json", "w"), indent=1)
if __name__ == "__main__":
    main()
