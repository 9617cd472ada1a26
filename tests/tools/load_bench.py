"""Cold load, file -> HBM: fma_load_file vs the loaders vLLM's default path builds on (safetensors' own GPU load and a
vLLM-style per-tensor safe_open + copy_), Llama-3-8B-shaped synthetic safetensors file in the page cache."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))  # repo root
import numpy as np, torch
import fma_b200
from fma_b200 import workloads as W, loader, _lib as L
from oracle import oracle as O
MODEL = sys.argv[1] if len(sys.argv) > 1 else "llama-3-8b"
path = f"/tmp/fma_{MODEL}.safetensors"
tensors = W.model_tensors(MODEL)
total = sum(b for _, b in tensors)
if not os.path.exists(path):
    t0 = time.time(); first = 0
    def gen():
        global first
        for n, b in tensors:
            raw = O.fill(b, 99, first); first += b // 8
            yield (n, "BF16", (b // 2,), raw.tobytes())
    loader.write_safetensors(path, gen())
    print("wrote", path, round(total / 2**30, 2), "GiB in", round(time.time() - t0, 1), "s", flush=True)
res = {"model": MODEL, "bytes": total}
# --- reference-style loaders -------------------------------------------------------------------------------
from safetensors import safe_open
from safetensors.torch import load_file
torch.cuda.init(); torch.cuda.synchronize()
for rep in range(2):
    t0 = time.perf_counter(); d = load_file(path, device="cuda:0"); torch.cuda.synchronize(); t = time.perf_counter() - t0
    res[f"safetensors_load_file_cuda_rep{rep}"] = dict(seconds=t, gbs=total / t / 1e9); del d; torch.cuda.empty_cache()
params = {n: torch.empty(b // 2, dtype=torch.bfloat16, device="cuda") for n, b in tensors}
for rep in range(2):
    t0 = time.perf_counter()
    with safe_open(path, framework="pt", device="cpu") as f:          # vLLM's safetensors_weights_iterator + param.copy_
        for n in f.keys():
            params[n].copy_(f.get_tensor(n))
    torch.cuda.synchronize(); t = time.perf_counter() - t0
    res[f"vllm_style_safe_open_copy_rep{rep}"] = dict(seconds=t, gbs=total / t / 1e9)
del params; torch.cuda.empty_cache()
# --- this engine -------------------------------------------------------------------------------------------
eng = fma_b200.Engine(0)
dst = {}
for s in W.simulate_segments(tensors, "weights"): pass
for n, b in tensors: dst[n] = eng.alloc(b, "weights")
for threads in (4, 8, 12, 16):
    for chunk in (8, 16, 32):
        eng.set_option("load_threads", threads); eng.set_option("load_chunk_bytes", chunk << 20); eng.set_option("load_slots", max(12, threads * 2))
        best = min((loader.load_safetensors(eng, path, dst) for _ in range(2)), key=lambda s: s["seconds"])
        res[f"fma_t{threads}_c{chunk}"] = best
        print(f"fma threads={threads} chunk={chunk}MiB: {best['seconds']:.3f} s {best['gbs']:.1f} GB/s (read {best['read_seconds']:.2f} thread-s)", flush=True)
# parity: K3 digest of every loaded tensor == oracle digest of the file bytes
first = 0; ok = True
dg = eng.digest_all(["weights"])
raw = np.memmap(path, dtype=np.uint8, mode="r")
ent = {t.name: t for t in loader.read_header(path)}
for i, (n, b) in enumerate(tensors[:40]):
    seg = eng.segment(i)
    if seg.bytes == b:
        ok = ok and dg[i] == O.digest(np.asarray(raw[ent[n].file_offset:ent[n].file_offset + b]), 0)
res["bit_exact_first_40_tensors"] = bool(ok)
for k, v in res.items():
    if isinstance(v, dict) and "gbs" in v and not k.startswith("fma_"): print(k, round(v["seconds"], 3), "s", round(v["gbs"], 1), "GB/s")
print("bit_exact", ok)
os.makedirs("gpurun_out/load", exist_ok=True); json.dump(res, open(f"gpurun_out/load/load_bench_{MODEL}.json", "w"), indent=1)
