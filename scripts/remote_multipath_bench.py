"""MULTI-PATH wake ACROSS PROCESSES at full size (BASELINE config 1 table): the owner process (this one; sees every GPU) runs the parking
service with its helper GPUs; the instance process sees ONLY GPU 0 (CUDA_VISIBLE_DEVICES=0, as under the launcher), attaches the
owner's staging buffers as remote paths and wakes a Llama-3-8B table: single link vs own link + k helper links driven by the owner.
Output: gpurun_out/remote_mp/remote_multipath.json"""
import json, os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

INSTANCE = r"""
import os, sys, json, statistics
sys.path.insert(0, %r)
import fma_b200
from fma_b200 import workloads as W, _lib as L
from fma_b200.parking import ParkingClient
k = int(sys.argv[1])
eng = fma_b200.Engine(0)
table = W.allocation_table("llama-3-8b", kv_cache_bytes=32 << 30)
for s in table: eng.alloc(s.bytes, s.tag)
first = 0
for i, s in enumerate(table):
    if s.tag == "weights":
        eng.fill(i, 1234, first); first += s.bytes // 8
Wb = W.weight_bytes(table)
before = eng.digest_all(["weights"])
cli = ParkingClient()
res = dict(visible=os.environ.get("CUDA_VISIBLE_DEVICES"), W=Wb, rows=[])
def cycles(label, pull):
    wakes = []
    for i in range(7):
        eng.sleep(["weights"])
        if pull: cli.request_pull(eng, "Ibench", 0, timeout_s=20.0)
        eng.wake(None); st = eng.stats()
        if i: wakes.append(st["wake_seconds"])
    tl = eng.timeline()
    chunks = {str(r["idx"]): round(r["bytes"] / 2**30, 2) for r in tl if r["kind"] == "path_chunks"}
    med = statistics.median(wakes)
    res["rows"].append(dict(label=label, wake_s_median=round(med, 5), wake_s_min_max=[round(min(wakes), 5), round(max(wakes), 5)],
                            e2e_gbs=round(Wb / med / 1e9, 1), gib_per_path=chunks))
eng.host_reserve(Wb)
cycles("single link (no paths)", False)
n = cli.attach_remote_paths(eng, "Ibench", 0, n_helpers=k, avoid=[0])
eng.host_reserve(Wb)
cycles("own link + %%d remote helper(s), pull requests served by the owner" %% n, True)
cycles("remote paths attached but no pull request (own link alone)", False)
res["bit_exact"] = eng.digest_all(["weights"]) == before
print(json.dumps(res), flush=True)
cli.release("Ibench")
eng.close()
""" % ROOT

def main():
    import torch
    import fma_b200
    from fma_b200.parking import ParkingService
    n = torch.cuda.device_count()
    k = min(int(sys.argv[1]) if len(sys.argv) > 1 else n - 1, n - 1)
    sock = os.path.join(tempfile.mkdtemp(), "agent.sock")
    svc = ParkingService(sock, n_devices=n)
    svc.start()
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="0", FMA_NODE_AGENT_SOCK=sock, FMA_HOST_STORE_SHM="1", FMA_PULL_TIMEOUT_S="20")
    script = os.path.join(os.path.dirname(sock), "instance.py")
    open(script, "w").write(INSTANCE)
    r = subprocess.run([sys.executable, script, str(k)], env=env, capture_output=True, text=True, timeout=900)
    svc.close()
    if r.returncode != 0:
        print(r.stdout[-2000:], r.stderr[-3000:])
        raise SystemExit(1)
    res = json.loads(r.stdout.strip().splitlines()[-1])
    res["owner_visible_gpus"] = n
    os.makedirs(os.path.join(ROOT, "gpurun_out", "remote_mp"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "remote_mp", "remote_multipath.json"), "w"), indent=1)
    for row in res["rows"]:
        print(json.dumps(row))
    print("bit_exact", res["bit_exact"], "instance saw GPUs:", res["visible"])

if __name__ == "__main__":
    main()
