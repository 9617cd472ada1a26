"""The compiled host side end to end on a GPU box: fma_served (csrc/fma_served.cpp: the reference's cmd/test-server with engines
behind it) holding a Llama-3-8B-shaped table per rank; POST /sleep, GET /is_sleeping, POST /wake_up timed by wall clock over
HTTP exactly as the dual-pods controller issues them (inference-server.go:1329-1339,1595-1607,1118-1137), K3 digests compared
before / after.  usage: python scripts/native_server_e2e.py [workload] [n_ranks] [pack]   -> gpurun_out/e2e/native_server_*.json"""
import json, os, subprocess, sys, time, urllib.request

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fma_b200  # noqa: E402,F401
from fma_b200 import workloads as W  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "llama-3-8b"
ranks = int(sys.argv[2]) if len(sys.argv) > 2 else 1
pack = int(sys.argv[3]) if len(sys.argv) > 3 else 0
table = W.allocation_table(workload, kv_cache_bytes=32 << 30)
args = [os.path.join(ROOT, "llm-d-fast-model-actuation_b200", "fma_served"), "--port", "0", "--pack", str(pack)]
for r in range(ranks):
    args += ["--device", str(r)]
for s in table:
    args += ["--seg", f"{s.tag}:{s.bytes >> 20}"]
p = subprocess.Popen(args, stdout=subprocess.PIPE, text=True)
out = {"workload": workload, "ranks": ranks, "pack": pack, "weights_gib_per_rank": W.weight_bytes(table) / 2**30, "rows": []}
try:
    line = p.stdout.readline()
    assert line.startswith("listening on "), line
    base = f"http://127.0.0.1:{int(line.split()[-1])}"

    def call(method, path):
        req = urllib.request.Request(base + path, data=b"" if method == "POST" else None, method=method)
        t0 = time.perf_counter()
        with urllib.request.urlopen(req, timeout=600) as r:
            return r.status, r.read().decode(), time.perf_counter() - t0

    before = call("GET", "/digests")[1]
    for rep in range(6):
        s_st, _, s_t = call("POST", "/sleep")
        asleep = json.loads(call("GET", "/is_sleeping")[1])["is_sleeping"]
        w_st, _, w_t = call("POST", "/wake_up")
        awake = not json.loads(call("GET", "/is_sleeping")[1])["is_sleeping"]
        st = json.loads(call("GET", "/stats")[1])
        out["rows"].append({"sleep_status": s_st, "sleep_s": round(s_t, 4), "wake_status": w_st, "wake_s": round(w_t, 4), "asleep": asleep, "awake": awake,
                            "engine_wake_s": [round(r["wake_seconds"], 4) for r in st["ranks"]], "image_store_gib": [round(r["image_store_bytes"] / 2**30, 3) for r in st["ranks"]]})
        print(out["rows"][-1], flush=True)
    out["bit_exact"] = call("GET", "/digests")[1] == before
    wakes = sorted(r["wake_s"] for r in out["rows"][1:])
    out["wake_s_median_http"] = wakes[len(wakes) // 2]
    out["wake_gbs_aggregate_http"] = round(ranks * W.weight_bytes(table) / out["wake_s_median_http"] / 1e9, 2)
finally:
    p.terminate(); p.wait(timeout=60)
    os.makedirs(os.path.join(ROOT, "gpurun_out", "e2e"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "e2e", f"native_server_{workload}_r{ranks}_pack{pack}.json"), "w"), indent=1)
    print(json.dumps({k: v for k, v in out.items() if k != "rows"}))
