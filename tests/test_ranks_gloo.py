"""N>1 host logic on CPU: world_size-2 gloo group — barrier, max/sum over ranks, per-rank shard plan."""
import os
import socket

import pytest
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import fma_b200  # noqa: F401
    from fma_b200 import ranks
    from fma_b200 import workloads as W

    g = ranks.RankGroup(backend="gloo")
    g.barrier()
    t_wake = 0.30 + 0.05 * rank                       # rank 1 is the slow shard
    Wb = W.weight_bytes(W.allocation_table("llama-3-70b-tp8"))
    out = dict(rank=g.rank, world=g.world, max=g.max(t_wake), sum=g.sum(float(Wb)), all_ok=g.all_true(True),
               one_bad=g.all_true(rank != 1), seed=ranks.shard_seed(rank), park=ranks.parking_device(rank, world), W=Wb)
    out["gbs"] = ranks.aggregate_wake(out["sum"], out["max"])
    g.phase_barrier()
    out["max_vec"] = g.max_vec([0.1 * (rank + 1), 0.5 - 0.2 * rank])      # per-step max over ranks
    g.barrier()
    g.close()
    q.put(out)


def test_two_rank_gloo_aggregation():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted([q.get(timeout=120) for _ in ps], key=lambda d: d["rank"])
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r["rank"] for r in res] == [0, 1] and all(r["world"] == 2 for r in res)
    assert all(abs(r["max"] - 0.35) < 1e-12 for r in res)                     # max over ranks, same on every rank
    assert all(r["sum"] == 2 * res[0]["W"] for r in res)                      # whole-job bytes
    assert all(r["all_ok"] for r in res) and not any(r["one_bad"] for r in res)
    assert [r["seed"] for r in res] == [1234, 1235]
    assert sorted(r["park"] for r in res) == [0, 1] and all(r["park"] != r["rank"] for r in res)
    assert all(abs(r["gbs"] - 2 * res[0]["W"] / 0.35 / 1e9) < 1e-6 for r in res)
    assert all(r["max_vec"] == pytest.approx([0.2, 0.5]) for r in res)


def test_parking_is_a_fixed_point_free_permutation():
    import fma_b200  # noqa: F401
    from fma_b200 import ranks

    for n in (2, 4, 8):
        m = [ranks.parking_device(r, n) for r in range(n)]
        assert sorted(m) == list(range(n)) and all(m[r] != r for r in range(n))
    with pytest.raises(ValueError):
        ranks.parking_device(0, 1)


def test_single_process_group_is_a_noop():
    import fma_b200  # noqa: F401
    from fma_b200 import ranks

    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        os.environ.pop(k, None)
    g = ranks.RankGroup()
    g.barrier()
    assert g.world == 1 and g.max(1.5) == 1.5 and g.sum(2.0) == 2.0 and g.all_true(True)
