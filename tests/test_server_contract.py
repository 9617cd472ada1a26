"""B1 wire contract on CPU: the three routes behave like the reference's executable spec (cmd/test-server/main.go:56-91)
and like vLLM's executor state machine (abstract.py:322-360).  A recording backend checks what reaches the worker level."""
import pytest
from fastapi.testclient import TestClient

import fma_b200  # noqa: F401
from fma_b200 import server


class Recorder:
    def __init__(self):
        self.calls = []

    def sleep(self, level):
        self.calls.append(("sleep", level))

    def wake_up(self, tags):
        self.calls.append(("wake_up", tags))


@pytest.fixture()
def client_and_backend():
    b = Recorder()
    return TestClient(server.create_app(b)), b


def test_controller_sequence(client_and_backend):
    """What ensureUnbound / querySleeping / wakeSleeper do (inference-server.go:1329-1339,1595-1607,1118-1137)."""
    c, b = client_and_backend
    assert c.get("/health").status_code == 200
    assert c.get("/is_sleeping").json() == {"is_sleeping": False}
    r = c.post("/sleep")                                   # the controller never sends level/mode: level 1
    assert r.status_code == 200 and r.content == b""
    assert c.get("/is_sleeping").json() == {"is_sleeping": True}
    r = c.post("/wake_up", headers={"Content-Type": "application/json"})   # doPost sends this content type, no body
    assert 200 <= r.status_code < 300
    assert c.get("/is_sleeping").json() == {"is_sleeping": False}
    assert b.calls == [("sleep", 1), ("wake_up", None)]


def test_idempotence_and_retries(client_and_backend):
    c, b = client_and_backend
    assert c.post("/wake_up").status_code == 200          # waking when awake: harmless, nothing reaches the workers
    assert b.calls == []
    c.post("/sleep"); c.post("/sleep")                     # sleeping twice: second is a no-op
    assert b.calls == [("sleep", 1)]
    c.post("/wake_up"); c.post("/wake_up"); c.post("/wake_up")   # the controller retries /wake_up (5 s timeout)
    assert b.calls == [("sleep", 1), ("wake_up", None)]


def test_level_and_tag_selective_wake(client_and_backend):
    c, b = client_and_backend
    c.post("/sleep?level=2&mode=abort")
    c.post("/wake_up?tags=weights")
    assert c.get("/is_sleeping").json() == {"is_sleeping": True}       # kv_cache still asleep
    c.post("/wake_up?tags=bogus")                                       # unknown tag: refused with a warning, still 200
    assert c.get("/is_sleeping").json() == {"is_sleeping": True}
    c.post("/wake_up?tags=weights")                                     # already woken tag: refused likewise
    c.post("/wake_up?tags=kv_cache")
    assert c.get("/is_sleeping").json() == {"is_sleeping": False}
    assert b.calls == [("sleep", 2), ("wake_up", ["weights"]), ("wake_up", ["kv_cache"])]


def test_startup_delay_like_test_server():
    c = TestClient(server.create_app(Recorder(), healthy_after=3600))
    assert c.get("/health").status_code == 503


def test_worker_policy_level1_offloads_weights_only():
    class FakeEngine:
        def __init__(self):
            self.calls = []

        def sleep(self, offload, tier=0):
            self.calls.append(("sleep", tuple(offload), tier))

        def wake(self, tags):
            self.calls.append(("wake", tags))

    e0, e1 = FakeEngine(), FakeEngine()
    be = server.EngineBackend([e0, e1], tier=0)
    be.sleep(1); be.wake_up(None); be.sleep(2); be.wake_up(["weights"])
    want = [("sleep", ("weights",), 0), ("wake", None), ("sleep", (), 0), ("wake", ["weights"])]
    assert e0.calls == want and e1.calls == want                       # every rank gets the same call (collective_rpc)


def test_cpu_worker_semantics_move_nothing():
    c = TestClient(server.create_app(server.CpuWorkerSemantics()))
    assert c.post("/sleep").status_code == 200 and c.get("/is_sleeping").json()["is_sleeping"] is True
    assert c.post("/wake_up").status_code == 200 and c.get("/is_sleeping").json()["is_sleeping"] is False


def test_a_retried_wake_up_never_overlaps_the_one_in_flight():
    """The controller retries POST /wake_up after its 5 s timeout (inference-server.go:1699-1716): the retry must wait for the
    call in flight (then find nothing to do), never enter the engine beside it — vLLM's engine-core RPC queue, go/fma/server.go
    and fma_served.cpp serialise the same way."""
    import threading
    import time

    from fma_b200.server import SleepState

    class SlowBackend:
        def __init__(self):
            self.inside = 0
            self.max_inside = 0
            self.wakes = 0
            self.lock = threading.Lock()

        def sleep(self, level):
            pass

        def wake_up(self, tags):
            with self.lock:
                self.inside += 1
                self.max_inside = max(self.max_inside, self.inside)
                self.wakes += 1
            time.sleep(0.2)
            with self.lock:
                self.inside -= 1

    b = SlowBackend()
    st = SleepState(b)
    st.sleep(1)
    ts = [threading.Thread(target=st.wake_up) for _ in range(4)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert b.max_inside == 1 and b.wakes == 1 and not st.is_sleeping
