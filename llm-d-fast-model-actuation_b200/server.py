"""The three dev-mode routes the dual-pods controller calls on an inference server, backed by the engine (B1).

Wire contract, restated from the reference's executable spec ``cmd/test-server/main.go:56-91`` and
vLLM's router ``vllm:entrypoints/serve/sleep/api_router.py:22-49``:

    POST /sleep[?level=1][&mode=abort]      -> 200, empty body, only after every rank finished
    POST /wake_up[?tags=weights&tags=...]   -> 200  (no tags = wake everything); safe to retry
    GET  /is_sleeping                       -> 200 {"is_sleeping": bool}
    GET  /health                            -> 200 once serving

and the executor-level state machine ``vllm:v1/executor/abstract.py:322-360`` (sleeping twice / waking when awake are
harmless; ``sleeping_tags = {"weights","kv_cache"}``; waking an unknown tag is refused with a warning, not an error).

With a real vLLM the routes are vLLM's own and this module is not used (the allocator is swapped underneath, see
``cumem.install_into_vllm``).  It serves (a) non-vLLM engines that adopt the C-ABI directly and (b) BASELINE config 0:
the unmodified reference launcher driving a child without a GPU, where — exactly like vLLM's CPU worker
(``vllm:v1/worker/cpu_worker.py:148-154``) — nothing moves and only the state flips (``CpuWorkerSemantics``; this is
plumbing, not a data-path fallback: it owns no weights).
"""
from __future__ import annotations

import logging
import threading
import time
from typing import Iterable, Protocol

from fastapi import FastAPI, Request
from fastapi.responses import JSONResponse, PlainTextResponse, Response

logger = logging.getLogger("fma_b200.server")

SLEEPING_TAGS = ("weights", "kv_cache")   # abstract.py:329


class Backend(Protocol):
    def sleep(self, level: int) -> None: ...
    def wake_up(self, tags: list[str] | None) -> None: ...


class EngineBackend:
    """Worker-level policy of ``Worker.sleep/wake_up`` (vllm:v1/worker/gpu_worker.py:157-196) over one or more engines
    (one per tensor-parallel rank hosted by this process)."""

    def __init__(self, engines: Iterable, tier: int = 0):
        self.engines = list(engines)
        self.tier = tier

    def sleep(self, level: int) -> None:
        offload = ("weights",) if level == 1 else tuple()    # gpu_worker.py:169-170
        for e in self.engines:
            e.sleep(offload, tier=self.tier)

    def wake_up(self, tags: list[str] | None) -> None:
        for e in self.engines:
            e.wake(tags)


class CpuWorkerSemantics:
    """vLLM's CPU worker: sleep and wake_up are no-ops with a warning (cpu_worker.py:148-154)."""

    def sleep(self, level: int) -> None:
        logger.warning("sleep mode is not supported on CPU, ignore it.")

    def wake_up(self, tags: list[str] | None) -> None:
        logger.warning("sleep mode is not supported on CPU, ignore it.")


class SleepState:
    """Executor.sleep / wake_up / is_sleeping (abstract.py:322-360), byte for byte in behaviour."""

    def __init__(self, backend: Backend):
        self.backend = backend
        self.is_sleeping = False
        self.sleeping_tags: set[str] = set()
        self.last_sleep_seconds = 0.0
        self.last_wake_seconds = 0.0
        # The routes are sync endpoints (thread pool): a controller retry after its 5 s /wake_up timeout
        # (inference-server.go:1699-1716) can overlap the call still in flight.  vLLM serialises these through the engine
        # core's RPC queue, go/fma/server.go and fma_served.cpp with a mutex; so does this.
        self._lock = threading.Lock()

    def sleep(self, level: int = 1) -> None:
        with self._lock:
            self._sleep(level)

    def wake_up(self, tags: list[str] | None = None) -> None:
        with self._lock:
            self._wake_up(tags)

    def _sleep(self, level: int = 1) -> None:
        if self.is_sleeping:
            logger.warning("Executor is already sleeping.")
            return
        t0 = time.perf_counter()
        self.backend.sleep(level)
        self.last_sleep_seconds = time.perf_counter() - t0
        self.sleeping_tags = set(SLEEPING_TAGS)
        self.is_sleeping = True
        logger.info("It took %.6f seconds to fall asleep.", self.last_sleep_seconds)

    def _wake_up(self, tags: list[str] | None = None) -> None:
        if not self.is_sleeping:
            logger.warning("Executor is not sleeping.")
            return
        if tags:
            for tag in tags:
                if tag not in self.sleeping_tags:
                    logger.warning("Tag %s is not in sleeping tags %s", tag, self.sleeping_tags)
                    return
        t0 = time.perf_counter()
        self.backend.wake_up(tags)
        self.last_wake_seconds = time.perf_counter() - t0
        logger.info("It took %.6f seconds to wake up tags %s.", self.last_wake_seconds,
                    tags if tags is not None else self.sleeping_tags)
        if tags:
            for tag in tags:
                self.sleeping_tags.remove(tag)
        else:
            self.sleeping_tags.clear()
        if not self.sleeping_tags:
            self.is_sleeping = False


def create_app(backend: Backend, healthy_after: float = 0.0):
    """FastAPI app with the four routes.  ``healthy_after`` mimics test-server's --startup-delay (main.go:37,46)."""
    app = FastAPI()
    state = SleepState(backend)
    app.state.sleep_state = state
    ready_at = time.time() + healthy_after

    @app.get("/health")
    async def health():
        if time.time() >= ready_at:
            return PlainTextResponse("OK\n", status_code=200)
        return PlainTextResponse("Service Unavailable\n", status_code=503)

    @app.post("/sleep")
    def sleep(raw_request: Request):
        level = raw_request.query_params.get("level", "1")          # api_router.py:25
        raw_request.query_params.get("mode", "abort")               # accepted, scheduling is the engine core's business
        state.sleep(int(level))
        return Response(status_code=200)

    @app.post("/wake_up")
    def wake_up(raw_request: Request):
        tags = raw_request.query_params.getlist("tags")
        state.wake_up(tags if tags else None)                       # [] -> None: wake everything (api_router.py:36-38)
        return Response(status_code=200)

    @app.get("/is_sleeping")
    async def is_sleeping():
        return JSONResponse(content={"is_sleeping": state.is_sleeping})   # pkg/api/interface.go:129-133

    return app


async def run_server(args) -> None:
    """Same calling convention as ``vllm.entrypoints.openai.api_server.run_server(args)``
    (inference_server/launcher/launcher.py:38,837): lets the unmodified launcher fork a child that serves the sleep
    routes with CPU-worker semantics on ``args.port`` (BASELINE config 0 staging)."""
    import uvicorn

    app = create_app(CpuWorkerSemantics())
    config = uvicorn.Config(app, host=getattr(args, "host", None) or "0.0.0.0", port=int(getattr(args, "port", 8000)), log_level="info")
    await uvicorn.Server(config).serve()
