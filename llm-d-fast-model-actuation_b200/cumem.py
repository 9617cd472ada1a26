"""Drop-in for ``vllm.device_allocator.cumem.CuMemAllocator`` (surface B2).

Same class name, singleton accessor, method names, argument meaning and error behaviour as the
reference's allocator shim, so vLLM's ``Worker.sleep`` / ``Worker.wake_up`` / ``load_model``
(vllm:v1/worker/gpu_worker.py:157-209,335-342) run unchanged on top of the B200 engine:

    reference (vllm:device_allocator/cumem.py)            this module
    ------------------------------------------            -----------------------------------
    CuMemAllocator.get_instance()          :118-128       CuMemAllocator.get_instance()
    .use_memory_pool(tag)                  :251-308       .use_memory_pool(tag)
    .sleep(offload_tags)                   :177-225       .sleep(offload_tags)   -> fma_sleep
    .wake_up(tags)                         :227-249       .wake_up(tags)         -> fma_wake
    .get_current_usage()                   :310-318       .get_current_usage()   -> fma_current_usage
    .pointer_to_data                       :131           .pointer_to_data (read-only view of the C table)
    my_malloc / my_free in cumem_allocator.abi3.so        my_malloc / my_free in libfma_b200.so

Differences that are the point of the rewrite: the registry lives in C (no Python callbacks from
inside the allocator), the host backup is ONE pre-pinned NUMA-local arena instead of a
``torch.empty(pin_memory=True)`` per segment inside the sleep loop, and copies are asynchronous
over several copy-engine streams overlapped with cuMemCreate/cuMemMap.

``install_into_vllm()`` swaps the class into an imported vLLM so the unmodified reference launcher
(inference_server/launcher/launcher.py:799-837) serves /sleep and /wake_up through this engine.
"""
from __future__ import annotations

import dataclasses
import gc
import json
import logging
import os
import threading
from contextlib import contextmanager
from typing import Any

from . import _lib as L
from .engine import Engine, EngineConfig

logger = logging.getLogger("vllm.fma_b200.cumem")  # child of vLLM's configured logger, so INFO lines reach the instance log

# py_device, py_alignedSize, py_d_mem, py_p_memHandle  (cumem.py:47-48); the 4th slot carries the
# engine's segment sequence number instead of a heap pointer to a CUmemGenericAllocationHandle.
HandleType = tuple[int, int, int, int]


@dataclasses.dataclass
class AllocationData:
    handle: HandleType
    tag: str
    cpu_backup_tensor: Any = None  # always None: the backup lives in the engine's packed host store
    mapped: bool = True
    has_backup: bool = False


def _tier_from_env() -> int:
    return {"host": L.FMA_TIER_HOST, "peer": L.FMA_TIER_PEER, "local": L.FMA_TIER_LOCAL}[
        os.environ.get("FMA_TIER", "host").lower()]


class CuMemAllocator:
    """Singleton per process, exactly like the reference (the C side keeps the current engine in a
    process global because torch's pluggable-allocator signature has no user pointer)."""

    instance: "CuMemAllocator | None" = None
    default_tag: str = "default"

    @staticmethod
    def get_instance() -> "CuMemAllocator":
        if CuMemAllocator.instance is None:
            CuMemAllocator.instance = CuMemAllocator()
        return CuMemAllocator.instance

    def __init__(self, device: int | None = None, config: EngineConfig | None = None, engine: Any = None):
        """``engine`` may be injected (tests of the host logic); normally the allocator creates its own."""
        if engine is not None:
            self.device = 0 if device is None else device
            self.engine = engine
        else:
            import torch

            if not torch.cuda.is_available():
                raise L.FmaError(L.FMA_ENODRIVER, "CuMemAllocator needs a CUDA device; there is no CPU fallback")
            self.device = torch.cuda.current_device() if device is None else device
            self.engine = Engine(self.device, config)
        self.engine.make_current()
        self.current_tag: str = CuMemAllocator.default_tag
        self.allocator_and_pools: dict[str, Any] = {}
        self._reserve_thread: threading.Thread | None = None
        # MULTI-PATH wake across processes (FMA_REMOTE_PATHS=k): the node agent's helper GPUs pull for this instance; that needs the
        # host store as a memfd the agent can map, so the switch is thrown before the store is first reserved
        self._remote = None
        if int(os.environ.get("FMA_REMOTE_PATHS", "0") or 0) > 0 and os.environ.get("FMA_NODE_AGENT_SOCK"):
            os.environ["FMA_HOST_STORE_SHM"] = "1"

    # ---- registry view -----------------------------------------------------------------
    @property
    def pointer_to_data(self) -> dict[int, AllocationData]:
        out: dict[int, AllocationData] = {}
        for s in self.engine.segments():
            out[s.va] = AllocationData((self.device, s.bytes, s.va, s.seq), s.tag, None, s.mapped, s.has_backup)
        return out

    def get_current_usage(self) -> int:
        return self.engine.current_usage()

    # ---- hot path ----------------------------------------------------------------------
    def sleep(self, offload_tags: tuple[str, ...] | str | None = None) -> None:
        if offload_tags is None:
            offload_tags = (CuMemAllocator.default_tag,)
        elif isinstance(offload_tags, str):
            offload_tags = (offload_tags,)
        assert isinstance(offload_tags, tuple)
        self._join_reserve()
        import torch

        tier = _tier_from_env()
        owner = None
        if tier == L.FMA_TIER_PEER:
            nbytes = sum(s.bytes for s in self.engine.segments() if s.tag in offload_tags)
            if os.environ.get("FMA_NODE_AGENT_SOCK"):
                # Under the launcher an instance sees only its own GPUs (launcher.py:171-187): the node agent owns the parking
                # buffer (parking.py), this rank attaches to it and deposits the image descriptor afterwards, so the parked
                # weights outlive this process.
                from .parking import ParkingClient

                owner = ParkingClient()
                try:
                    owner.park(self.engine, _instance_id(), _rank(), nbytes, avoid=_own_gpu_indices())
                except (RuntimeError, OSError, L.FmaError) as e:
                    # no peer has room (or the agent is gone): the sleep itself must not fail for that — the host tier always works
                    logger.warning("fma_b200: cannot park %.2f GiB on a peer GPU (%s): sleeping to the host tier instead", nbytes / 1024**3, e)
                    owner, tier = None, L.FMA_TIER_HOST
            else:
                # parking GPU for the NVLink tier: FMA_PEER_DEVICE (index among the devices this process sees)
                peer = int(os.environ.get("FMA_PEER_DEVICE", "-1"))
                if peer < 0:
                    raise L.FmaError(L.FMA_EINVAL, "FMA_TIER=peer needs FMA_NODE_AGENT_SOCK (node-level owner) or FMA_PEER_DEVICE=<visible device index of the parking GPU>")
                try:
                    self.engine.peer_reserve(peer, nbytes)
                except L.FmaError as e:
                    if e.code != L.FMA_ENOMEM:
                        raise
                    logger.warning("fma_b200: GPU %d cannot take %.2f GiB right now (%s): sleeping to the host tier instead", peer, nbytes / 1024**3, e.message)
                    tier = L.FMA_TIER_HOST
        self.engine.sleep(offload_tags, tier=tier)
        if owner is not None and self.engine.stats()["sleep_bytes_offloaded"]:
            owner.deposit(self.engine, _instance_id(), _rank(), tier)
        elif (tier == L.FMA_TIER_HOST and os.environ.get("FMA_DEPOSIT_HOST_IMAGE") == "1" and os.environ.get("FMA_NODE_AGENT_SOCK")
              and os.environ.get("FMA_HOST_STORE_SHM") == "1" and self.engine.stats()["sleep_bytes_offloaded"]):
            # HOST tier, opt-in: hand the memfd behind the store to the node agent so the image outlives this process.  The exported
            # store is read-only from then on: this engine's NEXT sleep takes (and pins) a fresh private store — right for an instance
            # that is about to be deleted, wrong for one that sleeps and wakes all day, hence not the default.
            from .parking import ParkingClient

            ParkingClient().deposit_host(self.engine, _instance_id(), _rank())
        st = self.engine.stats()
        total = st["sleep_bytes_offloaded"] + st["sleep_bytes_discarded"]
        # same INFO line the reference emits (cumem.py:215-222; parsed by llm-d-benchmark), plus GB/s
        logger.info(
            "CuMemAllocator: sleep freed %.2f GiB memory in total, of which %.2f GiB is backed up in CPU and the "
            "rest %.2f GiB is discarded directly.", total / 1024**3, st["sleep_bytes_offloaded"] / 1024**3,
            st["sleep_bytes_discarded"] / 1024**3)
        if st["sleep_copy_seconds"] > 0:
            logger.info("fma_b200: sleep %.3f s, D2H %.1f GB/s", st["sleep_seconds"],
                        st["sleep_bytes_offloaded"] / st["sleep_copy_seconds"] / 1e9)
        if st.get("image_packed"):
            logger.info("fma_b200: packed image: %.2f GiB stored for %.2f GiB of weights (%.3f)", st["image_store_bytes"] / 1024**3,
                        st["sleep_bytes_offloaded"] / 1024**3, st["image_store_bytes"] / max(st["sleep_bytes_offloaded"], 1))
        gc.collect()
        torch.cuda.empty_cache()

    def wake_up(self, tags: list[str] | None = None) -> None:
        moves_an_image = tags is None or any(t != "kv_cache" for t in tags)   # the kv cache is discarded on sleep: waking it copies nothing
        if self._remote is not None and moves_an_image and self.engine.is_sleeping() and _tier_from_env() == L.FMA_TIER_HOST:
            try:    # the owner's helpers serve the wake that follows; if the request fails, the instance's own link does it all
                self._remote.request_pull(self.engine, _instance_id(), _rank())
            except Exception as e:      # noqa: BLE001
                logger.warning("fma_b200: pull request to the node agent failed (%s): waking over the own link only", e)
        self.engine.wake(tags)
        st = self.engine.stats()
        if st["wake_copy_seconds"] > 0:
            logger.info("fma_b200: wake %.3f s, H2D %.1f GB/s", st["wake_seconds"],
                        st["wake_bytes_restored"] / st["wake_copy_seconds"] / 1e9)

    # ---- pool --------------------------------------------------------------------------
    @contextmanager
    def use_memory_pool(self, tag: str | None = None):
        import torch

        if tag is None:
            tag = CuMemAllocator.default_tag
        assert isinstance(tag, str)
        conf = os.environ.get("PYTORCH_CUDA_ALLOC_CONF", "")
        expandable_was_enabled = "expandable_segments:True" in conf  # incompatible with the pool (cumem.py:266-274)
        if expandable_was_enabled:
            torch.cuda.memory._set_allocator_settings("expandable_segments:False")
        old_tag = self.current_tag
        self.current_tag = tag
        self.engine.make_current()
        self.engine.set_current_tag(tag)
        try:
            new_alloc = torch.cuda.memory.CUDAPluggableAllocator(L.lib_path(), "my_malloc", "my_free")
            mem_pool = torch.cuda.memory.MemPool(new_alloc._allocator)
            data = (mem_pool, new_alloc)
            with torch.cuda.memory.use_mem_pool(mem_pool):
                self.allocator_and_pools[tag] = data  # keep alive (pytorch#146431, cumem.py:285-289)
                yield
                # release segments that ended up empty (e.g. load-time temporaries), cumem.py:299-303
                for allocation in mem_pool.snapshot():
                    if allocation["allocated_size"] == 0:
                        try:
                            self.engine.free(allocation["address"])
                        except L.FmaError:
                            pass
        finally:
            self.current_tag = old_tag
            self.engine.set_current_tag(old_tag)
            if expandable_was_enabled:
                torch.cuda.memory._set_allocator_settings("expandable_segments:True")
        if tag == "weights":
            # one line for tooling (scripts/e2e_launcher_vllm.py validates the synthetic tables of workloads.py against it)
            segs = [s for s in self.engine.segments() if s.tag == "weights"]
            hist: dict[int, int] = {}
            for sg in segs:
                hist[sg.bytes >> 20] = hist.get(sg.bytes >> 20, 0) + 1
            logger.info("fma_b200: weights pool closed: %s", json.dumps({"segments": len(segs), "bytes": sum(sg.bytes for sg in segs),
                                                                          "first_mib": [sg.bytes >> 20 for sg in segs[:6]],
                                                                          "mib_histogram": {str(k): v for k, v in sorted(hist.items())}}))
        if tag == "weights" and os.environ.get("FMA_ADOPT_PARKED") == "1" and os.environ.get("FMA_NODE_AGENT_SOCK"):
            self._adopt_parked_image()
        if tag == "weights" and self._remote is None and int(os.environ.get("FMA_REMOTE_PATHS", "0") or 0) > 0 \
                and os.environ.get("FMA_NODE_AGENT_SOCK") and _tier_from_env() == L.FMA_TIER_HOST:
            try:
                from .parking import ParkingClient

                cli = ParkingClient()
                n = cli.attach_remote_paths(self.engine, _instance_id(), _rank(), int(os.environ["FMA_REMOTE_PATHS"]), avoid=_own_gpu_indices())
                self._remote = cli
                logger.info("fma_b200: %d remote wake path(s) through the node agent's helper GPUs", n)
            except Exception as e:      # noqa: BLE001
                logger.warning("fma_b200: no remote wake paths (%s): host-tier wakes use the own link only", e)
        if tag == "weights" and os.environ.get("FMA_PREPIN", "1") != "0" and _tier_from_env() == L.FMA_TIER_HOST:
            self._start_reserve()

    def _adopt_parked_image(self) -> None:
        """Warm start: an earlier instance with this instance ID parked its weights with the node agent and died
        (or was deleted) asleep — the controller would cold-start here (inference-server.go:416-448).  The weights pool has just
        been filled by the loader (``--load-format dummy`` is enough: contents are about to be replaced); if the owner holds an image
        for (instance, rank) whose segment sequence matches, adopt it and wake: the weights come back over NVLink, digests
        verified.  Anything else (no image, another model) leaves the loaded weights as they are."""
        import time

        from .parking import ParkingClient

        t0 = time.perf_counter()
        try:
            cli = ParkingClient()
            if not cli.adopt(self.engine, _instance_id(), _rank(), tags=("weights",)) and \
                    not cli.adopt_host(self.engine, _instance_id(), _rank(), tags=("weights",)):
                logger.info("fma_b200: no parked image for instance %s rank %d: keeping the loaded weights", _instance_id(), _rank())
                return
            self.engine.wake(["weights"], flags=L.FMA_FLAG_VERIFY)
            st = self.engine.stats()
            logger.info("fma_b200: adopted the parked image of instance %s rank %d: %.2f GiB restored from %s in %.3f s (digests verified)",
                        _instance_id(), _rank(), st["wake_bytes_restored"] / 1024**3, "the host store" if st["tier"] == L.FMA_TIER_HOST else "peer HBM",
                        time.perf_counter() - t0)
        except L.FmaError as e:
            if self.engine.is_sleeping():
                raise                                   # half way: do not serve from unmapped weights
            logger.warning("fma_b200: parked image not adopted (%s): keeping the loaded weights", e)

    # ---- pre-pinning off the critical path -----------------------------------------------
    def _start_reserve(self) -> None:
        """Pin the host store for the weights as soon as they are loaded, in the background, so the
        first /sleep does not pay for it (the reference pins inside its sleep loop, cumem.py:204-209)."""
        nbytes = sum(s.bytes for s in self.engine.segments() if s.tag == "weights")
        if not nbytes:
            return

        def run():
            try:
                self.engine.host_reserve(nbytes)
            except L.FmaError as e:  # sleep will retry and report
                logger.warning("background host_reserve failed: %s", e)

        self._reserve_thread = threading.Thread(target=run, name="fma-prepin", daemon=True)
        self._reserve_thread.start()

    def _join_reserve(self) -> None:
        if self._reserve_thread is not None:
            self._reserve_thread.join()
            self._reserve_thread = None


def _instance_id() -> str:
    """The controller's instance ID (``"I" + base64url(sha256(...)) + "i"``, inference-server.go:801-829): the node agent passes it to
    its children as FMA_INSTANCE_ID; a bare process falls back to its pid."""
    return os.environ.get("FMA_INSTANCE_ID") or f"pid{os.getpid()}"


def _rank() -> int:
    return int(os.environ.get("RANK", os.environ.get("LOCAL_RANK", "0")) or 0)


def _own_gpu_indices() -> list[int]:
    """Node-level GPU indices this instance runs on (never park a model on its own GPU): FMA_NODE_GPU_INDICES, set by the node
    agent next to CUDA_VISIBLE_DEVICES."""
    v = os.environ.get("FMA_NODE_GPU_INDICES", "")
    return [int(x) for x in v.split(",") if x.strip().isdigit()]


def install_into_vllm() -> None:
    """Replace vLLM's allocator class with this one (call before the engine core starts workers)."""
    import vllm.device_allocator.cumem as ref

    ref.CuMemAllocator = CuMemAllocator  # type: ignore[misc]
    ref.cumem_available = True


__all__ = ["CuMemAllocator", "AllocationData", "HandleType", "install_into_vllm"]
