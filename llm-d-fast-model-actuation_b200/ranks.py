"""Multi-GPU plumbing for the path: one process per GPU, NO data-path collective.

Each tensor-parallel rank is its own process with its own engine and moves only its own shard — exactly how
vLLM fans ``("sleep", kwargs)`` / ``("wake_up", kwargs)`` out to its workers and waits for all of them
(vllm:v1/executor/multiproc_executor.py:339-379, abstract.py:327,347).  ``torch.distributed`` is used only for
the barrier around the timed region and for reducing the timings (max over ranks) — nccl on GPUs, gloo in
the CPU tests."""
from __future__ import annotations

import os


def rank_env() -> tuple[int, int, int]:
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


class RankGroup:
    def __init__(self, backend: str | None = None, device_index: int | None = None):
        self.rank, self.world, self.local_rank = rank_env()
        self.backend = backend
        self._dist = None
        self._cpu_group = None
        self._device = "cpu"
        if self.world > 1:
            import torch
            import torch.distributed as dist

            backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
            self.backend = backend
            kw = {}
            if backend == "nccl":
                idx = self.local_rank if device_index is None else device_index
                self._device = f"cuda:{idx}"
                kw["device_id"] = torch.device(self._device)
            if not dist.is_initialized():
                dist.init_process_group(backend, **kw)
            self._dist = dist
            # CPU-side group for the executor's "every rank has finished" hand-shake between the phases of a step
            # (vLLM waits for all workers over a shm message queue, multiproc_executor.py:339-379): no GPU work, no NCCL kernel
            self._cpu_group = dist.new_group(backend="gloo") if backend == "nccl" else None

    def barrier(self) -> None:
        if self._dist is not None:
            self._dist.barrier()

    def phase_barrier(self) -> None:
        """Executor semantics (abstract.py:327,347): ``sleep`` / ``wake_up`` are fanned out to every rank and return when ALL
        ranks are done, so no rank starts waking while another still sleeps.  Host-side only (gloo)."""
        if self._dist is not None:
            self._dist.barrier(group=self._cpu_group) if self._cpu_group is not None else self._dist.barrier()

    def max_vec(self, xs: list[float]) -> list[float]:
        """Element-wise max over ranks (per-step job latency = the slowest rank of that step)."""
        if self._dist is None:
            return [float(x) for x in xs]
        import torch

        t = torch.tensor([float(x) for x in xs], dtype=torch.float64, device=self._device)
        self._dist.all_reduce(t, op=self._dist.ReduceOp.MAX)
        return [float(v) for v in t.tolist()]

    def _reduce(self, x: float, op_name: str) -> float:
        if self._dist is None:
            return float(x)
        import torch

        t = torch.tensor([float(x)], dtype=torch.float64, device=self._device)
        self._dist.all_reduce(t, op=getattr(self._dist.ReduceOp, op_name))
        return float(t.item())

    def max(self, x: float) -> float:
        """Timings are reported as the max over ranks: the job is as slow as its slowest shard."""
        return self._reduce(x, "MAX")

    def sum(self, x: float) -> float:
        return self._reduce(x, "SUM")

    def all_true(self, ok: bool) -> bool:
        return self._reduce(0.0 if ok else 1.0, "SUM") == 0.0

    def close(self) -> None:
        if self._dist is not None and self._dist.is_initialized():
            self._dist.destroy_process_group()
            self._dist = None


def shard_seed(rank: int, base: int = 1234) -> int:
    """Synthetic content seed of a rank's shard (seed 1234 + rank)."""
    return base + rank


def parking_device(local_rank: int, world: int) -> int:
    """Peer-HBM tier placement used by the bench: rank r parks on GPU (r + max(1, N/2)) % N.  Through NVSwitch
    every peer is equally far, so placement is a capacity decision; this one is a fixed-point-free permutation,
    i.e. every GPU sends to exactly one peer and receives from exactly one."""
    if world < 2:
        raise ValueError("the peer tier needs at least 2 GPUs")
    return (local_rank + max(1, world // 2)) % world


def aggregate_wake(world_bytes: float, wake_seconds_max: float) -> float:
    """Whole-job GB/s = bytes all ranks restored / slowest rank's time."""
    return world_bytes / wake_seconds_max / 1e9
