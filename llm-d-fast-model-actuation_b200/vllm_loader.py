"""vLLM model loader ``--load-format fma``: checkpoint files reach HBM through the engine's cold-load mover
(``fma_load_file``: reader threads -> pinned bounce ring -> copy engines) instead of vLLM's tensor-by-tensor copies from
pageable, mmap'ed memory.

In the reference "load_model" is: the launcher forks vLLM (inference_server/launcher/launcher.py:799-837), the worker
calls ``get_model_loader(load_config).load_model`` (vllm:v1/worker/gpu_worker.py:335-342), and
``DefaultModelLoader`` feeds ``model.load_weights`` from ``safetensors_weights_iterator``
(vllm:model_executor/model_loader/default_loader.py:211-286).  This loader keeps ALL of that — file discovery, name
mapping, tensor-parallel sharding and qkv / gate_up fusion inside each parameter's ``weight_loader`` — and replaces only
the iterator: every safetensors file is streamed window by window (<= 4 GiB) into a device staging buffer, and the
iterator yields *device* tensors that alias it, so the copies vLLM then makes are HBM -> HBM.

The staging buffer belongs to a PRIVATE engine (its own VA arena): nothing of it lands in the ``weights`` pool, so the
model's weights stay one VA-contiguous run and the buffer disappears with the loader.

Selected per instance with ``--load-format fma`` in ``InferenceServerConfig.spec.modelServerConfig.options``
(api/fma/v1alpha1/inferenceserverconfig_types.go:42-44); registered by the vLLM plugin (plugin/fma_b200_vllm_plugin)
when ``FMA_B200=1``.  Not yet run on a GPU: the window / aliasing logic is tested against the host-simulated engine
(tests/test_engine_hostsim.py), the vLLM registration against the installed vLLM (tests/test_vllm_loader.py).
"""
from __future__ import annotations

import os

from typing import Callable, Iterable, Iterator

from . import loader as _fmt

WINDOW_BYTES = 4 << 30
_ALIGN = 256

# safetensors dtype tag -> (torch dtype name, bytes per element)
TORCH_DTYPES = {"F64": "float64", "F32": "float32", "F16": "float16", "BF16": "bfloat16", "I64": "int64", "I32": "int32",
                "I16": "int16", "I8": "int8", "U8": "uint8", "BOOL": "bool", "F8_E4M3": "float8_e4m3fn", "F8_E5M2": "float8_e5m2"}


def plan_windows(entries: list[_fmt.TensorEntry], window_bytes: int = WINDOW_BYTES) -> list[tuple[int, int, list[_fmt.TensorEntry]]]:
    """Cut a file's tensors (sorted by offset) into windows: (file offset of the window, bytes, tensors inside).
    A window starts on a 256-byte boundary of the file so that a tensor keeps its file alignment inside the staging buffer;
    a tensor larger than ``window_bytes`` gets a window of its own."""
    out: list[tuple[int, int, list[_fmt.TensorEntry]]] = []
    cur: list[_fmt.TensorEntry] = []
    start = end = 0
    for t in entries:
        if t.nbytes == 0:
            continue
        if cur and t.file_offset + t.nbytes - start > window_bytes:
            out.append((start, end - start, cur))
            cur = []
        if not cur:
            start = end = (t.file_offset // _ALIGN) * _ALIGN
        cur.append(t)
        end = max(end, t.file_offset + t.nbytes)
    if cur:
        out.append((start, end - start, cur))
    return out


def stream_tensors(engine, files: Iterable[str], view: Callable[[int, int, _fmt.TensorEntry], object],
                   drain: Callable[[], None] = lambda: None, window_bytes: int = WINDOW_BYTES,
                   stats: dict | None = None) -> Iterator[tuple[str, object]]:
    """The iterator's core, free of torch: for each file and window, (1) ``drain()`` — wait until the consumer's copies
    out of the staging buffer have finished, (2) stream the window's byte range into the staging segment with
    ``fma_load_file``, (3) yield ``(name, view(device_address, offset_in_window, entry))`` for every tensor in it.
    ``engine`` is the private staging engine; its single segment grows to the largest window seen."""
    seg_ptr, seg_bytes = 0, 0
    try:
        for path in files:
            entries = _fmt.read_header(path)
            for start, nbytes, tensors in plan_windows(entries, window_bytes):
                drain()
                if nbytes > seg_bytes:
                    if seg_ptr:
                        engine.free(seg_ptr)
                    seg_ptr, seg_bytes = engine.alloc(nbytes, "staging"), nbytes
                st = engine.load_file(path, [(start, nbytes, seg_ptr)])
                if stats is not None:
                    stats["bytes"] = stats.get("bytes", 0) + st["bytes"]
                    stats["seconds"] = stats.get("seconds", 0.0) + st["seconds"]
                    stats["windows"] = stats.get("windows", 0) + 1
                for t in tensors:
                    yield t.name, view(seg_ptr, t.file_offset - start, t)
        drain()
    finally:
        if seg_ptr:
            engine.free(seg_ptr)


class _CudaArray:
    """``__cuda_array_interface__`` holder: lets torch alias engine-owned device memory without copying."""

    def __init__(self, ptr: int, nbytes: int, owner):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}
        self._owner = owner


def torch_view(torch, owner, alias: bool | None = None):
    """view(device_address, offset, entry) -> torch tensor of the entry's dtype and shape.

    Default: a private device copy of the staging bytes (one HBM-speed D2D copy per tensor).  The staging segment is
    overwritten by the next window and freed at the end, and vLLM weight loaders may keep ``loaded_weight`` past their
    iteration (fused / MoE expert stacking, deferred quantisation, copies on another stream), so handing out aliases is only
    safe for loaders known to consume each tensor at once: ``FMA_LOADER_ALIAS=1`` (or ``alias=True``) turns the zero-copy
    aliasing on (a misaligned tensor is cloned either way)."""
    if alias is None:
        alias = os.environ.get("FMA_LOADER_ALIAS") == "1"

    def view(ptr: int, off: int, t: _fmt.TensorEntry):
        raw = torch.as_tensor(_CudaArray(ptr + off, t.nbytes, owner), device="cuda")
        dtype = getattr(torch, TORCH_DTYPES[t.dtype])
        if not alias or (ptr + off) % max(_fmt.DTYPE_BYTES[t.dtype], 1):
            raw = raw.clone()
        return raw.view(dtype).reshape(t.shape)
    return view


def register() -> None:
    """Register ``--load-format fma`` with vLLM (called from the general plugin)."""
    import torch
    from vllm.model_executor.model_loader import register_model_loader
    from vllm.model_executor.model_loader.default_loader import DefaultModelLoader

    @register_model_loader("fma")
    class FmaModelLoader(DefaultModelLoader):
        """DefaultModelLoader with the safetensors iterator replaced by the engine's file -> HBM stream."""

        def _get_weights_iterator(self, source):
            hf_folder, files, use_safetensors = self._prepare_weights(
                source.model_or_path, source.subfolder, source.revision, source.fall_back_to_pt, source.allow_patterns_overrides)
            if not use_safetensors:                                   # *.bin / *.pt checkpoints: vLLM's own path
                return super()._get_weights_iterator(source)
            from . import Engine

            eng = Engine(torch.cuda.current_device())
            stats: dict = {}
            self.fma_load_stats = stats

            def gen():
                try:
                    for name, tensor in stream_tensors(eng, sorted(files), torch_view(torch, eng),
                                                       drain=torch.cuda.current_stream().synchronize, stats=stats):
                        yield source.prefix + name, tensor
                finally:
                    eng.close()
            return gen()

    return FmaModelLoader
