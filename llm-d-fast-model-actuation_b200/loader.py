"""Cold load: safetensors file -> HBM through the engine's pinned-ring + copy-engine mover.

In the reference "load_model" is create-instance -> vLLM's checkpoint loader (inference_server/launcher/launcher.py:
656-669,799-837; vllm:v1/worker/gpu_worker.py:335-342): tensor by tensor, from pageable mmap'ed memory.  Here the file's
byte ranges are streamed straight to their device addresses by ``fma_load_file`` (reader threads + bounce ring + copy
engines); this module only parses the container.

safetensors layout (public format): 8-byte little-endian header length N, N bytes of JSON
``{name: {"dtype", "shape", "data_offsets": [begin, end]}, "__metadata__": {...}}``, then the data buffer; offsets are
relative to the start of the data buffer.
"""
from __future__ import annotations

import dataclasses
import json
import struct
from typing import Iterable, Mapping

DTYPE_BYTES = {"F64": 8, "F32": 4, "F16": 2, "BF16": 2, "I64": 8, "I32": 4, "I16": 2, "I8": 1, "U8": 1, "BOOL": 1,
               "F8_E4M3": 1, "F8_E5M2": 1}


@dataclasses.dataclass(frozen=True)
class TensorEntry:
    name: str
    dtype: str
    shape: tuple[int, ...]
    file_offset: int      # absolute offset of the tensor's bytes in the file
    nbytes: int


def read_header(path: str) -> list[TensorEntry]:
    with open(path, "rb") as f:
        (n,) = struct.unpack("<Q", f.read(8))
        if n > (100 << 20):
            raise ValueError(f"{path}: implausible safetensors header length {n}")
        header = json.loads(f.read(n))
    base = 8 + n
    out = []
    for name, meta in header.items():
        if name == "__metadata__":
            continue
        b, e = meta["data_offsets"]
        numel = 1
        for d in meta["shape"]:
            numel *= d
        if meta["dtype"] in DTYPE_BYTES and numel * DTYPE_BYTES[meta["dtype"]] != e - b:
            raise ValueError(f"{path}: tensor {name} shape/dtype disagree with its byte range")
        out.append(TensorEntry(name, meta["dtype"], tuple(meta["shape"]), base + b, e - b))
    out.sort(key=lambda t: t.file_offset)
    return out


def spans_for(entries: Iterable[TensorEntry], destinations: Mapping[str, int]) -> list[tuple[int, int, int]]:
    """(file_offset, nbytes, device_address) for every tensor that has a destination, merged when both the file
    ranges and the device ranges are adjacent (fewer, larger DMA transfers)."""
    spans: list[list[int]] = []
    for t in entries:
        if t.name not in destinations or t.nbytes == 0:
            continue
        dst = destinations[t.name]
        if spans and spans[-1][0] + spans[-1][1] == t.file_offset and spans[-1][2] + spans[-1][1] == dst:
            spans[-1][1] += t.nbytes
        else:
            spans.append([t.file_offset, t.nbytes, dst])
    return [tuple(s) for s in spans]


def load_safetensors(engine, path: str, destinations: Mapping[str, int], o_direct: bool = False) -> dict:
    """Stream every tensor of ``path`` named in ``destinations`` (name -> device address inside an engine segment,
    e.g. ``param.data_ptr()`` of tensors allocated in the engine's memory pool) into HBM.  Returns load statistics."""
    entries = read_header(path)
    missing = [n for n in destinations if n not in {t.name for t in entries}]
    if missing:
        raise KeyError(f"{path} has no tensor(s) {missing[:5]}")
    st = engine.load_file(path, spans_for(entries, destinations), o_direct=o_direct)
    st["gbs"] = st["bytes"] / st["seconds"] / 1e9 if st["seconds"] > 0 else 0.0
    return st


def write_safetensors(path: str, tensors: Iterable[tuple[str, str, tuple[int, ...], bytes | memoryview]], align: int = 8) -> None:
    """Minimal writer (tests / synthetic checkpoints): ``tensors`` yields (name, dtype, shape, raw little-endian bytes)."""
    items = list(tensors)
    header, off = {}, 0
    for name, dtype, shape, raw in items:
        header[name] = {"dtype": dtype, "shape": list(shape), "data_offsets": [off, off + len(raw)]}
        off += len(raw)
    blob = json.dumps(header, separators=(",", ":")).encode()
    blob += b" " * ((-(8 + len(blob))) % align)
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(blob)))
        f.write(blob)
        for _, _, _, raw in items:
            f.write(raw)
