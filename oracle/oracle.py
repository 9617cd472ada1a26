"""ctypes + numpy front end of the C oracle (oracle/fma_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_DIR, "libfma_oracle.so")
PAGE = 2 << 20


def build() -> str:
    src = os.path.join(_DIR, "fma_oracle.c")
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _DIR, "libfma_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


class seg_t(C.Structure):
    _fields_ = [("dev", C.c_void_p), ("bytes", C.c_uint64), ("tag", C.c_int32), ("pad", C.c_int32),
                ("backup", C.c_void_p)]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        l = C.CDLL(build())
        l.fma_oracle_splitmix64.restype = C.c_uint64
        l.fma_oracle_splitmix64.argtypes = [C.c_uint64, C.c_uint64]
        l.fma_oracle_fill.restype = None
        l.fma_oracle_fill.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64]
        l.fma_oracle_digest.restype = C.c_uint64
        l.fma_oracle_digest.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64]
        l.fma_oracle_gather.restype = None
        l.fma_oracle_gather.argtypes = [C.POINTER(C.c_void_p), C.c_uint32, C.c_void_p]
        l.fma_oracle_scatter.restype = None
        l.fma_oracle_scatter.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_uint32]
        l.fma_oracle_sleep.restype = C.c_uint64
        l.fma_oracle_sleep.argtypes = [C.POINTER(seg_t), C.c_uint32, C.c_uint64]
        l.fma_oracle_wake.restype = C.c_uint64
        l.fma_oracle_wake.argtypes = [C.POINTER(seg_t), C.c_uint32, C.c_uint64, C.c_uint8]
        l.fma_oracle_pack_page.restype = C.c_uint32
        l.fma_oracle_pack_page.argtypes = [C.c_void_p, C.c_void_p]
        l.fma_oracle_unpack_page.restype = C.c_int
        l.fma_oracle_unpack_page.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        _lib = l
    return _lib


PACKED_PAGE = (3 << 19) + (16 << 10)   # stored size of a page in the "FMP4" code (csrc/fma_codec.h)


def pack_page(page: np.ndarray) -> np.ndarray:
    """Stored form of one 2 MiB page: PACKED_PAGE bytes in the FMP4 code, or the page verbatim (PAGE bytes)."""
    b = np.ascontiguousarray(page).view(np.uint8)
    assert b.size == PAGE
    out = np.empty(PAGE, dtype=np.uint8)
    n = int(lib().fma_oracle_pack_page(b.ctypes.data, out.ctypes.data))
    return out[:n].copy()


def unpack_page(stored: np.ndarray) -> np.ndarray:
    b = np.ascontiguousarray(stored).view(np.uint8)
    out = np.empty(PAGE, dtype=np.uint8)
    if lib().fma_oracle_unpack_page(b.ctypes.data, b.size, out.ctypes.data) != 0:
        raise ValueError("malformed stored page")
    return out


def bf16_weights(n_values: int, seed: int, scale: float = 1e-3) -> np.ndarray:
    """uint16 bit patterns of bf16 values ~ U(-scale, scale) (what vLLM's dummy loader fills parameters with,
    vllm:model_executor/model_loader/weight_utils.py:1451-1471), truncated from float32."""
    rng = np.random.default_rng(seed)
    f = rng.uniform(-scale, scale, n_values).astype(np.float32)
    return (f.view(np.uint32) >> 16).astype(np.uint16)


def splitmix64(seed: int, k: int) -> int:
    return int(lib().fma_oracle_splitmix64(seed, k))


def fill(nbytes: int, seed: int, first_word: int = 0) -> np.ndarray:
    """uint8 array of nbytes (multiple of 8) holding splitmix64(seed, first_word + j) little-endian."""
    assert nbytes % 8 == 0
    a = np.empty(nbytes // 8, dtype=np.uint64)
    lib().fma_oracle_fill(a.ctypes.data, a.size, seed, first_word)
    return a.view(np.uint8)


def digest(buf: np.ndarray, first_word: int = 0) -> int:
    b = np.ascontiguousarray(buf).view(np.uint8)
    return int(lib().fma_oracle_digest(b.ctypes.data, b.size, first_word))


def gather(pages: list[np.ndarray]) -> np.ndarray:
    """Packed image of a list of 2 MiB pages (K1)."""
    n = len(pages)
    ptrs = (C.c_void_p * n)(*[p.ctypes.data for p in pages])
    out = np.empty(n * PAGE, dtype=np.uint8)
    lib().fma_oracle_gather(ptrs, n, out.ctypes.data)
    return out


def scatter(image: np.ndarray, pages: list[np.ndarray]) -> None:
    n = len(pages)
    ptrs = (C.c_void_p * n)(*[p.ctypes.data for p in pages])
    lib().fma_oracle_scatter(image.ctypes.data, ptrs, n)


def packed_image(segments: list[np.ndarray]) -> np.ndarray:
    """Packed image of whole segments in table order (each a multiple of PAGE)."""
    pages = []
    for s in segments:
        assert s.size % PAGE == 0
        pages += [s[o:o + PAGE] for o in range(0, s.size, PAGE)]
    return gather(pages) if pages else np.empty(0, dtype=np.uint8)


class CuMemModel:
    """State-machine restatement of the reference allocator over numpy arrays
    (vllm:device_allocator/cumem.py:131-249).  ``dev[i] is None`` models an unmapped segment."""

    def __init__(self):
        self.sizes: list[int] = []
        self.tags: list[str] = []
        self.dev: list[np.ndarray | None] = []
        self.backup: list[np.ndarray | None] = []

    def malloc(self, nbytes: int, tag: str, data: np.ndarray | None = None) -> int:
        self.sizes.append(nbytes)
        self.tags.append(tag)
        self.dev.append(np.zeros(nbytes, dtype=np.uint8) if data is None else data.copy())
        self.backup.append(None)
        return len(self.sizes) - 1

    def sleep(self, offload_tags=("default",)) -> tuple[int, int]:
        if isinstance(offload_tags, str):
            offload_tags = (offload_tags,)
        backed = total = 0
        if any(d is None for d in self.dev):      # Executor.sleep guard, abstract.py:323-325
            return 0, 0
        for i in range(len(self.sizes)):            # cumem.py:198-213
            total += self.sizes[i]
            if self.tags[i] in offload_tags:
                self.backup[i] = self.dev[i].copy()
                backed += self.sizes[i]
            self.dev[i] = None
        return total, backed

    def wake_up(self, tags=None, poison: int = 0) -> int:
        restored = 0
        for i in range(len(self.sizes)):            # cumem.py:237-249
            if tags is None or self.tags[i] in tags:
                if self.dev[i] is not None:
                    continue
                if self.backup[i] is not None:
                    self.dev[i] = self.backup[i]
                    self.backup[i] = None
                    restored += self.sizes[i]
                else:
                    self.dev[i] = np.full(self.sizes[i], poison, dtype=np.uint8)
        return restored

    def is_sleeping(self) -> bool:
        return any(d is None for d in self.dev)

    def get_current_usage(self) -> int:
        return sum(self.sizes)
