"""The oracle pinned against what exists to pin it with: the published splitmix64
known-answer vector, algebraic properties of the digest, gather/scatter inverses, and agreement
between the two independent restatements of cumem.py:177-249 (C over malloc, numpy state machine).
The vLLM-generated golden fixture is checked in test_golden.py."""
import ctypes as C

import numpy as np
import pytest

PAGE = 2 << 20


def test_splitmix64_known_answer(oracle):
    # Vigna's reference splitmix64.c, seed 1234567: first five outputs (published test vector)
    kat = [6457827717110365317, 3203168211198807973, 9817491932198370423, 4593380528125082431, 16408922859458223821]
    assert [oracle.splitmix64(1234567, k) for k in range(5)] == kat


def test_fill_is_counter_based(oracle):
    a = oracle.fill(1 << 16, 1234, 0).view(np.uint64)
    b = oracle.fill(1 << 15, 1234, (1 << 15) // 8).view(np.uint64)
    assert np.array_equal(a[(1 << 15) // 8:], b)           # any slice can be generated independently
    assert a[7] == oracle.splitmix64(1234, 7)
    assert not np.array_equal(a, oracle.fill(1 << 16, 1235, 0).view(np.uint64))


def _py_digest(buf, first_word=0):
    m = (1 << 64) - 1
    G = 0x9E3779B97F4A7C15
    acc = 0
    for j, w in enumerate(buf.view(np.uint64).tolist()):
        z = (w + (first_word + j + 1) * G) & m
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & m
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & m
        acc = (acc + (z ^ (z >> 31))) & m
    return acc


def test_digest_matches_pure_python_definition(oracle):
    buf = oracle.fill(4096, 99, 5)
    assert oracle.digest(buf, 17) == _py_digest(buf, 17)
    assert oracle.digest(np.zeros(0, dtype=np.uint8)) == 0


def test_digest_is_additive_and_position_sensitive(oracle):
    buf = oracle.fill(1 << 20, 7, 0)
    half = buf.size // 2
    whole = oracle.digest(buf, 0)
    parts = (oracle.digest(buf[:half], 0) + oracle.digest(buf[half:], half // 8)) & ((1 << 64) - 1)
    assert whole == parts                                   # checksum of checksums
    swapped = buf.copy()
    swapped[:8], swapped[8:16] = buf[8:16].copy(), buf[:8].copy()
    assert oracle.digest(swapped, 0) != whole               # misplacement is detected
    flipped = buf.copy()
    flipped[12345] ^= 1
    assert oracle.digest(flipped, 0) != whole


def test_gather_scatter_inverse(oracle):
    rng = np.random.default_rng(0)
    pages = [rng.integers(0, 256, PAGE, dtype=np.uint8) for _ in range(5)]
    order = [3, 0, 4, 1, 2]
    image = oracle.gather([pages[i] for i in order])
    for k, i in enumerate(order):
        assert np.array_equal(image[k * PAGE:(k + 1) * PAGE], pages[i])
    out = [np.zeros(PAGE, dtype=np.uint8) for _ in range(5)]
    oracle.scatter(image, [out[i] for i in order])
    assert all(np.array_equal(out[i], pages[i]) for i in range(5))
    assert oracle.packed_image([]).size == 0


def test_c_and_numpy_restatements_agree(oracle):
    """fma_oracle_sleep/wake (C, cumem.py:198-213,237-249) vs CuMemModel (numpy)."""
    libc = C.CDLL(None)
    libc.malloc.restype = C.c_void_p
    libc.malloc.argtypes = [C.c_size_t]
    sizes = [PAGE, 3 * PAGE, PAGE, 2 * PAGE]
    tags = [0, 1, 0, 1]  # 0 = weights, 1 = kv_cache
    rng = np.random.default_rng(1)
    data = [rng.integers(0, 256, s, dtype=np.uint8) for s in sizes]
    segs = (oracle.seg_t * 4)()
    model = oracle.CuMemModel()
    for i in range(4):
        p = libc.malloc(sizes[i])
        C.memmove(p, data[i].ctypes.data, sizes[i])
        segs[i].dev, segs[i].bytes, segs[i].tag, segs[i].backup = p, sizes[i], tags[i], None
        model.malloc(sizes[i], "weights" if tags[i] == 0 else "kv_cache", data[i])
    backed = oracle.lib().fma_oracle_sleep(segs, 4, 1 << 0)
    total, backed_m = model.sleep(("weights",))
    assert backed == backed_m == sizes[0] + sizes[2] and total == sum(sizes)
    assert all(not segs[i].dev for i in range(4)) and model.is_sleeping()
    assert oracle.lib().fma_oracle_sleep(segs, 4, 1 << 0) == 0 and model.sleep(("weights",)) == (0, 0)  # idempotent
    restored = oracle.lib().fma_oracle_wake(segs, 4, 0, 0xAB)
    assert restored == model.wake_up(None, poison=0xAB) == backed
    for i in range(4):
        got = np.ctypeslib.as_array((C.c_uint8 * sizes[i]).from_address(segs[i].dev))
        assert np.array_equal(got, model.dev[i])
        if tags[i] == 0:
            assert np.array_equal(got, data[i])              # offloaded tags come back bit-identical
        else:
            assert (got == 0xAB).all()                       # discarded tags come back without their contents
    assert not model.is_sleeping()
    assert oracle.lib().fma_oracle_wake(segs, 4, 0, 0) == 0  # waking when awake is harmless


def test_tag_selective_wake(oracle):
    model = oracle.CuMemModel()
    a = model.malloc(PAGE, "weights", np.full(PAGE, 1, np.uint8))
    b = model.malloc(PAGE, "kv_cache", np.full(PAGE, 2, np.uint8))
    model.sleep(("weights",))
    model.wake_up(["weights"])
    assert model.dev[a] is not None and model.dev[b] is None and model.is_sleeping()
    assert model.sleep(("weights",)) == (0, 0)               # still "sleeping": Executor guard
    model.wake_up(["kv_cache"])
    assert not model.is_sleeping() and (model.dev[a] == 1).all()
