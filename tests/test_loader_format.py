"""safetensors container handling on CPU: our writer/reader against the `safetensors` package (independent implementation),
span merging, and the library's argument checking (no GPU needed for those)."""
import numpy as np
import pytest

import fma_b200  # noqa: F401
from fma_b200 import loader


def _mk(tmp_path):
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, 4096, dtype=np.uint8)
    b = rng.standard_normal((8, 16)).astype(np.float32)
    c = rng.integers(-5, 5, (3, 5), dtype=np.int64)
    p = str(tmp_path / "m.safetensors")
    loader.write_safetensors(p, [("a", "U8", a.shape, a.tobytes()), ("b", "F32", b.shape, b.tobytes()), ("c", "I64", c.shape, c.tobytes())])
    return p, {"a": a, "b": b, "c": c}


def test_reader_matches_the_safetensors_package(tmp_path):
    st = pytest.importorskip("safetensors.numpy")
    p, ref = _mk(tmp_path)
    got = st.load_file(p)                                   # the independent implementation accepts our writer's file
    assert all(np.array_equal(got[k], ref[k]) for k in ref)
    q = str(tmp_path / "theirs.safetensors")
    st.save_file(ref, q)                                    # and our reader parses theirs
    raw = open(q, "rb").read()
    for t in loader.read_header(q):
        assert raw[t.file_offset:t.file_offset + t.nbytes] == ref[t.name].tobytes()
        assert t.shape == ref[t.name].shape


def test_spans_merge_only_when_file_and_device_ranges_are_both_adjacent(tmp_path):
    p, ref = _mk(tmp_path)
    ent = loader.read_header(p)
    by = {t.name: t for t in ent}
    base = 0x7000_0000_0000
    dst = {"a": base, "b": base + by["a"].nbytes, "c": base + (1 << 21)}
    spans = loader.spans_for(ent, dst)
    assert spans == [(by["a"].file_offset, by["a"].nbytes + by["b"].nbytes, base), (by["c"].file_offset, by["c"].nbytes, base + (1 << 21))]
    assert loader.spans_for(ent, {"b": base}) == [(by["b"].file_offset, by["b"].nbytes, base)]


def test_corrupt_header_is_rejected(tmp_path):
    p = str(tmp_path / "bad.safetensors")
    open(p, "wb").write((1 << 40).to_bytes(8, "little") + b"{}")
    with pytest.raises(ValueError):
        loader.read_header(p)
