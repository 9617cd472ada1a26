"""Generates tests/golden/fmp4_pages.json: SHA-256 of the STORED form of seeded 2 MiB pages under the "FMP4" page code
(csrc/fma_codec.h), produced by the oracle (oracle/fma_oracle.c).  The code has no counterpart in the reference — this
fixture pins the FORMAT: any later change to the oracle, the kernels or the header that alters a stored byte (or the
packed/raw decision) breaks tests/test_codec_oracle.py::test_format_is_pinned_by_the_golden_fixture.

    python tests/golden/make_fmp4_golden.py        # run from the repo root, CPU only
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

N = 1 << 20


def pages():
    """(name, page bytes): deterministic inputs built from numpy's PCG64 with fixed seeds and from closed forms."""
    yield "bf16_uniform_1e-3_seed1", O.bf16_weights(N, 1).view(np.uint8)
    g = np.random.default_rng(2).normal(0, 0.02, N).astype(np.float32)
    yield "bf16_normal_0.02_seed2", (g.view(np.uint32) >> 16).astype(np.uint16).view(np.uint8)
    yield "zeros", np.zeros(2 * N, np.uint8)
    v = np.full(N, 0x3F80, np.uint16); v[np.arange(2048) * 509 + 7] = 0x00D5
    yield "ones_with_2048_exceptions", v.view(np.uint8)
    v = np.full(N, 0x3F80, np.uint16); v[np.arange(2049) * 509 + 7] = 0x00D5
    yield "ones_with_2049_exceptions_raw", v.view(np.uint8)
    yield "ramp_u16", np.arange(N, dtype=np.uint32).astype(np.uint16).view(np.uint8)
    yield "noise_seed3", np.random.default_rng(3).integers(0, 256, 2 * N, dtype=np.uint8)


def main():
    out = {"format": "FMP4", "packed_bytes": O.PACKED_PAGE, "pages": []}
    for name, page in pages():
        stored = O.pack_page(page)
        assert np.array_equal(O.unpack_page(stored), page)
        out["pages"].append({"name": name, "input_sha256": hashlib.sha256(page.tobytes()).hexdigest(), "stored_bytes": int(stored.size),
                             "stored_sha256": hashlib.sha256(stored.tobytes()).hexdigest()})
    with open(os.path.join(ROOT, "tests", "golden", "fmp4_pages.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
