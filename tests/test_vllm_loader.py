"""--load-format fma (llm-d-fast-model-actuation_b200/vllm_loader.py): window planning on CPU, and registration with the
installed vLLM's loader registry.  The streaming core runs against the host-simulated engine in
tests/test_engine_hostsim.py; a GPU run with a real vLLM server is pending (scripts/e2e_launcher_vllm.py)."""
import importlib
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VL = importlib.import_module("llm-d-fast-model-actuation_b200.vllm_loader")
FMT = importlib.import_module("llm-d-fast-model-actuation_b200.loader")


def _entries(sizes, base=1000):
    out, off = [], base
    for i, n in enumerate(sizes):
        out.append(FMT.TensorEntry(f"t{i}", "U8", (n,), off, n))
        off += n
    return out


def test_windows_cover_every_tensor_once_and_respect_the_limit():
    e = _entries([300, 0, 700, 5000, 100, 100, 2500], base=1000)
    w = VL.plan_windows(e, window_bytes=2048)
    names = [t.name for _, _, ts in w for t in ts]
    assert names == ["t0", "t2", "t3", "t4", "t5", "t6"]                      # the empty tensor is skipped
    for start, n, ts in w:
        assert start % 256 == 0 and start <= ts[0].file_offset
        assert start + n == ts[-1].file_offset + ts[-1].nbytes
        assert n <= 2048 or len(ts) == 1                                     # only a lone oversized tensor exceeds the window
    assert [len(ts) for _, _, ts in w] == [2, 1, 2, 1]
    one = VL.plan_windows(e, window_bytes=1 << 30)
    assert len(one) == 1 and one[0][0] == 768 and one[0][1] == 1000 + sum(t.nbytes for t in e) - 768


def test_dtype_table_covers_the_container_format():
    assert set(VL.TORCH_DTYPES) == set(FMT.DTYPE_BYTES)


def test_registers_with_the_installed_vllm():
    code = ("import sys; sys.path.insert(0, %r)\n"
            "import fma_b200\n"
            "from fma_b200 import vllm_loader\n"
            "cls = vllm_loader.register()\n"
            "from vllm.config.load import LoadConfig\n"
            "from vllm.model_executor.model_loader import get_model_loader\n"
            "from vllm.model_executor.model_loader.default_loader import DefaultModelLoader\n"
            "l = get_model_loader(LoadConfig(load_format='fma'))\n"
            "assert type(l) is cls and isinstance(l, DefaultModelLoader)\n"
            "print('ok')\n") % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ok" in r.stdout, (r.stdout + r.stderr)[-2000:]
