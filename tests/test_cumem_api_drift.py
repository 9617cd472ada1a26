"""B2 is kept exactly: the mirror class llm-d-fast-model-actuation_b200/cumem.py::CuMemAllocator has the same public methods,
parameter names and defaults as the installed vLLM's ``vllm/device_allocator/cumem.py::CuMemAllocator`` (the class
``Worker.sleep / wake_up / load_model`` call, vllm:v1/worker/gpu_worker.py:157-209).  Compared statically (ast): vLLM's module
cannot be imported on a box without libcuda."""
import ast
import importlib.util
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METHODS = ["get_instance", "sleep", "wake_up", "use_memory_pool", "get_current_usage"]


def _class_methods(path, cls):
    tree = ast.parse(open(path).read())
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == cls:
            out = {}
            for f in node.body:
                if isinstance(f, ast.FunctionDef):
                    a = f.args
                    names = [x.arg for x in a.posonlyargs + a.args]
                    defaults = [ast.unparse(d) for d in a.defaults]
                    out[f.name] = (names, defaults, [d.id if isinstance(d, ast.Name) else ast.unparse(d) for d in f.decorator_list])
            return out
    raise AssertionError(f"{cls} not found in {path}")


def test_mirror_has_the_reference_class_surface():
    spec = importlib.util.find_spec("vllm")
    if spec is None:
        pytest.skip("vllm not installed")
    ref_path = os.path.join(os.path.dirname(spec.origin), "device_allocator", "cumem.py")
    ref = _class_methods(ref_path, "CuMemAllocator")
    ours = _class_methods(os.path.join(ROOT, "llm-d-fast-model-actuation_b200", "cumem.py"), "CuMemAllocator")
    for m in METHODS:
        assert m in ref, f"vLLM's CuMemAllocator lost {m}: revisit the mirror"
        assert m in ours, f"mirror lacks {m}"
        rn, rd, rdec = ref[m]
        on, od, odec = ours[m]
        assert on == rn, (m, on, rn)                                   # same parameter names, same order
        assert od == rd, (m, od, rd)                                   # same defaults
        assert ("staticmethod" in rdec) == ("staticmethod" in odec) and ("contextmanager" in " ".join(rdec)) == ("contextmanager" in " ".join(odec)), m
