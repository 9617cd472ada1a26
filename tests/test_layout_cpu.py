"""The engine's address-space layout logic (VA arenas with first-fit reuse, run planning) is plain C++ in
csrc/fma_layout.h and is what the engine (csrc/fma_engine.cu, fma_wake.cu) compiles in; exercised here on CPU with g++ (incl. a 20k-step randomised
alloc/free invariant check)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_layout_logic(tmp_path):
    exe = str(tmp_path / "layout_test")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I",
                           os.path.join(ROOT, "llm-d-fast-model-actuation_b200", "csrc"),
                           os.path.join(ROOT, "tests", "cpp", "layout_test.cpp"), "-o", exe])
    out = subprocess.check_output([exe], text=True)
    assert "layout ok" in out


def test_engine_uses_the_tested_header():
    csrc = os.path.join(ROOT, "llm-d-fast-model-actuation_b200", "csrc")
    src = "".join(open(os.path.join(csrc, f)).read() for f in ("fma_internal.h", "fma_engine.cu", "fma_wake.cu"))
    assert '#include "fma_layout.h"' in src and "fma_layout::plan_runs(" in src and "fma_layout::arena_take(" in src
