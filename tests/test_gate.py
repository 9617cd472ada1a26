"""The cross-process VMM gate (csrc/fma_gate.cu) with real processes: mutual exclusion, priority order, bounded waits, recovery from a
holder that was SIGKILLed.  The gate is on by default and sits on the wake's critical path at N > 1; plain and under
ThreadSanitizer."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "llm-d-fast-model-actuation_b200", "csrc")


@pytest.mark.parametrize("san", ["plain", "tsan"])
def test_gate_across_processes(tmp_path, san):
    exe = str(tmp_path / f"gate_test_{san}")
    flags = ["-O1", "-fsanitize=thread"] if san == "tsan" else ["-O2"]
    subprocess.check_call(["g++", "-std=c++17", "-g", *flags, "-I" + CSRC, os.path.join(ROOT, "tests", "cpp", "gate", "gate_test.cpp"),
                           "-x", "c++", os.path.join(CSRC, "fma_gate.cu"), "-o", exe, "-lpthread", "-lrt"])
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=1")
    if san == "tsan":
        env["GATE_TEST_SKIP_ROBUST"] = "1"      # TSan's mutex model has no EOWNERDEAD recovery (false "unlock of an unlocked mutex")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0 and "gate test ok" in r.stdout, (r.stdout + r.stderr)[-3000:]
    assert "WARNING: ThreadSanitizer" not in r.stderr, r.stderr[-3000:]
