#!/bin/bash
# Round 2, GPU call 11 (2 GPUs): MULTI-PATH wake across processes — parity test with a restricted instance, then the 8B table.
set -u
out=gpurun_out/r2c11
mkdir -p "$out"
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -rfEs --timeout 200 -k "across_processes or multipath_wake_matches or parking_buffer" > "$out/pytest.log" 2>&1; echo "pytest rc=$?" | tee "$out/status.txt"; tail -6 "$out/pytest.log"
timeout 400 python scripts/remote_multipath_bench.py 1 > "$out/remote_bench.log" 2>&1; echo "remote bench rc=$?" | tee -a "$out/status.txt"; tail -6 "$out/remote_bench.log" | cut -c1-600
cp gpurun_out/remote_mp/remote_multipath.json "$out/" 2>/dev/null
