"""Stand-in for the `vllm` package used ONLY to stage BASELINE config 0 on a machine without a GPU.

The installed vLLM is the CUDA wheel; without a GPU it resolves to UnspecifiedPlatform and cannot serve, and the CPU
wheel of dockerfiles/Dockerfile.launcher.cpu is not reachable.  vLLM's CPU worker makes sleep/wake_up no-ops
(vllm:v1/worker/cpu_worker.py:148-154), so the plumbing under test is: unmodified reference launcher -> fork ->
`run_server(args)` -> the three dev-mode routes.  This package provides exactly the four names launcher.py imports
(inference_server/launcher/launcher.py:38-41) and forwards run_server to fma_b200.server.run_server."""
__version__ = "0.0-cpu-standin"
