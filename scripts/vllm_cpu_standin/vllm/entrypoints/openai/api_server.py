from fma_b200.server import run_server  # noqa: F401  (CPU-worker semantics: state flips, nothing moves)
