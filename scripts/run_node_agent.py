"""Start the node agent (fma_b200.node_agent) — same CLI as the reference launcher (inference_server/launcher/launcher.py:857-899)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fma_b200  # noqa: F401,E402
from fma_b200 import node_agent  # noqa: E402

if __name__ == "__main__":
    node_agent.main()
