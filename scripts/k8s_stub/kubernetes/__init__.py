"""Minimal stand-in for the `kubernetes` package, which is not installed in this image.
The reference launcher only needs it to IMPORT gputranslator.py (inference_server/launcher/gputranslator.py:26);
its mock-GPU / ConfigMap path is never exercised by the engine tests."""
from . import client, config  # noqa: F401
