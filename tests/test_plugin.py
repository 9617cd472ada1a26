"""The vLLM general plugin that routes sleep/wake to the engine is discoverable the way vLLM discovers plugins
(importlib.metadata entry points, group vllm.general_plugins) and is inert unless FMA_B200=1."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(code, env_extra=None):
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "plugin"), ROOT, env.get("PYTHONPATH", "")])
    env.update(env_extra or {})
    return subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)


def test_entry_point_is_discoverable_and_inert_by_default():
    r = _run("from importlib.metadata import entry_points\n"
             "eps=[e for e in entry_points(group='vllm.general_plugins') if e.name=='fma_b200']\n"
             "assert len(eps)==1 and eps[0].value=='fma_b200_vllm_plugin:register'\n"
             "import sys; eps[0].load()()\n"
             "assert 'fma_b200.cumem' not in sys.modules   # FMA_B200 unset: nothing is touched\n"
             "print('ok')")
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-800:]


def test_kubernetes_stub_lets_the_reference_translator_import():
    """inference_server/launcher/gputranslator.py:26 only needs `from kubernetes import client, config` to resolve."""
    env = {"PYTHONPATH": os.pathsep.join([os.path.join(ROOT, "scripts", "k8s_stub"), os.environ.get("PYTHONPATH", "")])}
    r = subprocess.run([sys.executable, "-c", "from kubernetes import client, config\n"
                        "try:\n    config.load_incluster_config()\nexcept config.ConfigException:\n    print('ok')"],
                       env={**os.environ, **env}, capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-500:]
