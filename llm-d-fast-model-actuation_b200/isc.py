"""InferenceServerConfig -> (VllmConfig, instance ID): the decision "same ID => wake the sleeper instead of creating"
(SURVEY.md §8f-2).

Restates ``controller.configInferenceServer`` (pkg/controller/dual-pods/inference-server.go:801-829):

    options      = spec.modelServerConfig.options + " --port " + str(port)
    VllmConfig   = {options, gpu_uuids, env_vars, annotations: {"isc-name": name, "inference-port": str(port)}}
                   (launcherclient.go:47-49,72-78)
    instance ID  = "I" + base64url_nopad(sha256(yaml(spec.modelServerConfig) + ";gpus=" + ",".join(gpu_uuids))) + "i"

The hashed bytes are ``sigs.k8s.io/yaml.Marshal`` output (go.mod:14, v1.6.0, on go.yaml.in/yaml/v2 v2.4.2, go.mod:38).
Neither module is under /root/reference (only go.sum lines) and there is no Go toolchain here, so this file restates
their PUBLISHED behaviour:

* sigs.k8s.io/yaml.Marshal = encoding/json.Marshal, then the JSON is read back as YAML into ``interface{}`` and
  marshalled by yaml.v2 — so struct field order is lost and **every map, the top level included, is emitted with
  sorted keys** (yaml.v2 ``keyList.Less``: a natural order that compares digit runs numerically and puts non-letters
  before letters);
* ``omitempty`` on options / env_vars / labels / annotations (api/fma/v1alpha1/inferenceserverconfig_types.go:35-62):
  empty strings and empty or nil maps do not appear; ``port`` always does;
* yaml.v2 ``encoder.stringv``: a string containing a newline is a literal block scalar; a string that would *resolve*
  to something other than a string when read back plain (``resolve()``: bools incl. y/n/yes/no/on/off, null, ints in
  Go base-0 syntax, floats, timestamps, base-60 floats) is double-quoted; everything else is handed to the emitter as
  plain, and the libyaml-derived emitter falls back to single quotes, then double quotes, when the text cannot be a
  plain scalar (``: ``, `` #``, leading indicators, leading/trailing blanks ...);
* the emitter folds plain and quoted scalars at blanks once a line passes 80 columns (v2.4 restored wrapping;
  sigs.k8s.io/yaml does not call ``FutureLineWrap``), indents by 2 and allows unicode.

PyYAML's emitter is a port of the same libyaml emitter yaml.v2's ``emitterc.go`` was ported from, so it is driven here
event by event (bypassing PyYAML's own resolver and representer, whose YAML-1.1 rules differ from yaml.v2's).

Pinning.  The MARSHALLER is pinned against the reference's own files: its three generated CRDs (config/crd/*.yaml, 8834
lines written by controller-gen v0.19.0 through the same sigs.k8s.io/yaml stack: sorted keys, 80-column folding, quoting,
literal blocks, indentless sequences) are parsed and re-emitted by ``go_yaml_marshal`` byte for byte
(tests/test_isc.py::test_marshal_reproduces_the_references_generated_crds, run where /root/reference exists).  The ID
composition around it (what is hashed, separator, base64 flavour, "I"/"i" wrapping) follows inference-server.go:801-829
line by line but stays **unpinned**: the reference has no test or document that fixes an expected ID for any input
(SURVEY.md §8f-2); only a live controller could confirm it.
"""
from __future__ import annotations

import base64
import functools
import hashlib
import io
import re
from typing import Dict, List, Mapping, Optional, Tuple

import yaml as _pyyaml
from yaml import events as _ev

ISC_NAME_ANNOTATION = "isc-name"          # launcherclient.go:48
ISC_PORT_ANNOTATION = "inference-port"    # launcherclient.go:49

# ------------------------------------------------------------------------------------------------------------------
# yaml.v2 resolve(): would this text, read back as a plain scalar, be something other than a string?
# ------------------------------------------------------------------------------------------------------------------
_RESOLVE_MAP = set(
    "y Y yes Yes YES on On ON true True TRUE n N no No NO off Off OFF false False FALSE "
    "~ null Null NULL .nan .NaN .NAN .inf .Inf .INF +.inf +.Inf +.INF -.inf -.Inf -.INF".split()) | {""}
_YAML_STYLE_FLOAT = re.compile(r"^[-+]?(\.[0-9]+|[0-9]+(\.[0-9]*)?)([eE][-+]?[0-9]+)?$")
_BASE60_FLOAT = re.compile(r"^[-+]?[0-9][0-9_]*(?::[0-5]?[0-9])+(?:\.[0-9_]*)?$")
_DOT_FLOAT = re.compile(r"^\.[0-9]+([eE][-+]?[0-9]+)?$")
_TS_DATE = r"(\d{4})-(\d{1,2})-(\d{1,2})"
_TS_TIME = r"(\d{1,2}):(\d{1,2}):(\d{1,2})(?:\.\d{1,9})?"
_TIMESTAMPS = [re.compile(rf"^{_TS_DATE}[Tt]{_TS_TIME}(?:Z|[+-]\d\d:\d\d)$"),
               re.compile(rf"^{_TS_DATE} {_TS_TIME}$"),
               re.compile(rf"^{_TS_DATE}$")]


def _go_parse_int(s: str) -> bool:
    """strconv.ParseInt / ParseUint(s, 0, 64) succeeds (sign, then 0x / 0o / 0b / leading-0 octal / decimal)."""
    neg = s.startswith("-")
    body = s[1:] if s[:1] in "+-" else s
    base = 10
    low = body.lower()
    if low.startswith("0x"):
        base, body = 16, body[2:]
    elif low.startswith("0o"):
        base, body = 8, body[2:]
    elif low.startswith("0b"):
        base, body = 2, body[2:]
    elif len(body) > 1 and body[0] == "0":
        base, body = 8, body[1:]
    if not body or any(c not in "0123456789abcdef"[:base] for c in body.lower()):
        return False
    v = int(body, base)
    return v <= (1 << 63) if neg else (v < (1 << 64) if s[:1] != "+" else v < (1 << 63))


def _is_timestamp(s: str) -> bool:
    for i, rx in enumerate(_TIMESTAMPS):
        m = rx.match(s)
        if not m:
            continue
        g = [int(x) for x in m.groups() if x is not None]
        if not (1 <= g[1] <= 12 and 1 <= g[2] <= 31):
            return False
        return len(g) == 3 or (g[3] < 24 and g[4] < 60 and g[5] < 60)
    return False


def resolves_to_non_string(s: str) -> bool:
    """yaml.v2 ``resolve("", s)`` returns a tag other than !!str, or ``isBase60Float(s)`` (so stringv must quote it)."""
    if s == "":
        return True
    c = s[0]
    if c in "yYnNtTfFoO~":
        hint = "M"
    elif c in "+-":
        hint = "S"
    elif c in "0123456789":
        hint = "D"
    elif c == ".":
        hint = "."
    else:
        return False
    if s in _RESOLVE_MAP:
        return True
    if hint == ".":
        return bool(_DOT_FLOAT.match(s))
    if hint in "DS":
        if _is_timestamp(s):
            return True
        plain = s.replace("_", "")
        if _go_parse_int(plain):
            return True
        if _YAML_STYLE_FLOAT.match(plain):
            return True
        for pre in ("0b", "-0b"):
            if plain.startswith(pre) and plain[len(pre):] and set(plain[len(pre):]) <= {"0", "1"}:
                return True
        if ":" in s and _BASE60_FLOAT.match(s):
            return True
    return False


# ------------------------------------------------------------------------------------------------------------------
# yaml.v2 map key order (sorter.go keyList.Less) for string keys
# ------------------------------------------------------------------------------------------------------------------
def _key_less(a: str, b: str) -> bool:
    n = min(len(a), len(b))
    for i in range(n):
        if a[i] == b[i]:
            continue
        al, bl = a[i].isalpha(), b[i].isalpha()
        if al and bl:
            return a[i] < b[i]
        if al or bl:
            return bl
        an = bn = 0
        if a[i] == "0" or b[i] == "0":
            j = i - 1
            while j >= 0 and a[j].isdigit():
                if a[j] != "0":
                    an = bn = 1
                    break
                j -= 1
        ai = i
        while ai < len(a) and a[ai].isdigit():
            an = an * 10 + ord(a[ai]) - 48
            ai += 1
        bi = i
        while bi < len(b) and b[bi].isdigit():
            bn = bn * 10 + ord(b[bi]) - 48
            bi += 1
        if an != bn:
            return an < bn
        if ai != bi:
            return ai < bi
        return a[i] < b[i]
    return len(a) < len(b)


def sorted_keys(keys) -> List[str]:
    return sorted(keys, key=functools.cmp_to_key(lambda a, b: -1 if _key_less(a, b) else (1 if _key_less(b, a) else 0)))


# ------------------------------------------------------------------------------------------------------------------
# emitter
# ------------------------------------------------------------------------------------------------------------------
def _scalar(text: str, style: Optional[str]) -> _ev.ScalarEvent:
    return _ev.ScalarEvent(anchor=None, tag=None, implicit=(True, True), value=text, style=style)


def _emit_value(out: list, v) -> None:
    if isinstance(v, (list, tuple)):
        out.append(_ev.SequenceStartEvent(anchor=None, tag=None, implicit=True, flow_style=not v))
        for x in v:
            _emit_value(out, x)
        out.append(_ev.SequenceEndEvent())
    elif isinstance(v, Mapping):
        out.append(_ev.MappingStartEvent(anchor=None, tag=None, implicit=True, flow_style=not v))
        for k in sorted_keys(v.keys()):
            _emit_value(out, k)
            _emit_value(out, v[k])
        out.append(_ev.MappingEndEvent())
    elif isinstance(v, bool):
        out.append(_scalar("true" if v else "false", None))
    elif isinstance(v, int):
        out.append(_scalar(str(v), None))
    elif isinstance(v, float):
        out.append(_scalar(repr(v), None))
    elif v is None:
        out.append(_scalar("null", None))
    elif isinstance(v, str):
        if "\n" in v:
            style = "|"
        elif resolves_to_non_string(v):
            style = '"'
        else:
            style = None
        out.append(_scalar(v, style))
    else:
        raise TypeError(f"unsupported value in a ModelServerConfig: {type(v).__name__}")


def go_yaml_marshal(obj: Mapping) -> bytes:
    """What ``sigs.k8s.io/yaml.Marshal`` writes for a JSON-shaped value (maps, lists, strings, ints, bools, null)."""
    evs: list = [_ev.StreamStartEvent(encoding=None), _ev.DocumentStartEvent(explicit=False)]
    _emit_value(evs, obj)
    evs += [_ev.DocumentEndEvent(explicit=False), _ev.StreamEndEvent()]
    buf = io.StringIO()
    em = _pyyaml.emitter.Emitter(buf, canonical=False, indent=2, width=80, allow_unicode=True, line_break="\n")
    for e in evs:
        em.emit(e)
    return buf.getvalue().encode("utf-8")


# ------------------------------------------------------------------------------------------------------------------
# configInferenceServer
# ------------------------------------------------------------------------------------------------------------------
def model_server_config_json(port: int, options: str = "", env_vars: Optional[Mapping[str, str]] = None,
                             labels: Optional[Mapping[str, str]] = None,
                             annotations: Optional[Mapping[str, str]] = None) -> Dict:
    """The JSON object encoding/json produces for a ModelServerConfig (omitempty applied)."""
    d: Dict = {"port": int(port)}
    if options:
        d["options"] = options
    for key, m in (("env_vars", env_vars), ("labels", labels), ("annotations", annotations)):
        if m:
            d[key] = dict(m)
    return d


def config_inference_server(isc_name: str, port: int, options: str = "", env_vars: Optional[Mapping[str, str]] = None,
                            labels: Optional[Mapping[str, str]] = None, annotations: Optional[Mapping[str, str]] = None,
                            gpu_uuids: Optional[List[str]] = None) -> Tuple[Dict, str]:
    """-> (VllmConfig as the launcher's JSON body, instance ID).  inference-server.go:801-829."""
    if not (1 <= int(port) <= 65535):                 # +kubebuilder:validation:Minimum=1/Maximum=65535
        raise ValueError(f"port {port} outside 1..65535")
    gpu_uuids = list(gpu_uuids or [])
    port_s = str(int(port))
    cfg: Dict = {"options": options + " --port " + port_s}
    if gpu_uuids:                                     # json:"gpu_uuids,omitempty"
        cfg["gpu_uuids"] = gpu_uuids
    if env_vars:
        cfg["env_vars"] = dict(env_vars)
    cfg["annotations"] = {ISC_NAME_ANNOTATION: isc_name, ISC_PORT_ANNOTATION: port_s}
    h = hashlib.sha256()
    h.update(go_yaml_marshal(model_server_config_json(port, options, env_vars, labels, annotations)))
    h.update(b";gpus=")
    h.update(",".join(gpu_uuids).encode())
    return cfg, "I" + base64.urlsafe_b64encode(h.digest()).rstrip(b"=").decode() + "i"


def instance_id(isc_spec: Mapping, gpu_uuids: Optional[List[str]] = None) -> str:
    """Instance ID from an InferenceServerConfig manifest's ``spec`` (a dict as read from YAML/JSON)."""
    msc = isc_spec["modelServerConfig"]
    return config_inference_server("", msc["port"], msc.get("options", ""), msc.get("env_vars"), msc.get("labels"),
                                   msc.get("annotations"), gpu_uuids)[1]


# ------------------------------------------------------------------------------------------------------------------
# "same ID => wake": what the controller decides from a launcher's instance list
# ------------------------------------------------------------------------------------------------------------------
INFERENCE_PORT_ANNOTATION = "inference-port"      # launcherclient.go:49


def instance_port(inst: Mapping) -> int:
    """``getVLLMInstancePort`` (inference-server.go:968-977): the port comes from the instance's annotation, nowhere else."""
    ann = inst.get("annotations") or {}
    if INFERENCE_PORT_ANNOTATION not in ann:
        raise ValueError(f"missing annotations[{INFERENCE_PORT_ANNOTATION}]")
    v = str(ann[INFERENCE_PORT_ANNOTATION])
    if not re.fullmatch(r"[+-]?[0-9]+", v) or not -(1 << 31) <= int(v) < (1 << 31):     # strconv.ParseInt(value, 10, 32)
        raise ValueError(f"parse annotations[{INFERENCE_PORT_ANNOTATION}] value {v!r}")
    return int(v)


def select_launcher(launchers, isc_hash: str, desired_port: int, max_others: int):
    """Restatement of ``selectBestLauncherPod`` (inference-server.go:680-785) over launcher states as the launcher REST returns
    them (``GET /v2/vllm/instances``: ``{total_instances, running_instances, instances: [{instance_id, status, annotations}]}``).

    ``launchers``: iterable of ``(name, state_or_None)`` — ``None`` = the launcher is not ready / could not be synced.
    Returns ``(name | None, has_sleeping_instance, some_not_ready)``:
      priority 1  a launcher that already holds an instance with THIS id (status != "stopped") -> wake it (the fast path);
      priority 2  the first launcher with room for one more instance (total_instances <= max_others) -> create there;
      a launcher where another instance already uses the desired port, or an instance has no usable port, is skipped."""
    candidate = None
    some_not_ready = False
    for name, state in launchers:
        if state is None:
            some_not_ready = True
            continue
        has_sleeping = False
        port_conflict = False
        for inst in state.get("instances", []):
            try:
                port = instance_port(inst)
            except ValueError:
                port_conflict = True
                break
            if port == desired_port and inst.get("instance_id") != isc_hash:
                port_conflict = True
                break
            if inst.get("instance_id") == isc_hash and inst.get("status") != "stopped":
                has_sleeping = True
        if port_conflict:
            continue
        if has_sleeping:
            return name, True, False
        if state.get("total_instances", 0) <= max_others and candidate is None:
            candidate = name
    if candidate is not None:
        return candidate, False, False
    if some_not_ready:
        return None, False, True
    return None, False, False


def plan_actuation(launchers, isc_spec: Mapping, gpu_uuids=None, max_others: int = 1) -> dict:
    """ISC + GPUs -> instance ID (``instance_id``) -> what to do on which launcher: ``wake`` the instance with that ID where it
    sleeps, ``create`` it on a launcher with room, ``retry`` while launchers are not ready, or ``new_launcher``."""
    iid = instance_id(isc_spec, gpu_uuids)
    port = int(isc_spec["modelServerConfig"]["port"])
    name, sleeping, not_ready = select_launcher(launchers, iid, port, max_others)
    action = "wake" if sleeping else "create" if name is not None else "retry" if not_ready else "new_launcher"
    return {"action": action, "launcher": name, "instance_id": iid, "port": port}
