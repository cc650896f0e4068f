"""TConfig: the universal typed key/value config object (reference: include/common/config.hpp:34-103,
src/common/config.cpp:89-223).  Serialises to type-tagged lossless JSON, e.g.
``{"name": "fc", "type": "dense", "parameters": {"input_features": {"type": "size_t", "value": 512}}}``.
"""
from __future__ import annotations

import json
from typing import Any


def _tag(v: Any):
    """Type-tagged JSON value with the reference's type names (src/common/config.cpp:89-126): nested configs and
    arbitrary structures are ``json`` (the reference stores ``LayerConfig::to_json()`` objects as nlohmann::json),
    homogeneous lists are ``vector<size_t|int|float|string>``."""
    if isinstance(v, bool):
        return {"type": "bool", "value": v}
    if isinstance(v, int):
        return {"type": "size_t" if v >= 0 else "int", "value": v}
    if isinstance(v, float):
        return {"type": "float", "value": v}
    if isinstance(v, str):
        return {"type": "string", "value": v}
    if isinstance(v, TConfig):
        return {"type": "json", "value": v.to_json()}
    if isinstance(v, (list, tuple)):
        if len(v) > 0 and all(isinstance(e, TConfig) for e in v):
            return {"type": "json", "value": [e.to_json() for e in v]}
        if len(v) > 0 and all(isinstance(e, bool) for e in v):
            return {"type": "json", "value": list(v)}
        if all(isinstance(e, int) and not isinstance(e, bool) for e in v):
            return {"type": "vector<size_t>" if all(e >= 0 for e in v) else "vector<int>", "value": list(v)}
        if all(isinstance(e, (int, float)) and not isinstance(e, bool) for e in v):
            return {"type": "vector<float>", "value": [float(e) for e in v]}
        if all(isinstance(e, str) for e in v):
            return {"type": "vector<string>", "value": list(v)}
        return {"type": "json", "value": [_plain(e) for e in v]}
    if isinstance(v, dict):
        return {"type": "json", "value": _plain(v)}
    if v is None:
        return {"type": "json", "value": None}
    raise TypeError(f"unsupported config value type {type(v)}")


def _plain(v: Any):
    """Plain JSON form of a value nested inside a ``json`` parameter (configs become their to_json())."""
    if isinstance(v, TConfig):
        return v.to_json()
    if isinstance(v, (list, tuple)):
        return [_plain(e) for e in v]
    if isinstance(v, dict):
        return {k: _plain(e) for k, e in v.items()}
    return v


def _is_config_json(d: Any) -> bool:
    return isinstance(d, dict) and "parameters" in d and "type" in d and "name" in d and isinstance(d["parameters"], dict)


def _revive(v: Any):
    """Inverse of _plain: dicts shaped like TConfig::to_json() come back as TConfig objects."""
    if _is_config_json(v):
        return TConfig.from_json(v)
    if isinstance(v, list):
        return [_revive(e) for e in v]
    return v


def _untag(d: Any):
    if not isinstance(d, dict) or "type" not in d:
        return d
    t, v = d["type"], d.get("value")
    if t == "config":                       # round-1 files
        return TConfig.from_json(v)
    if t == "config_array":
        return [TConfig.from_json(e) for e in v]
    if t in ("size_t", "int", "int64", "long", "long long", "unsigned int", "unsigned long", "unsigned long long"):
        return int(v)
    if t in ("float", "double"):
        return float(v)
    if t == "json":
        if isinstance(v, list) and v and all(isinstance(e, dict) and set(e) == {"type", "value"} for e in v):
            return [_untag(e) for e in v]   # round-1 heterogeneous lists
        return _revive(v)
    if t in ("vector<float>", "vector<double>", "float_array"):
        return [float(e) for e in v]
    return v


class TConfig:
    def __init__(self, name: str = "", type: str = "", **parameters):
        self.name = name
        self.type = type
        self.parameters: dict[str, Any] = dict(parameters)

    def set(self, key: str, value: Any) -> "TConfig":
        self.parameters[key] = value
        return self

    def get(self, key: str, default: Any = None) -> Any:
        return self.parameters.get(key, default)

    def has(self, key: str) -> bool:
        return key in self.parameters

    def __getitem__(self, key):
        return self.parameters[key]

    def __contains__(self, key):
        return key in self.parameters

    def to_json(self) -> dict:
        return {"name": self.name, "type": self.type,
                "parameters": {k: _tag(v) for k, v in self.parameters.items()}}

    def dump(self, indent: int | None = 4) -> str:
        return json.dumps(self.to_json(), indent=indent)

    @staticmethod
    def from_json(j) -> "TConfig":
        if isinstance(j, str):
            j = json.loads(j)
        c = TConfig(j.get("name", ""), j.get("type", ""))
        for k, v in j.get("parameters", {}).items():
            c.parameters[k] = _untag(v)
        return c

    def __eq__(self, other):
        return isinstance(other, TConfig) and self.to_json() == other.to_json()

    def __repr__(self):
        return f"TConfig(name={self.name!r}, type={self.type!r}, parameters={self.parameters!r})"
