"""TConfig: the universal typed key/value config object (reference: include/common/config.hpp:34-103,
src/common/config.cpp:89-223).  Serialises to type-tagged lossless JSON, e.g.
``{"name": "fc", "type": "dense", "parameters": {"input_features": {"type": "size_t", "value": 512}}}``.
"""
from __future__ import annotations

import json
from typing import Any


def _tag(v: Any):
    if isinstance(v, bool):
        return {"type": "bool", "value": v}
    if isinstance(v, int):
        return {"type": "size_t" if v >= 0 else "int", "value": v}
    if isinstance(v, float):
        return {"type": "float", "value": v}
    if isinstance(v, str):
        return {"type": "string", "value": v}
    if isinstance(v, TConfig):
        return {"type": "config", "value": v.to_json()}
    if isinstance(v, (list, tuple)):
        if all(isinstance(e, TConfig) for e in v) and len(v) > 0:
            return {"type": "config_array", "value": [e.to_json() for e in v]}
        if all(isinstance(e, bool) for e in v) and len(v) > 0:
            return {"type": "bool_array", "value": list(v)}
        if all(isinstance(e, int) for e in v):
            return {"type": "size_t_array", "value": list(v)}
        if all(isinstance(e, (int, float)) for e in v):
            return {"type": "float_array", "value": [float(e) for e in v]}
        if all(isinstance(e, str) for e in v):
            return {"type": "string_array", "value": list(v)}
        return {"type": "json", "value": [_tag(e) for e in v]}
    if isinstance(v, dict):
        return {"type": "json", "value": v}
    if v is None:
        return {"type": "null", "value": None}
    raise TypeError(f"unsupported config value type {type(v)}")


def _untag(d: Any):
    if not isinstance(d, dict) or "type" not in d:
        return d
    t, v = d["type"], d.get("value")
    if t == "config":
        return TConfig.from_json(v)
    if t == "config_array":
        return [TConfig.from_json(e) for e in v]
    if t in ("size_t", "int", "int64"):
        return int(v)
    if t in ("float", "double"):
        return float(v)
    if t == "json" and isinstance(v, list):
        return [_untag(e) for e in v]
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
