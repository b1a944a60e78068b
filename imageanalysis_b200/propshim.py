"""Minimal stand-in for the external `props` property-tree package
(aura-props "props-legacy"; reference environment.yml:44) that the reference
uses for all configuration (`from props import getNode`,
scripts/lib/matcher.py:15).  `imageanalysis_b200.matcher` prefers the real
package when it is installed and uses this only when it is absent, so that
the drop-in module keeps reading the same nodes
(`/config/detector/{detector,scale}`, `/config/matcher/{match_ratio,min_pairs,
min_dist,max_dist}`) the reference's `configure()` reads (matcher.py:49-80).

Only the API surface the matching path touches is provided.
"""
from __future__ import annotations

import json
from typing import Any, List


class PropertyNode:
    def __init__(self):
        pass

    # -- tree navigation ----------------------------------------------------
    def hasChild(self, name: str) -> bool:
        return name in self.__dict__

    def getChild(self, path: str, create: bool = False):
        node = self
        for tok in [t for t in path.split("/") if t]:
            nxt = node.__dict__.get(tok)
            if not isinstance(nxt, PropertyNode):
                if not create:
                    return None
                nxt = PropertyNode()
                node.__dict__[tok] = nxt
            node = nxt
        return node

    def getChildren(self, expand: bool = True) -> List[str]:
        return sorted(self.__dict__.keys())

    def isEnum(self, name: str) -> bool:
        return isinstance(self.__dict__.get(name), list)

    # -- typed getters (missing -> zero value, like props-legacy) -------------
    def getFloat(self, name: str) -> float:
        v = self.__dict__.get(name)
        try:
            return float(v) if v is not None and not isinstance(v, (PropertyNode, list)) else 0.0
        except (TypeError, ValueError):
            return 0.0

    def getInt(self, name: str) -> int:
        v = self.__dict__.get(name)
        try:
            return int(v) if v is not None and not isinstance(v, (PropertyNode, list)) else 0
        except (TypeError, ValueError):
            return 0

    def getBool(self, name: str) -> bool:
        v = self.__dict__.get(name)
        if isinstance(v, str):
            return v.lower() in ("true", "1", "yes")
        return bool(v) if not isinstance(v, (PropertyNode, list)) else False

    def getString(self, name: str) -> str:
        v = self.__dict__.get(name)
        return "" if v is None or isinstance(v, (PropertyNode, list)) else str(v)

    # -- setters ----------------------------------------------------------------
    def setFloat(self, name: str, val: float):
        self.__dict__[name] = float(val)

    def setInt(self, name: str, val: int):
        self.__dict__[name] = int(val)

    def setBool(self, name: str, val: bool):
        self.__dict__[name] = bool(val)

    def setString(self, name: str, val: str):
        self.__dict__[name] = str(val)

    # -- enumerated (list) values ---------------------------------------------
    def setLen(self, name: str, size: int, init_val: Any = None):
        cur = self.__dict__.get(name)
        if not isinstance(cur, list):
            cur = []
        while len(cur) < size:
            cur.append(init_val)
        self.__dict__[name] = cur[:size]

    def getLen(self, name: str) -> int:
        v = self.__dict__.get(name)
        return len(v) if isinstance(v, list) else 0

    def getFloatEnum(self, name: str, index: int) -> float:
        v = self.__dict__.get(name)
        if isinstance(v, list) and index < len(v) and v[index] is not None:
            return float(v[index])
        return 0.0

    def setFloatEnum(self, name: str, index: int, val: float):
        self.setLen(name, max(index + 1, self.getLen(name)), 0.0)
        self.__dict__[name][index] = float(val)

    def getStringEnum(self, name: str, index: int) -> str:
        v = self.__dict__.get(name)
        if isinstance(v, list) and index < len(v) and v[index] is not None:
            return str(v[index])
        return ""

    # -- (de)serialisation ---------------------------------------------------------
    def to_dict(self):
        out = {}
        for k, v in self.__dict__.items():
            out[k] = v.to_dict() if isinstance(v, PropertyNode) else v
        return out

    def from_dict(self, d: dict):
        for k, v in d.items():
            if isinstance(v, dict):
                self.getChild(k, True).from_dict(v)
            else:
                self.__dict__[k] = v

    def pretty_print(self, indent: str = ""):
        print(json.dumps(self.to_dict(), indent=2, default=str))


root = PropertyNode()


def getNode(path: str, create: bool = False):
    if path in ("", "/"):
        return root
    return root.getChild(path, create)


def load(filename: str, node: PropertyNode) -> bool:
    """props_json.load look-alike."""
    try:
        with open(filename) as f:
            node.from_dict(json.load(f))
        return True
    except (OSError, ValueError):
        return False


def save(filename: str, node: PropertyNode) -> bool:
    """props_json.save look-alike."""
    try:
        with open(filename, "w") as f:
            json.dump(node.to_dict(), f, indent=2, default=str)
        return True
    except OSError:
        return False
