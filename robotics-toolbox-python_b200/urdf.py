"""URDF / xacro ingestion: a robot description file -> the link tree ``Robot`` is built from.

Mirrors what the reference's loader produces (reference tools/urdf/urdf.py:1694-1780, Robot.URDF_read): one
``Link`` per URDF link, in file order; a child link's ETS is the joint's constant origin -- ``ET.SE3(transl(xyz) *
rpy2tr(rpy))`` with URDF's fixed-axis roll-pitch-yaw, R = Rz(yaw) Ry(pitch) Rx(roll) -- followed by ONE variable
transform for a revolute / continuous / prismatic joint about +-x, +-y or +-z (a negative axis becomes ``flip``); a
joint axis that is not a coordinate axis is rotated onto z by a constant folded into the origin (urdf.py:1711-1724).
Mass and centre of mass come from ``<inertial>``; limits from ``<limit>``.

xacro files are expanded first by the small macro processor below (properties, ``${}`` expressions, macros with
value / default / block parameters, include with ``$(find pkg)``, arg, if / unless, insert_block) -- enough for the
descriptions the reference ships (ur_description, franka_description, ...).  Geometry / collision / transmission
elements are parsed past: they do not touch the kinematics or dynamics path.
"""
from __future__ import annotations

import copy
import math
import os
import re
import xml.etree.ElementTree as ET_
from typing import Dict, List, Optional, Tuple

import numpy as np


# ------------------------------------------------------------------ a small xacro processor
def _is_xacro(tag) -> bool:
    return isinstance(tag, str) and tag.startswith("{") and "xacro" in tag.split("}")[0]


def _local(tag: str) -> str:
    return tag.split("}")[-1]


_MATH = {k: getattr(math, k) for k in ("pi", "sin", "cos", "tan", "asin", "acos", "atan", "atan2", "sqrt", "radians", "degrees",
                                       "fabs", "floor", "ceil", "pow", "log", "exp")}
_MATH.update({"abs": abs, "min": min, "max": max, "round": round, "float": float, "int": int, "str": str, "True": True, "False": False,
              "true": True, "false": False})


class _Scope:
    def __init__(self, parent: Optional["_Scope"] = None):
        self.vars: Dict[str, object] = {}
        self.blocks: Dict[str, List[ET_.Element]] = {}
        self.parent = parent

    def lookup(self, name):
        s = self
        while s is not None:
            if name in s.vars:
                return s, s.vars[name]
            s = s.parent
        raise KeyError(name)

    def block(self, name):
        s = self
        while s is not None:
            if name in s.blocks:
                return s.blocks[name]
            s = s.parent
        raise KeyError(name)


class _Names(dict):
    """eval() namespace: properties are evaluated on first use (a property may be defined through others)."""

    def __init__(self, xp, scope):
        super().__init__(_MATH)
        self.xp, self.scope = xp, scope

    def __missing__(self, key):
        try:
            s, v = self.scope.lookup(key)
        except KeyError:
            raise NameError(f"xacro: undefined property {key!r}")
        if isinstance(v, str):
            v = self.xp.eval_text(v, s, raw=True)
            s.vars[key] = v
        return v


class Xacro:
    def __init__(self, args: Optional[Dict[str, str]] = None):
        self.macros: Dict[str, Tuple[List[Tuple[str, Optional[str], int]], ET_.Element]] = {}
        self.args: Dict[str, str] = dict(args or {})
        self.root_dir = None

    # ---- text substitution
    def _find(self, pkg: str, here: str) -> str:
        d = os.path.abspath(here)
        while True:
            cand = os.path.join(d, pkg)
            if os.path.isdir(cand):
                return cand
            if os.path.basename(d) == pkg:
                return d
            nd = os.path.dirname(d)
            if nd == d:
                raise FileNotFoundError(f"xacro: cannot locate package {pkg!r} above {here}")
            d = nd

    def eval_text(self, text: str, scope: _Scope, raw: bool = False, here: str = "."):
        if text is None:
            return text

        def dollar_paren(m):
            words = m.group(1).split()
            if words[0] == "find":
                return self._find(words[1], here)
            if words[0] == "arg":
                if words[1] not in self.args:
                    raise KeyError(f"xacro: undefined arg {words[1]!r}")
                return str(self.args[words[1]])
            raise ValueError(f"xacro: unsupported substitution $({m.group(1)})")

        text = re.sub(r"\$\(([^)]*)\)", dollar_paren, text)
        parts = re.split(r"(\$\{[^}]*\})", text)
        if raw and len(parts) == 3 and parts[0] == "" and parts[2] == "":
            return eval(parts[1][2:-1], {"__builtins__": {}}, _Names(self, scope))  # the value itself (number, bool, ...)
        out = []
        for p in parts:
            if p.startswith("${") and p.endswith("}"):
                v = eval(p[2:-1], {"__builtins__": {}}, _Names(self, scope))
                out.append(repr(v) if isinstance(v, float) else str(v))
            else:
                out.append(p)
        s = "".join(out)
        if raw:
            try:
                return float(s) if re.fullmatch(r"\s*[-+]?(\d+\.?\d*|\.\d+)([eE][-+]?\d+)?\s*", s) else s
            except ValueError:
                return s
        return s

    @staticmethod
    def _truth(v) -> bool:
        if isinstance(v, str):
            t = v.strip().lower()
            if t in ("true", "1", "1.0"):
                return True
            if t in ("false", "0", "0.0", ""):
                return False
            raise ValueError(f"xacro: cannot interpret {v!r} as a boolean")
        return bool(v)

    # ---- tree expansion
    def expand_children(self, elem: ET_.Element, scope: _Scope, here: str) -> List[ET_.Element]:
        out: List[ET_.Element] = []
        for child in list(elem):
            out.extend(self.expand(child, scope, here))
        return out

    def expand(self, e: ET_.Element, scope: _Scope, here: str) -> List[ET_.Element]:
        if not isinstance(e.tag, str):  # comments / processing instructions
            return []
        if _is_xacro(e.tag):
            name = _local(e.tag)
            if name == "include":
                fn = self.eval_text(e.get("filename"), scope, here=here)
                if not os.path.isabs(fn):
                    fn = os.path.join(here, fn)
                root = ET_.parse(fn).getroot()
                return self.expand_children(root, scope, os.path.dirname(fn))
            if name == "property":
                pname = e.get("name")
                if e.get("value") is not None:
                    scope.vars[pname] = e.get("value")  # evaluated on first use
                elif e.get("default") is not None:
                    if pname not in scope.vars:
                        scope.vars[pname] = e.get("default")
                else:
                    scope.blocks[pname] = [copy.deepcopy(c) for c in e]
                return []
            if name == "arg":
                self.args.setdefault(e.get("name"), self.eval_text(e.get("default", ""), scope, here=here))
                return []
            if name == "macro":
                params = []
                for tok in re.findall(r"""\*{0,2}[\w.]+(?::=(?:'[^']*'|"[^"]*"|\S+))?""", e.get("params") or ""):
                    stars = len(tok) - len(tok.lstrip("*"))
                    tok = tok.lstrip("*")
                    default = None
                    if ":=" in tok:
                        tok, default = tok.split(":=", 1)
                        if default.startswith("^"):
                            default = default.lstrip("^|") or None
                        if default is not None and len(default) >= 2 and default[0] == default[-1] and default[0] in "'\"":
                            default = default[1:-1]
                    params.append((tok, default, stars))
                self.macros[e.get("name").replace("xacro:", "")] = (params, e)
                return []
            if name in ("if", "unless"):
                v = self._truth(self.eval_text(e.get("value"), scope, raw=True, here=here))
                return self.expand_children(e, scope, here) if v == (name == "if") else []
            if name == "insert_block":
                res = []
                for b in scope.block(e.get("name")):
                    res.extend(self.expand(copy.deepcopy(b), scope, here))
                return res
            if name in self.macros:
                params, body = self.macros[name]
                inner = _Scope(scope)
                kids = [c for c in e if isinstance(c.tag, str)]
                for pname, default, stars in params:
                    if stars == 0:
                        if e.get(pname) is not None:
                            inner.vars[pname] = self.eval_text(e.get(pname), scope, raw=True, here=here)
                        elif default is not None:
                            inner.vars[pname] = self.eval_text(default, scope, raw=True, here=here)
                        else:
                            raise ValueError(f"xacro: macro {name} is missing parameter {pname}")
                    else:
                        if not kids:
                            raise ValueError(f"xacro: macro {name} is missing block parameter {pname}")
                        k = kids.pop(0)
                        blk = self.expand(k, scope, here) if stars == 1 else self.expand_children(k, scope, here)
                        inner.blocks[pname] = blk
                return self.expand_children(body, inner, here)
            raise ValueError(f"xacro: unknown element or macro <xacro:{name}>")
        new = ET_.Element(e.tag, {k: self.eval_text(v, scope, here=here) for k, v in e.attrib.items() if not k.startswith("{")})
        new.text = self.eval_text(e.text, scope, here=here) if e.text and e.text.strip() else e.text
        for c in self.expand_children(e, scope, here):
            new.append(c)
        return [new]

    def process_file(self, path: str) -> ET_.Element:
        root = ET_.parse(path).getroot()
        here = os.path.dirname(os.path.abspath(path))
        scope = _Scope()
        out = ET_.Element(_local(root.tag) if _is_xacro(root.tag) else root.tag, {k: v for k, v in root.attrib.items() if not k.startswith("{")})
        for c in self.expand_children(root, scope, here):
            out.append(c)
        return out


# ------------------------------------------------------------------ URDF -> links
def _floats(text, n, default):
    if text is None:
        return np.array(default, dtype=np.float64)
    v = np.array([float(x) for x in text.split()], dtype=np.float64)
    if v.size != n:
        raise ValueError(f"URDF: expected {n} numbers, got {text!r}")
    return v


def _rpy2r(rpy):
    """URDF fixed-axis roll-pitch-yaw = spatialmath SE3.RPY(order='zyx'): R = Rz(yaw) Ry(pitch) Rx(roll)"""
    r, p, y = rpy
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def _origin(elem) -> np.ndarray:
    T = np.eye(4)
    if elem is not None:
        T[:3, 3] = _floats(elem.get("xyz"), 3, [0, 0, 0])
        T[:3, :3] = _rpy2r(_floats(elem.get("rpy"), 3, [0, 0, 0]))
    return T


def _angvec2r(theta, v):
    """Rodrigues (spatialmath angvec2r)"""
    v = np.asarray(v, dtype=np.float64)
    sk = np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])
    return np.eye(3) + math.sin(theta) * sk + (1 - math.cos(theta)) * sk @ sk


def urdf_to_links(file_path: str, args: Optional[Dict[str, str]] = None):
    """-> (links, robot name).  ``file_path`` is a .urdf or a .xacro / .urdf.xacro file."""
    from .ET import ET
    from .ETS import ETS
    from .Robot import Link

    if not os.path.exists(file_path):
        raise FileNotFoundError(file_path)
    if file_path.endswith(".xacro"):
        root = Xacro(args).process_file(file_path)
    else:
        root = ET_.parse(file_path).getroot()
    if _local(root.tag) != "robot":
        raise ValueError("URDF: the root element must be <robot>")
    name = root.get("name", "")
    links, by_name = [], {}
    for le in root.findall("link"):
        m, r, I = 0.0, None, None
        ine = le.find("inertial")
        if ine is not None:
            if ine.find("mass") is not None:
                m = float(ine.find("mass").get("value"))
            r = _origin(ine.find("origin"))[:3, 3]
            it = ine.find("inertia")
            if it is not None:
                g = lambda k: float(it.get(k, 0.0))  # noqa: E731
                I = np.array([[g("ixx"), g("ixy"), g("ixz")], [g("ixy"), g("iyy"), g("iyz")], [g("ixz"), g("iyz"), g("izz")]])  # noqa: E741
        link = Link(ETS(), name=le.get("name"), m=m, r=r, I=I)
        links.append(link)
        if link.name in by_name:
            raise ValueError(f"URDF: duplicate link name {link.name}")
        by_name[link.name] = link
    seen_children = set()
    for je in root.findall("joint"):
        parent, child = je.find("parent").get("link"), je.find("child").get("link")
        if parent not in by_name or child not in by_name:
            raise ValueError(f"URDF: joint {je.get('name')} connects unknown links {parent!r} -> {child!r}")
        if child in seen_children:
            raise ValueError(f"URDF: link {child} has two parent joints")
        seen_children.add(child)
        T = _origin(je.find("origin"))
        jtype = je.get("type")
        axis = _floats(je.find("axis").get("xyz") if je.find("axis") is not None else None, 3, [1, 0, 0])
        if jtype in ("revolute", "continuous", "prismatic") and np.count_nonzero(axis) >= 2:
            # not a coordinate axis: rotate it onto z with a constant (urdf.py:1711-1724)
            u = axis / np.linalg.norm(axis)
            z = np.array([0.0, 0.0, 1.0])
            c = float(np.dot(z, u))
            k = np.cross(z, u)
            nk = np.linalg.norm(k)
            R = np.eye(3) if nk < 1e-12 else _angvec2r(math.atan2(nk, c), k / nk)
            T[:3, :3] = T[:3, :3] @ R
            axis = z
        ets = [ET.SE3(T)]
        if jtype in ("revolute", "continuous", "prismatic"):
            k = int(np.argmax(np.abs(axis)))
            kind = ("Rx", "Ry", "Rz")[k] if jtype != "prismatic" else ("tx", "ty", "tz")[k]
            qlim = None
            lim = je.find("limit")
            if lim is not None and jtype != "continuous" and lim.get("lower") is not None and lim.get("upper") is not None:
                qlim = [float(lim.get("lower")), float(lim.get("upper"))]
            ets.append(ET(kind, flip=bool(axis[k] < 0), qlim=qlim))
        elif jtype not in ("fixed",):
            raise ValueError(f"URDF: joint type {jtype!r} is not supported (revolute, continuous, prismatic, fixed)")
        cl = by_name[child]
        cl.ets = ETS(ets)
        cl.parent = by_name[parent]
        cl.joint_name = je.get("name")
        dyn = je.find("dynamics")
        if dyn is not None and dyn.get("friction") is not None:
            cl.B = float(dyn.get("friction"))
    return links, name
