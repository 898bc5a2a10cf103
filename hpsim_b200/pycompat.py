"""Runs the reference's own Python-2 layer -- `pylib/hpsim.py`, `lcsutil.py`, `nputil.py`, `sqldb.py` and the
user scripts under `pytest/` -- UNCHANGED on Python 3, on top of this package's `HPSim` module.

The reference ships a CPython-2.7 extension `HPSim.so` and a Python-2 convenience layer `pylib/hpsim.py`
(SURVEY.md section 1).  The files are neither copied nor rewritten here: they are read from a reference
checkout at import time, translated in memory and executed with `HPSim` bound to `hpsim_b200.HPSim`:

  * text level: `print` statements -> `print(...)` calls, `<>` -> `!=`;
  * AST level: every `/` becomes a call that keeps Python-2 semantics (int / int floors), so expressions
    such as `num=1024*256/4` (pytest/sim-lbeg.py:58) stay integers;
  * namespace level: `map`, `filter`, `zip`, `range` return lists as in Python 2;
  * matplotlib, which the layer imports for its plots, is replaced by inert stand-ins when it is not
    installed (this image has none); everything numerical runs as written.

    from hpsim_b200 import pycompat
    hps = pycompat.load_pylib("/path/to/hpsim/pylib")        # the reference's hpsim.py as a module
    pycompat.run_script("/path/to/hpsim/pytest/sim-lbeg.py")   # cwd = the script's directory, like `python sim-lbeg.py`
"""
from __future__ import annotations

import ast
import builtins
import importlib.abc
import importlib.util
import os
import re
import sys
import types

_PRINT = re.compile(r"^(\s*)print(\s+.*|\s*)$")


def _split_top_level(s: str):
    """Splits a print statement's argument list at top-level commas (outside brackets and strings)."""
    parts, depth, cur, quote, i = [], 0, [], None, 0
    while i < len(s):
        ch = s[i]
        if quote:
            cur.append(ch)
            if ch == "\\":
                i += 1
                if i < len(s):
                    cur.append(s[i])
            elif s.startswith(quote, i):
                cur.extend(quote[1:])
                i += len(quote) - 1
                quote = None
        elif ch in "\"'":
            quote = s[i:i + 3] if s[i:i + 3] in ('"""', "'''") else ch
            cur.extend(quote)
            i += len(quote) - 1
        elif ch in "([{":
            depth += 1
            cur.append(ch)
        elif ch in ")]}":
            depth -= 1
            cur.append(ch)
        elif ch == "," and depth == 0:
            parts.append("".join(cur))
            cur = []
        elif ch == "#" and depth == 0:
            break
        else:
            cur.append(ch)
        i += 1
    parts.append("".join(cur))
    return parts


def _fix_print(line: str) -> str:
    m = _PRINT.match(line)
    if not m:
        return line
    indent, rest = m.group(1), m.group(2).strip()
    if rest.startswith("(") and _balanced_call(rest):
        return line  # already a call with a single parenthesised argument list
    if not rest:
        return f"{indent}print()"
    target = ""
    if rest.startswith(">>"):
        parts = _split_top_level(rest[2:])
        target = f", file={parts[0].strip()}"
        parts = parts[1:]
    else:
        parts = _split_top_level(rest)
    end = ""
    if parts and parts[-1].strip() == "":
        parts = parts[:-1]
        end = ", end=' '"
    return f"{indent}print({', '.join(p.strip() for p in parts)}{target}{end})"


def _balanced_call(rest: str) -> bool:
    """True when `rest` is exactly one parenthesised group (print("a", b)), not `("a") % x` etc."""
    depth = 0
    for i, ch in enumerate(rest):
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
            if depth == 0:
                return rest[i + 1:].split("#")[0].strip() == ""
    return False


def translate(source: str) -> str:
    """Python-2 source of the reference's pylib / scripts -> source that parses on Python 3."""
    out, pending = [], None
    for line in source.expandtabs(8).splitlines():
        if pending is not None:  # a print statement continued over several lines
            pending += " " + line.strip()
            if _open_brackets(pending) > 0 or pending.rstrip().endswith("\\"):
                continue
            line, pending = pending.replace("\\ ", " "), None
        elif _PRINT.match(line) and (_open_brackets(line) > 0 or line.rstrip().endswith("\\")):
            pending = line.rstrip().rstrip("\\")
            continue
        line = _fix_print(line)
        if "<>" in line:
            line = line.replace("<>", "!=")
        out.append(line)
    return "\n".join(out) + "\n"


def _open_brackets(s: str) -> int:
    depth, quote = 0, None
    for ch in s.split("#")[0] if '"' not in s and "'" not in s else s:
        if quote:
            if ch == quote:
                quote = None
        elif ch in "\"'":
            quote = ch
        elif ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
    return depth


def py2div(a, b):
    """Python-2 `/`: floor division for two integers, true division otherwise."""
    if isinstance(a, (int,)) and isinstance(b, (int,)) and not isinstance(a, bool) and not isinstance(b, bool):
        return a // b
    return a / b


class _Div(ast.NodeTransformer):
    def visit_BinOp(self, node):
        self.generic_visit(node)
        if isinstance(node.op, ast.Div):
            return ast.copy_location(ast.Call(func=ast.Name(id="__py2div__", ctx=ast.Load()), args=[node.left, node.right], keywords=[]), node)
        return node

    def visit_AugAssign(self, node):
        self.generic_visit(node)
        if isinstance(node.op, ast.Div):
            load = ast.parse(ast.unparse(node.target), mode="eval").body
            call = ast.Call(func=ast.Name(id="__py2div__", ctx=ast.Load()), args=[load, node.value], keywords=[])
            return ast.copy_location(ast.Assign(targets=[node.target], value=call), node)
        return node


def compile_py2(source: str, filename: str):
    tree = ast.parse(translate(source), filename=filename)
    tree = ast.fix_missing_locations(_Div().visit(tree))
    return compile(tree, filename, "exec")


def _py2_namespace(ns: dict) -> dict:
    ns["__py2div__"] = py2div
    ns["map"] = lambda *a: list(builtins.map(*a))
    ns["filter"] = lambda *a: list(builtins.filter(*a))
    ns["zip"] = lambda *a: list(builtins.zip(*a))
    ns["range"] = lambda *a: list(builtins.range(*a))
    ns["xrange"] = builtins.range
    ns["raw_input"] = builtins.input
    return ns


# ------------------------------------------------------------------ inert matplotlib when it is not installed
class _Inert:
    """Accepts any attribute access, call, indexing or unpacking; `plt.hist(...)` unpacks into three of them."""

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Inert()

    def __call__(self, *a, **k):
        return _Inert()

    def __iter__(self):
        return iter((_Inert(), _Inert(), _Inert()))

    def __getitem__(self, k):
        return _Inert()

    def __setitem__(self, k, v):
        pass


def _inert_module(name: str) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__getattr__ = lambda attr: _Inert()  # module-level __getattr__ (PEP 562)
    m.__path__ = []
    return m


def install_matplotlib_stubs() -> bool:
    """Registers inert `matplotlib` / `mpl_toolkits` modules unless the real ones import.  Returns True if stubbed."""
    try:
        import matplotlib  # noqa: F401
        return False
    except Exception:
        pass
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.mlab", "matplotlib.cm", "matplotlib.colors",
                 "matplotlib.collections", "mpl_toolkits", "mpl_toolkits.mplot3d"):
        sys.modules.setdefault(name, _inert_module(name))
    return True


# ------------------------------------------------------------------ numpy arguments removed since the layer was written
class _NumpyCompat:
    """`np` as the layer sees it: today's numpy, with the `normed=` keyword of histogram / histogram2d (removed in
    numpy 1.24) mapped to `density=`."""

    def __init__(self):
        import numpy
        self._np = numpy

    def __getattr__(self, name):
        return getattr(self._np, name)

    @staticmethod
    def _kw(k):
        if "normed" in k:
            k = dict(k)
            k["density"] = k.pop("normed")
        return k

    def histogram(self, *a, **k):
        return self._np.histogram(*a, **self._kw(k))

    def histogram2d(self, *a, **k):
        return self._np.histogram2d(*a, **self._kw(k))


def _patch_numpy(ns: dict) -> None:
    import numpy
    for key, val in list(ns.items()):
        if val is numpy:
            ns[key] = _NumpyCompat()


# ------------------------------------------------------------------ importing the reference's modules
class _Py2Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """Imports `<dir>/<name>.py` of a reference pylib directory through compile_py2."""

    def __init__(self, directory: str):
        self.directory = os.path.abspath(directory)

    def find_spec(self, fullname, path=None, target=None):
        if "." in fullname:
            return None
        candidate = os.path.join(self.directory, fullname + ".py")
        if os.path.exists(candidate):
            return importlib.util.spec_from_loader(fullname, self, origin=candidate)
        return None

    def create_module(self, spec):
        return None

    def exec_module(self, module):
        path = module.__spec__.origin
        module.__file__ = path
        _py2_namespace(module.__dict__)
        with open(path, encoding="latin-1") as f:
            code = compile_py2(f.read(), path)
        exec(code, module.__dict__)
        _patch_numpy(module.__dict__)


def _bind_backend():
    from hpsim_b200 import HPSim
    sys.modules["HPSim"] = HPSim
    return HPSim


def load_pylib(pylib_dir: str):
    """The reference's `pylib/hpsim.py`, imported as module `hpsim` on top of hpsim_b200.HPSim."""
    if not os.path.exists(os.path.join(pylib_dir, "hpsim.py")):
        raise FileNotFoundError(f"{pylib_dir}: no hpsim.py (expected the reference's pylib directory)")
    install_matplotlib_stubs()
    _bind_backend()
    if not any(isinstance(f, _Py2Finder) and f.directory == os.path.abspath(pylib_dir) for f in sys.meta_path):
        sys.meta_path.insert(0, _Py2Finder(pylib_dir))
    import importlib
    return importlib.import_module("hpsim")


def run_script(path: str, pylib_dir: str | None = None, argv=None) -> dict:
    """Executes one of the reference's Python-2 user scripts (pytest/*.py) as `__main__` with the working directory
    set to the script's directory -- exactly how the reference documents running them.  Returns the script's globals."""
    path = os.path.abspath(path)
    if pylib_dir is None:
        pylib_dir = os.path.join(os.path.dirname(os.path.dirname(path)), "pylib")
    load_pylib(pylib_dir)
    ns = _py2_namespace({"__name__": "__main__", "__file__": path, "__builtins__": builtins})
    ns["exit"] = ns["quit"] = _ScriptExit.throw
    old_cwd, old_argv = os.getcwd(), sys.argv
    os.chdir(os.path.dirname(path))
    sys.argv = [path] + list(argv or [])
    try:
        with open(path, encoding="latin-1") as f:
            code = compile_py2(f.read(), path)
        try:
            exec(code, ns)
        except _ScriptExit:
            pass
        _patch_numpy(ns)
    finally:
        os.chdir(old_cwd)
        sys.argv = old_argv
    return ns


class _ScriptExit(Exception):
    @staticmethod
    def throw(*_a):
        raise _ScriptExit()
