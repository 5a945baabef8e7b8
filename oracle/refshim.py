"""TEST INFRASTRUCTURE ONLY -- in-memory loader for the *unmodified* pysfm reference.

The reference (``/root/reference``) is package-less Python 2.  It cannot be imported by the
Python 3.12 interpreter in this image, it cannot travel to the GPU box, and none of its source
may be copied into this repository.  This module installs a ``sys.meta_path`` finder that
reads a reference module's text from ``/root/reference/<name>.py``, applies purely syntactic
py2->py3 transforms (no arithmetic is touched) and ``exec``s the result in memory:

  1. backslash line continuations are joined;
  2. ``print X`` statements -> ``print(X)``;
  3. ``raise E, msg`` -> ``raise E(msg)``;
  4. ``.viewkeys()/.iteritems()/xrange`` -> py3 spellings, ``import StringIO`` -> ``io``;
  5. each module's globals get ``reduce`` and list-returning ``map``/``zip``.

It is used ONLY by ``oracle/make_golden.py`` (to generate ``tests/golden/*.npz``) and by
``-m "not gpu"`` tests that re-validate the numpy restatement in ``oracle/ba_oracle.py``
when ``/root/reference`` happens to be present.  Nothing under ``pysfm_b200/``, ``bench.py``
or the ``-m gpu`` tests may import it.
"""
import functools
import importlib.abc
import importlib.util
import os
import re
import sys
import types

REFERENCE_ROOT = os.environ.get("PYSFM_REFERENCE_ROOT", "/root/reference")

# Reference modules reachable from the bundle-adjustment path and its unit tests.
_SHIMMED = {
    "algebra", "lie", "sensor_model", "triangulate", "bundle", "optimize", "schur",
    "bundle_adjuster", "bundle_io", "numpy_test", "finite_differences",
    "bundle_unittest", "bundle_adjuster_unittest", "finite_differences_unittest",
    "test_bundle", "synthetic_data", "sequence", "geometry", "window_slam",
}
# matplotlib-only modules (absent in this image): every attribute is a no-op callable
_STUBBED = {"draw_bundle", "draw_bundle_pca", "matplotlib", "matplotlib.pyplot"}


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "bundle_adjuster.py"))


_PRINT_RE = re.compile(r"^(\s*)print\b(?!\s*\()(.*)$")
_PRINT_PAREN_RE = re.compile(r"^(\s*)print\s*(\(.*\))\s*%\s*(.*)$")
_RAISE_RE = re.compile(r"^(\s*)raise\s+([A-Za-z_][\w.]*)\s*,\s*(.+)$")


def _split_trailing_comment(code):
    """Split ``code`` into (statement, comment) ignoring '#' inside string literals."""
    quote = None
    i = 0
    while i < len(code):
        c = code[i]
        if quote:
            if c == "\\":
                i += 2
                continue
            if c == quote:
                quote = None
        elif c in "'\"":
            quote = c
        elif c == "#":
            return code[:i], code[i:]
        i += 1
    return code, ""


def py2_to_py3(text):
    text = text.replace("\\\n", " ")
    out = []
    for line in text.split("\n"):
        m = _PRINT_RE.match(line)
        if m:
            indent, rest = m.group(1), m.group(2)
            stmt, comment = _split_trailing_comment(rest)
            stmt = stmt.strip()
            if stmt.endswith(","):
                stmt = stmt[:-1]
            line = "%sprint(%s) %s" % (indent, stmt, comment)
        else:
            m = _PRINT_PAREN_RE.match(line)
            if m:  # print (fmt) % args  ->  print((fmt) % args)
                line = "%sprint(%s %% %s)" % m.groups()
        m = _RAISE_RE.match(line)
        if m:
            indent, exc, rest = m.groups()
            stmt, comment = _split_trailing_comment(rest)
            line = "%sraise %s(%s) %s" % (indent, exc, stmt.strip(), comment)
        out.append(line)
    text = "\n".join(out)
    text = text.replace(".viewkeys()", ".keys()").replace(".iteritems()", ".items()")
    text = text.replace(".itervalues()", ".values()")
    text = re.sub(r"\bxrange\(", "range(", text)
    text = re.sub(r"^(\s*)import StringIO\s*$", r"\1import io as StringIO", text, flags=re.M)
    return text


def _py2_builtins():
    return {
        "reduce": functools.reduce,
        "map": lambda *a: list(map(*a)),
        "zip": lambda *a: list(zip(*a)),
    }


class _RefLoader(importlib.abc.Loader):
    def __init__(self, name, path):
        self.name, self.path = name, path

    def create_module(self, spec):
        return None

    def exec_module(self, module):
        with open(self.path, "r") as f:
            src = py2_to_py3(f.read())
        module.__dict__.update(_py2_builtins())
        module.__file__ = self.path
        code = compile(src, self.path, "exec")
        exec(code, module.__dict__)
        # "from numpy import *" inside the module re-binds map/zip-free names only, but a
        # star import of numpy can shadow our list-returning helpers -- restore them.
        module.__dict__.update(_py2_builtins())


class _StubLoader(importlib.abc.Loader):
    def create_module(self, spec):
        mod = types.ModuleType(spec.name)
        mod.__getattr__ = lambda name: (lambda *a, **k: None)
        mod.__path__ = []   # lets "import matplotlib.pyplot" treat the stub as a package
        return mod

    def exec_module(self, module):
        pass


class _RefFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path=None, target=None):
        if fullname in _STUBBED:
            return importlib.util.spec_from_loader(fullname, _StubLoader())
        if fullname in _SHIMMED:
            p = os.path.join(REFERENCE_ROOT, fullname + ".py")
            if os.path.isfile(p):
                return importlib.util.spec_from_loader(fullname, _RefLoader(fullname, p))
        return None


_finder = None


def install():
    """Make ``import bundle_adjuster`` etc. resolve to the shim-loaded reference."""
    global _finder
    if not available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    if _finder is None:
        _finder = _RefFinder()
        sys.meta_path.insert(0, _finder)
    return _finder


def uninstall():
    global _finder
    if _finder is not None:
        sys.meta_path.remove(_finder)
        _finder = None
    for name in list(_SHIMMED | _STUBBED):
        sys.modules.pop(name, None)


def load(name):
    """Import one reference module through the shim and return it."""
    install()
    return importlib.import_module(name)
