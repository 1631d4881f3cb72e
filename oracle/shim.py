"""TEST INFRASTRUCTURE ONLY -- import shim for the *unmodified* reference (antoine311200/Syngular).

Only usable in the build container, where the reference is mounted read-only at
/root/reference.  It is used by `oracle/gen_golden.py` (to produce the committed
fixtures under tests/golden/) and by the CPU tests that cross-check the numpy
restatement `oracle/ref_numpy.py` against the real reference when it is present.
Nothing in the product package (`syngular_b200/`, `syngular/`) may import this.

The reference has no packaging and imports several modules that are absent here
(opt_einsum, django, matplotlib, turtle) -- see SURVEY.md Appendix A.  We stub them:
`opt_einsum.contract` is replaced by a sequential pairwise np.einsum evaluation of the
same interleaved spec (contraction ORDER is the only thing opt_einsum decides; the
arithmetic is numpy's either way).
"""
import os
import sys
import types
import warnings
from collections import Counter

import numpy as np

REFERENCE_ROOT = os.environ.get("SYNGULAR_REFERENCE", "/root/reference")


def _contract(*args, **kw):
    """opt_einsum.contract (interleaved form, arbitrary integer labels) as pairwise np.einsum."""
    ops, subs, out, i = [], [], None, 0
    while i < len(args):
        if isinstance(args[i], np.ndarray):
            ops.append(args[i])
            subs.append([s for s in args[i + 1] if s is not Ellipsis])
            i += 2
        else:
            out = [s for s in args[i] if s is not Ellipsis]
            i += 1
    count = Counter(l for s in subs for l in s)
    if out is None:
        out = sorted(l for l, c in count.items() if c == 1)
    cur, cur_s = ops[0], list(subs[0])
    for k in range(1, len(ops)):
        nxt_s = list(subs[k])
        later = Counter(l for s in subs[k + 1:] for l in s)
        every = list(dict.fromkeys(cur_s + nxt_s))
        keep = [l for l in every if l in out or later[l] > 0]
        m = {l: j for j, l in enumerate(every)}
        cur = np.einsum(cur, [m[l] for l in cur_s], ops[k], [m[l] for l in nxt_s], [m[l] for l in keep], optimize=True)
        cur_s = keep
    m = {l: j for j, l in enumerate(cur_s)}
    if cur_s == out and len(ops) > 1:
        return cur
    return np.einsum(cur, [m[l] for l in cur_s], [m[l] for l in out])


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "tensor"))


def load_reference():
    """Return the reference's `syngular` package (imported from /root/reference), or raise."""
    if not available():
        raise RuntimeError("reference not mounted at %s" % REFERENCE_ROOT)
    if "syngular" in sys.modules and getattr(sys.modules["syngular"], "__graft_reference__", False):
        return sys.modules["syngular"]
    # the product's drop-in package is also called `syngular`; never mix the two in a process
    for name in list(sys.modules):
        if name == "syngular" or name.startswith("syngular."):
            raise RuntimeError("a different `syngular` package is already imported in this process")
    warnings.filterwarnings("ignore")
    _stub("opt_einsum", contract=_contract)
    _stub("django", forms=_stub("django.forms", PasswordInput=object))
    _stub("turtle", right=None)
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib.pyplot  # noqa: F401
        except Exception:
            _stub("matplotlib", pyplot=_stub("matplotlib.pyplot", isinteractive=lambda: False))
    try:
        import cv2  # noqa: F401
    except Exception:
        _stub("cv2", eigen=None)
    if not hasattr(np, "complex"):
        np.complex = complex
    link_dir = os.path.join("/tmp", "syngular_ref_link_%d" % os.getpid())
    os.makedirs(link_dir, exist_ok=True)
    link = os.path.join(link_dir, "syngular")
    if not os.path.islink(link):
        os.symlink(REFERENCE_ROOT, link)
    sys.path.insert(0, link_dir)
    try:
        import syngular  # noqa: F401  (the reference)
        import syngular.tensor  # noqa: F401
    finally:
        sys.path.remove(link_dir)
    sys.modules["syngular"].__graft_reference__ = True
    return sys.modules["syngular"]
