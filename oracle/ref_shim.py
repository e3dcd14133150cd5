"""Oracle (test infrastructure): import the UNMODIFIED reference from /root/reference.

Only usable in the build container (the GPU box has no /root/reference); it is used by
`tests/golden/make_golden.py` to pin the oracle restatement against the reference itself and to
generate the committed golden vectors.  Nothing is copied from the reference: it is imported from
where it lies, with the three shims SURVEY.md section 8(c) lists:

  1. stub `matplotlib` / `matplotlib.pyplot` modules (vit_model.py:11; not installed here),
  2. a scratch CWD holding a generated `palette.json` (vit_model.py:204-210 reads it at import;
     the 256-entry VOC colour map is produced with the bit-interleave rule of predict.py:31-48),
  3. on a CPU-only host `torch.Tensor.cuda` is patched to the identity (vit_model.py:331,348,368
     call `.cuda()` unconditionally).
"""
from __future__ import annotations

import json
import os
import sys
import tempfile
import types

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_reference() -> str:
    """/root/reference in the build container; on the GPU box the byte-identical copy that oracle/make_ref.py wrote into the
    git-ignored oracle/_ref/ (only vit_model.py travels: enough to run the reference model, not make_golden.py)."""
    env = os.environ.get("VTC_REFERENCE_DIR")
    if env:
        return env
    for cand in ("/root/reference", os.path.join(_HERE, "_ref")):
        if os.path.isfile(os.path.join(cand, "vit_model.py")):
            return cand
    return "/root/reference"


REFERENCE_DIR = _find_reference()


def voc_palette():
    """VOC colour map: bit j of the class index goes to bit (7-j) of r/g/b, 3 bits per round."""
    pal = {}
    for i in range(256):
        c, r, g, b = i, 0, 0, 0
        for j in range(8):
            r |= ((c >> 0) & 1) << (7 - j)
            g |= ((c >> 1) & 1) << (7 - j)
            b |= ((c >> 2) & 1) << (7 - j)
            c >>= 3
        pal[str(i)] = [r, g, b]
    return pal


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "vit_model.py"))


def import_reference():
    """Returns the reference `vit_model` module (imported once)."""
    if "vit_model" in sys.modules and getattr(sys.modules["vit_model"], "__file__", "").startswith(REFERENCE_DIR):
        return sys.modules["vit_model"]
    if not available():
        raise RuntimeError(f"reference not found at {REFERENCE_DIR} (only present in the build container)")
    import torch

    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
            import matplotlib.pyplot  # noqa: F401
        except Exception:
            m = types.ModuleType("matplotlib")
            mp = types.ModuleType("matplotlib.pyplot")
            m.pyplot = mp
            sys.modules["matplotlib"] = m
            sys.modules["matplotlib.pyplot"] = mp
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
    scratch = tempfile.mkdtemp(prefix="vtc_ref_")
    with open(os.path.join(scratch, "palette.json"), "w") as f:
        json.dump(voc_palette(), f)
    cwd = os.getcwd()
    popts = torch._tensor_str.PRINT_OPTS.threshold
    os.chdir(scratch)
    sys.path.insert(0, REFERENCE_DIR)
    try:
        saved = sys.modules.pop("vit_model", None)
        import vit_model as ref
        if saved is not None:
            sys.modules["vit_model_product"] = saved
    finally:
        sys.path.remove(REFERENCE_DIR)
        os.chdir(cwd)
        torch.set_printoptions(threshold=popts)
    return ref


class on_cpu:
    """`with ref_shim.on_cpu():` runs the reference model on the host of a machine that HAS a GPU: its unconditional
    `.cuda()` calls (vit_model.py:331,348,368) are the identity inside the block and restored afterwards (the same model
    code then runs unmodified on the GPU for the on-box eager baseline)."""

    def __enter__(self):
        import torch
        self._saved = torch.Tensor.cuda
        torch.Tensor.cuda = lambda t, *a, **k: t
        return self

    def __exit__(self, *exc):
        import torch
        torch.Tensor.cuda = self._saved
        return False


def reference_source_lines(fname: str, first: int, last: int) -> str:
    """Dedented source of reference lines [first, last] (1-based, inclusive), read in place."""
    import textwrap
    with open(os.path.join(REFERENCE_DIR, fname), encoding="utf-8") as f:
        lines = f.readlines()[first - 1:last]
    return textwrap.dedent("".join(lines))
