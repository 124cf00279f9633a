"""TEST INFRASTRUCTURE ONLY — never imported by the product path.

Loads the *unmodified* reference (`/root/reference/model.py`, Python 2.7 / torch 0.1.12 code) in this
container (Python 3.12 / torch 2.11) so that its own `Sender` / `Receiver` / `Baseline` / `exchange` /
loss functions and its per-iteration update block (`model.py:1243-1330`) can generate golden vectors.

Nothing from the reference is copied into this repo: the module is imported from where it lies and the
update block is read from the file at run time (`reference_update_block`).  `/root/reference` exists
only in the build container, so only `tests/golden/make_golden.py` (fixture generator) and the
`not gpu` cross-check tests (skipped when the path is absent) use this file.

The shim restores the torch-0.1.12 / py2 semantics the reference relies on:

  (i)   `t.sum(dim)` / `t.max(dim)` keep the reduced dim (no implicit squeeze)      model.py:395,449,911
  (ii)  truthiness of a Variable is "non-empty" (py2 `__len__`), not its value         model.py:123,423,838
  (iii) `.data[0]` on a scalar result returns the Python number                        model.py:866,947
  (iv)  `map` is eager (py2 returns a list)                                            model.py:886,956
  (v)   uint8 masks are valid for `masked_select` / advanced indexing                  model.py:896,941-944
  (vi)  `nn.utils.clip_grad_norm` exists (renamed `clip_grad_norm_` later)             model.py:1310
  (vii) `misc.xavier_normal` must not treat every Tensor as a Variable                 misc.py:379-381
  (viii) `gflags` -> `absl.flags`; `h5py`, `nltk`, `visdom`, `parse` are stubbed (unused on the path)

Patches (i)-(iii),(v) are active only inside `legacy_semantics()`.
"""
import builtins
import contextlib
import importlib
import os
import sys
import types
import warnings

import numpy as np
import torch

REFERENCE_PATH = os.environ.get("MMG_REFERENCE_PATH", "/root/reference")


def reference_available(path=None):
    return os.path.isfile(os.path.join(path or REFERENCE_PATH, "model.py"))


# --------------------------------------------------------------------------------------------
# module stubs
# --------------------------------------------------------------------------------------------
def _install_stubs():
    from absl import flags as absl_flags

    if "gflags" not in sys.modules:
        sys.modules["gflags"] = absl_flags
    for name in ("h5py", "visdom", "parse"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    try:
        importlib.import_module("nltk.tokenize")
        importlib.import_module("nltk.corpus")
    except Exception:
        nltk = types.ModuleType("nltk")
        tok = types.ModuleType("nltk.tokenize")
        tok.word_tokenize = lambda s: s.split()
        corp = types.ModuleType("nltk.corpus")

        class _SW(object):
            @staticmethod
            def words(_lang):
                return []

        corp.stopwords = _SW()
        nltk.tokenize, nltk.corpus = tok, corp
        sys.modules["nltk"] = nltk
        sys.modules["nltk.tokenize"] = tok
        sys.modules["nltk.corpus"] = corp


_REF = {}


def load_reference(path=None):
    """Import the reference `model` module (cached).  Returns the module object."""
    path = path or REFERENCE_PATH
    if path in _REF:
        return _REF[path]
    if not reference_available(path):
        raise FileNotFoundError("reference not present at %s" % path)
    _install_stubs()
    sys.path.insert(0, path)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            misc = importlib.import_module("misc")
            # (vii) every Tensor is a `Variable` now -> infinite recursion in misc.xavier_normal
            misc.Variable = type("LegacyVariable", (), {})
            model = importlib.import_module("model")
    finally:
        sys.path.remove(path)
    if not hasattr(model.FLAGS, "use_binary") or "use_binary" not in model.FLAGS:
        model.flags()
    if not model.FLAGS.is_parsed():
        model.FLAGS(["ref_shim"])
    # (iv) eager map inside the reference module only
    model.map = lambda *a: list(builtins.map(*a))
    # (vi)
    if not hasattr(torch.nn.utils, "clip_grad_norm"):
        torch.nn.utils.clip_grad_norm = torch.nn.utils.clip_grad_norm_
    # absl spells it flag_values_dict
    _REF[path] = model
    return model


def set_flags(model, **kw):
    """Set reference FLAGS (the reference reads the global FLAGS inside ctor/forward)."""
    for k, v in kw.items():
        setattr(model.FLAGS, k, v)


# --------------------------------------------------------------------------------------------
# torch 0.1.12 semantics
# --------------------------------------------------------------------------------------------
@contextlib.contextmanager
def legacy_semantics():
    T = torch.Tensor
    saved = dict(sum=T.sum, max=T.max, bool=T.__bool__, getitem=T.__getitem__,
                 tsum=torch.sum, tmax=torch.max, msel=torch.masked_select)

    def _sum(self, *a, **k):
        if a and isinstance(a[0], int) and len(a) == 1 and "keepdim" not in k:
            return saved["sum"](self, a[0], keepdim=True, **k)
        return saved["sum"](self, *a, **k)

    def _max(self, *a, **k):
        if a and isinstance(a[0], int) and len(a) == 1 and "keepdim" not in k:
            return saved["max"](self, a[0], keepdim=True, **k)
        return saved["max"](self, *a, **k)

    def _bool(self):
        if self.dim() >= 1:
            return self.shape[0] != 0
        return saved["bool"](self)

    def _getitem(self, idx):
        if self.dim() == 0 and isinstance(idx, int) and idx == 0:
            return self.item()
        if isinstance(idx, torch.Tensor) and idx.dtype == torch.uint8:
            idx = idx.bool()
        return saved["getitem"](self, idx)

    def _msel(inp, mask, **k):
        if mask.dtype == torch.uint8:
            mask = mask.bool()
        return saved["msel"](inp, mask, **k)

    T.sum, T.max, T.__bool__, T.__getitem__ = _sum, _max, _bool, _getitem
    torch.masked_select = _msel
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            yield
    finally:
        T.sum, T.max, T.__bool__, T.__getitem__ = saved["sum"], saved["max"], saved["bool"], saved["getitem"]
        torch.masked_select = saved["msel"]


# --------------------------------------------------------------------------------------------
# numpy RNG recording: the reference samples with np.random.rand (model.py:227,420,460)
# --------------------------------------------------------------------------------------------
@contextlib.contextmanager
def record_uniforms(seed, sink):
    """Seed numpy's global RNG and append every `np.random.rand` draw (float64 array) to `sink`."""
    orig = np.random.rand
    np.random.seed(seed)

    def _rand(*shape):
        u = orig(*shape)
        sink.append(np.array(u, copy=True))
        return u

    np.random.rand = _rand
    try:
        yield
    finally:
        np.random.rand = orig


def reference_update_block(path=None):
    """Return the reference's per-iteration update block (`model.py:1243-1330`, from the unpacking of
    `exchange`'s result to the last optimizer step) as dedented source text, read from the reference file
    at run time.  Executed by tests/golden/make_golden.py inside `legacy_semantics()`."""
    import textwrap
    path = path or REFERENCE_PATH
    with open(os.path.join(path, "model.py")) as f:
        lines = f.read().split("\n")
    start = next(i for i, l in enumerate(lines) if l.strip() == "s_masks, s_feats, s_probs = s" and i > 1200)
    end = next(i for i, l in enumerate(lines) if l.strip() == "optimizer_bas_sen.step()" and i > start)
    return textwrap.dedent("\n".join(lines[start:end + 1]))
