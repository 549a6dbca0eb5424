"""Import the UNMODIFIED reference Python hot-path modules under this interpreter.

TEST INFRASTRUCTURE (build container only; reads /root/reference).  Used by oracle/make_golden.py to
produce tests/golden/* and by bench.py --impl reference when /root/reference happens to exist.

`import compressai` from /root/reference fails here (cp38 binaries on py3.12; timm / detectron2 /
pytorch_msssim absent; compressai/models/__init__.py imports detectron2-dependent models).  The
hot-path files import fine through a scratch package that symlinks them unmodified and supplies three
shims:
  * compressai/__init__.py  -- the coder registry of /root/reference/compressai/__init__.py:22-62
  * timm.models.layers      -- to_2tuple / trunc_normal_ / DropPath (stf.py:5)
  * compressai.ans / compressai._CXX -- the reference's own binaries through oracle/refbin.py
    (or oracle/coder.py, the pinned C restatement, when `coder="port"`).
"""
import importlib
import os
import sys
import tempfile

REF = "/root/reference"

_INIT = '''
_entropy_coder = "ans"
_available_entropy_coders = [_entropy_coder]
def set_entropy_coder(entropy_coder):
    global _entropy_coder
    if entropy_coder not in _available_entropy_coders:
        raise ValueError(f'Invalid entropy coder "{entropy_coder}"')
    _entropy_coder = entropy_coder
def get_entropy_coder():
    return _entropy_coder
def available_entropy_coders():
    return _available_entropy_coders
'''

_TIMM = '''
import collections.abc, itertools, torch
from torch.nn.init import trunc_normal_
def to_2tuple(x):
    if isinstance(x, collections.abc.Iterable) and not isinstance(x, str):
        return tuple(x)
    return tuple(itertools.repeat(x, 2))
class DropPath(torch.nn.Module):
    def __init__(self, drop_prob=0.0):
        super().__init__(); self.drop_prob = drop_prob
    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        return x * mask / keep
'''

_ANS = '''
import numpy as np
from {backend} import RansEncoder as _E, BufferedRansEncoder as _B, RansDecoder as _D
def _t(cdfs): return np.asarray(cdfs, dtype=np.int32)
class RansEncoder:
    def __init__(self): self._e = _E()
    def encode_with_indexes(self, s, i, c, l, o): return self._e.encode_with_indexes(s, i, _t(c), l, o)
class BufferedRansEncoder:
    def __init__(self): self._e = _B()
    def encode_with_indexes(self, s, i, c, l, o): return self._e.encode_with_indexes(s, i, _t(c), l, o)
    def flush(self): return self._e.flush()
class RansDecoder:
    def __init__(self): self._d = _D()
    def set_stream(self, s): return self._d.set_stream(s)
    def decode_stream(self, i, c, l, o): return self._d.decode_stream(i, _t(c), l, o).tolist()
    def decode_with_indexes(self, s, i, c, l, o): return self._d.decode_with_indexes(s, i, _t(c), l, o).tolist()
'''

_CXX = '''
from {backend} import pmf_to_quantized_cdf as _f
def pmf_to_quantized_cdf(pmf, precision): return [int(v) for v in _f(pmf, precision)]
'''


def available():
    return os.path.isdir(os.path.join(REF, "compressai"))


def install(coder="binary"):
    """Create the scratch package, put it FIRST on sys.path and return the imported modules.

    coder = "binary": reference's shipped .so via oracle/refbin.py; "port": oracle/coder.py.
    """
    assert available(), "/root/reference is not present on this machine"
    for name in list(sys.modules):
        if name == "compressai" or name.startswith("compressai.") or name == "timm" or name.startswith("timm."):
            del sys.modules[name]
    root = tempfile.mkdtemp(prefix="refshim_")
    pkg = os.path.join(root, "compressai")
    os.makedirs(os.path.join(pkg, "models"))
    os.makedirs(os.path.join(root, "timm", "models"))
    src = os.path.join(REF, "compressai")
    for d in ("entropy_models", "layers", "ops"):
        os.symlink(os.path.join(src, d), os.path.join(pkg, d))
    for f in ("base.py", "utils.py", "stf.py", "cnn.py"):
        os.symlink(os.path.join(src, "models", f), os.path.join(pkg, "models", f))
    open(os.path.join(pkg, "models", "__init__.py"), "w").close()
    backend = "oracle.refbin" if coder == "binary" else "oracle.coder"
    with open(os.path.join(pkg, "__init__.py"), "w") as f:
        f.write(_INIT)
    with open(os.path.join(pkg, "ans.py"), "w") as f:
        f.write(_ANS.format(backend=backend))
    with open(os.path.join(pkg, "_CXX.py"), "w") as f:
        f.write(_CXX.format(backend=backend))
    open(os.path.join(root, "timm", "__init__.py"), "w").close()
    open(os.path.join(root, "timm", "models", "__init__.py"), "w").close()
    with open(os.path.join(root, "timm", "models", "layers.py"), "w") as f:
        f.write(_TIMM)
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if repo not in sys.path:
        sys.path.append(repo)  # for `oracle.*`
    sys.path.insert(0, root)
    importlib.invalidate_caches()
    stf = importlib.import_module("compressai.models.stf")
    cnn = importlib.import_module("compressai.models.cnn")
    em = importlib.import_module("compressai.entropy_models")
    return {"root": root, "stf": stf, "cnn": cnn, "entropy_models": em}


def uninstall(handle):
    if handle["root"] in sys.path:
        sys.path.remove(handle["root"])
    for name in list(sys.modules):
        if name == "compressai" or name.startswith("compressai.") or name == "timm" or name.startswith("timm."):
            del sys.modules[name]
