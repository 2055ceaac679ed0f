"""Recipe for oracle/_ref/ (TEST INFRASTRUCTURE ONLY, like everything under oracle/).

The reference is pure Python: "compiling" it means byte-compiling its own sources, where they lie under /root/reference,
into sourceless modules under oracle/_ref/ (git-ignored, NOT gpurun-ignored: the .pyc files travel to the GPU box like the
built libotb200.so; no reference source is copied into the repository).  `bench.py --impl reference` then times the
reference's OWN SpeechToTextRecognizer.recognize on the box's host cores (cpu_baseline.kind = "reference"); without
oracle/_ref it falls back to the oracle port (kind = "port").  Only bench.py's reference arm and the tests may import this.

    python oracle/build_ref.py          # in the build container (needs /root/reference); __graft_entry__.build() calls it
"""
import importlib
import os
import py_compile
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = '/root/reference'
DST = os.path.join(HERE, '_ref')


def build(dst=DST):
    """Byte-compile /root/reference/otrans/**.py -> dst/otrans/**.pyc (legacy sourceless layout).  Returns dst or None."""
    src = os.path.join(REF_ROOT, 'otrans')
    if not os.path.isdir(src):
        return None
    n = 0
    for root, _, files in os.walk(src):
        rel = os.path.relpath(root, REF_ROOT)
        for f in files:
            if not f.endswith('.py'):
                continue
            out = os.path.join(dst, rel, f[:-3] + '.pyc')
            os.makedirs(os.path.dirname(out), exist_ok=True)
            with warnings.catch_warnings():      # the reference has a docstring with an invalid escape (train/__init__.py:1)
                warnings.simplefilter('ignore', SyntaxWarning)
                py_compile.compile(os.path.join(root, f), cfile=out, dfile=os.path.join(rel, f), doraise=True,
                                   invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
            n += 1
    with open(os.path.join(dst, 'BUILD_INFO'), 'w') as fh:
        fh.write(f'{n} modules byte-compiled from {src} with python {sys.version.split()[0]} by oracle/build_ref.py\n')
    return dst


def load(dst=DST):
    """Import the byte-compiled reference package; returns the `otrans` module or None when oracle/_ref is absent / unusable."""
    if not os.path.isfile(os.path.join(dst, 'otrans', '__init__.pyc')):
        return None
    for p in (os.path.join(dst, 'otrans', 'module'), dst):      # ffn.py:9 does `from activation import Swish` (non-package import)
        if p not in sys.path:
            sys.path.insert(0, p)
    try:
        otrans = importlib.import_module('otrans')
        importlib.import_module('otrans.model')
        importlib.import_module('otrans.recognize.speech2text')
        return otrans
    except Exception:
        return None


if __name__ == '__main__':
    out = build()
    print('oracle/_ref built at' if out else 'no /root/reference here: nothing built', out or '')
