"""Build oracle/_ref/ from the reference where it lies.  TEST INFRASTRUCTURE ONLY.

    python oracle/build_ref.py            (also called by __graft_entry__.build())

The reference (weilonghu/KGC-GCN) is pure Python; its four modules are byte-compiled from /root/reference into
``oracle/_ref/{model,data_loader,main,utils}.refbc`` (CPython bytecode: a BUILT artefact like a .so - git-ignored, it
travels to the GPU box with the snapshot; no reference source is copied into the repository).  With ``oracle/shims`` on
the path (stand-ins for the reference's absent third-party imports) they import unmodified, which is what
``bench.py --impl reference`` times (cpu_baseline.kind = "reference") and what tests/test_oracle_golden.py cross-checks
the port against when the directory is present.  /root/reference does not exist on the GPU box: there the prebuilt
files are used as they are; when they are missing everything falls back to the pinned port (oracle/mgcn_oracle.py).
"""
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, '_ref')
MODULES = ('utils', 'model', 'data_loader', 'main')      # import order: main imports the other three
EXT = '.refbc'                 # not '.pyc': snapshot tools commonly drop *.pyc


def build_ref(ref_dir=None):
    """-> list of written files ([] when the reference tree is absent)."""
    ref_dir = ref_dir or os.environ.get('KGC_REFERENCE_DIR', '/root/reference')
    if not all(os.path.exists(os.path.join(ref_dir, m + '.py')) for m in MODULES):
        return []
    os.makedirs(OUT, exist_ok=True)
    written = []
    for m in MODULES:
        dst = os.path.join(OUT, m + EXT)
        py_compile.compile(os.path.join(ref_dir, m + '.py'), cfile=dst, dfile='reference/' + m + '.py', doraise=True)
        written.append(dst)
    return written


def load_reference():
    """(model, data_loader, main) modules of the UNMODIFIED reference from oracle/_ref, or None when it was not built
    (or was built by another interpreter version).  Puts oracle/shims on sys.path and registers the four modules under
    the names the reference imports them by (utils, model, data_loader, main)."""
    import importlib.machinery
    import importlib.util
    paths = {m: os.path.join(OUT, m + EXT) for m in MODULES}
    if not all(os.path.exists(p) for p in paths.values()):
        return None
    shims = os.path.join(HERE, 'shims')
    if shims not in sys.path:
        sys.path.insert(0, shims)
    loaded = {}
    try:
        for m in MODULES:
            cur = sys.modules.get(m)
            if cur is not None and getattr(cur, '__file__', None) == paths[m]:
                loaded[m] = cur
                continue
            if cur is not None:
                return None                     # some other 'model' / 'main' / 'utils' module is already imported
            loader = importlib.machinery.SourcelessFileLoader(m, paths[m])
            spec = importlib.util.spec_from_loader(m, loader, origin=paths[m])
            mod = importlib.util.module_from_spec(spec)
            mod.__file__ = paths[m]
            sys.modules[m] = mod
            loader.exec_module(mod)
            loaded[m] = mod
    except Exception:                           # stale bytecode (other interpreter), missing third-party module, ...
        for m in MODULES:
            if m in sys.modules and getattr(sys.modules[m], '__file__', None) == paths[m]:
                del sys.modules[m]
        return None
    return loaded['model'], loaded['data_loader'], loaded['main']


if __name__ == '__main__':
    print(build_ref() or 'reference tree not found: nothing built')
