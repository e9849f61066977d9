"""Build recipe for `oracle/_ref/`: the reference's OWN range coders, compiled from
the sources where they lie under /root/reference (never copied into this repo).

TEST INFRASTRUCTURE ONLY. Outputs go to oracle/_ref/ (git-ignored, NOT gpurun-ignored,
so the built .so travels to the GPU box where /root/reference does not exist).

Sources compiled (unmodified, in place):
  * lib/entropy_models/rans_coder/rans_wrapper.cpp (+ cdf_ops.cpp, rans_byte.h)
      -> oracle/_ref/rans_ext_cpp/rans_ext_cpp.so
  * models/convolutional/lossy_coord_v3/rans_coder/simple_rans_wrapper.cpp
      -> oracle/_ref/simple_rans_ext_cpp/simple_rans_ext_cpp.so
Flags mirror the reference's own JIT recipe
(lib/entropy_models/rans_coder/__init__.py:35-46,
 models/convolutional/lossy_coord_v3/rans_coder/__init__.py:12-24) minus -march=native
(the .so must run on the GPU box's host CPU as well).
"""
import importlib.util
import os
import os.path as osp
import sys

HERE = osp.dirname(osp.abspath(__file__))
REF_ROOT = os.environ.get('FASTPCC_REFERENCE_ROOT', '/root/reference')
OUT = osp.join(HERE, '_ref')

_SPECS = {
    'rans_ext_cpp': dict(
        sources=['lib/entropy_models/rans_coder/rans_wrapper.cpp'],
        includes=['lib/entropy_models/rans_coder'],
        cflags=['-fopenmp', '-O3'], ldflags=[]),
    'simple_rans_ext_cpp': dict(
        sources=['models/convolutional/lossy_coord_v3/rans_coder/simple_rans_wrapper.cpp'],
        includes=['models/convolutional/lossy_coord_v3/rans_coder', 'lib/entropy_models/rans_coder'],
        cflags=['-O3'], ldflags=[]),
}


def so_path(name):
    return osp.join(OUT, name, name + '.so')


def reference_available():
    return all(osp.isfile(osp.join(REF_ROOT, s)) for spec in _SPECS.values() for s in spec['sources'])


def build(verbose=False):
    """Compile both reference extensions into oracle/_ref (no-op for those already built)."""
    if not reference_available():
        return False
    from torch.utils.cpp_extension import load
    for name, spec in _SPECS.items():
        if osp.isfile(so_path(name)):
            continue
        bdir = osp.join(OUT, name)
        os.makedirs(bdir, exist_ok=True)
        load(name=name,
             sources=[osp.join(REF_ROOT, s) for s in spec['sources']],
             extra_include_paths=[osp.join(REF_ROOT, i) for i in spec['includes']],
             extra_cflags=spec['cflags'], extra_ldflags=spec['ldflags'],
             build_directory=bdir, verbose=verbose)
    stage_python()
    return True


PYREF_ZIP = osp.join(OUT, 'pyref.zip')
_PY_TREES = ('lib', 'models', 'scripts')


def stage_python(force=False):
    """Archive the reference's .py files (lib/, models/) into oracle/_ref/pyref.zip so that the drop-in test
    (tests/test_gpu_dropin.py: the UNMODIFIED reference model.py + cuda_ops.py running on this repo's extension
    shim) can run on the GPU box, where /root/reference does not exist.  Same status as a `pip install --target` of
    the reference: git-ignored, travels with the snapshot, never committed, extracted only into a temporary directory."""
    if not osp.isdir(osp.join(REF_ROOT, 'models')):
        return False
    if osp.isfile(PYREF_ZIP) and not force:
        return True
    import zipfile
    os.makedirs(OUT, exist_ok=True)
    with zipfile.ZipFile(PYREF_ZIP, 'w', zipfile.ZIP_DEFLATED) as z:
        for tree in _PY_TREES:
            for root, dirs, files in os.walk(osp.join(REF_ROOT, tree)):
                dirs[:] = [d for d in dirs if d not in ('build', '__pycache__')]
                for f in files:
                    if f.endswith('.py'):
                        full = osp.join(root, f)
                        z.write(full, osp.relpath(full, REF_ROOT))
    return True


def load_ref(name):
    """Import a prebuilt reference extension from oracle/_ref; returns None if absent."""
    path = so_path(name)
    if not osp.isfile(path):
        return None
    import torch  # noqa: F401  (libtorch symbols must be loaded first)
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sys.modules[name] = mod
    return mod


if __name__ == '__main__':
    ok = build(verbose='-v' in sys.argv)
    print('reference built' if ok else 'reference sources not found; nothing built')
    for n in _SPECS:
        print(n, 'OK' if load_ref(n) is not None else 'MISSING')
