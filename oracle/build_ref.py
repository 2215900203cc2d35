"""Build the UNMODIFIED reference C++ decoder into oracle/_ref/ (test infrastructure only).

The sources are compiled where they lie under /root/reference (never copied
into this repo): /root/reference/src/openpifpaf/csrc/src/*.cpp with
csrc/include on the include path, against the libtorch that ships with the
installed PyTorch.  Output: oracle/_ref/refcpp.so which registers
torch.classes.openpifpaf_decoder.* / openpifpaf_decoder_utils.* exactly as
the reference's own extension does (csrc/src/module.cpp:19-118).

Gotcha: the container exports CXX=/opt/gcc/bin/g++ (a wrapper);
objects built that way crash in the first OPENPIFPAF_INFO print.  We force
/usr/bin/g++.

This file is TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may load oracle/_ref.
"""
import glob
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_CSRC = '/root/reference/src/openpifpaf/csrc'
OUT_DIR = os.path.join(HERE, '_ref')
OUT_SO = os.path.join(OUT_DIR, 'refcpp.so')


def build(force=False, verbose=False):
    if os.path.exists(OUT_SO) and not force:
        return OUT_SO
    if not os.path.isdir(REF_CSRC):
        raise FileNotFoundError(
            f'{REF_CSRC} not present (GPU box?): oracle/_ref must be prebuilt and shipped')
    os.environ['CXX'] = '/usr/bin/g++'
    os.environ['CC'] = '/usr/bin/gcc'
    from torch.utils.cpp_extension import load
    build_dir = os.path.join(OUT_DIR, 'build')
    os.makedirs(build_dir, exist_ok=True)
    load(name='refcpp',
         sources=sorted(glob.glob(os.path.join(REF_CSRC, 'src', '*.cpp'))),
         extra_include_paths=[os.path.join(REF_CSRC, 'include')],
         extra_cflags=['-std=c++17', '-O2', '-D_SILENCE_ALL_CXX17_DEPRECATION_WARNINGS'],
         build_directory=build_dir, is_python_module=False, verbose=verbose)
    shutil.copy(os.path.join(build_dir, 'refcpp.so'), OUT_SO)
    shutil.rmtree(build_dir, ignore_errors=True)
    return OUT_SO


REF_PY = '/root/reference/src/openpifpaf'
PKG_DIR = os.path.join(HERE, '_ref_pkg')


def stage_package(force=False):
    """Stage the UNMODIFIED reference Python package next to its compiled extension so that it travels to the GPU
    box (like oracle/_ref, this directory is git-ignored but not gpurun-ignored; nothing of it enters the history):

      oracle/_ref_pkg/openpifpaf/        <- copy of /root/reference/src/openpifpaf (csrc/ sources left out)
      oracle/_ref_pkg/openpifpaf/_cpp.so <- oracle/_ref/refcpp.so (what cpp_extension.py:6-26 looks for)
      oracle/_ref_pkg/pysparkling.py     <- one-line stub of the optional dependency (gotcha 2)

    Users: the -m gpu plugin test (the reference's own Predictor / Multi.batch with CifCafB200 selected) and
    bench.py --impl reference (the reference's own Shell + Decoder.batch on the host cores)."""
    marker = os.path.join(PKG_DIR, 'openpifpaf', '_cpp.so')
    if os.path.exists(marker) and not force:
        return PKG_DIR
    if not os.path.isdir(REF_PY):
        raise FileNotFoundError(f'{REF_PY} not present (GPU box?): oracle/_ref_pkg must be staged beforehand')
    so = build()
    shutil.rmtree(PKG_DIR, ignore_errors=True)
    shutil.copytree(REF_PY, os.path.join(PKG_DIR, 'openpifpaf'),
                    ignore=shutil.ignore_patterns('csrc', '__pycache__', '*.pyc'))
    shutil.copy(so, marker)
    with open(os.path.join(PKG_DIR, 'pysparkling.py'), 'w') as f:
        f.write('class Context:\n    pass\n')
    return PKG_DIR


def package_available():
    return os.path.exists(os.path.join(PKG_DIR, 'openpifpaf', '_cpp.so'))


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))
    print(stage_package(force='--force' in sys.argv))
