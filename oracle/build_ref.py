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


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))
