"""Build libcomic_b200.so in-tree with nvcc for sm_100a.

    python build.py            # incremental (skips when up to date)
    python build.py --force
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, os.environ.get('COMIC_B200_LIBNAME', 'libcomic_b200.so'))   # env: experiment builds only
_OBJ_PREFIX = (os.path.splitext(os.path.basename(LIB))[0] + '_') if 'COMIC_B200_LIBNAME' in os.environ else ''
SOURCES = ['api.cu', 'encoder.cu', 'encoder_train.cu', 'decoder.cu', 'persistent.cu', 'train.cu']
NVCC_FLAGS = ['-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo',
              '-Xcompiler', '-fPIC']


def _nvcc():
    for p in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if p and (os.path.isabs(p) and os.path.exists(p) or not os.path.isabs(p)):
            return p
    return 'nvcc'


def up_to_date():
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(os.path.dirname(HERE), 'include', 'comic_b200.h'))
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force=False, verbose=False):
    if not force and up_to_date():
        return LIB
    objs = []
    procs = []
    for s in SOURCES:
        o = os.path.join(CSRC, _OBJ_PREFIX + s.replace('.cu', '.o'))
        cmd = [_nvcc()] + NVCC_FLAGS + os.environ.get('COMIC_B200_NVCC_EXTRA', '').split() + ['-c', os.path.join(CSRC, s), '-o', o]
        if verbose:
            cmd += ['-Xptxas', '-v']
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        objs.append(o)
    for s, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out.decode())
        if p.returncode:
            raise RuntimeError('nvcc failed on %s' % s)
    cmd = [_nvcc(), '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a']
    subprocess.check_call(cmd)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
