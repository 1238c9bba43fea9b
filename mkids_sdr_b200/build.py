"""Build libmkidgpu.so (hand-written sm_100a CUDA + C ABI) in-tree with nvcc.

    python -m mkids_sdr_b200.build [--force] [--verbose]

Cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box with
the gpurun snapshot; nothing is JIT-compiled at run time.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OUT = os.path.join(HERE, 'libmkidgpu.so')
STAMP = os.path.join(HERE, '.libmkidgpu.stamp')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
EXTRA = os.environ.get('MKID_NVCC_EXTRA', '').split()      # e.g. -DK4_... for kernel experiments
FLAGS = EXTRA + ['-std=c++17', '-O3', '-lineinfo', '-gencode', 'arch=compute_100a,code=sm_100a',
         '-Xcompiler', '-fPIC', '-Xcompiler', '-O3', '--expt-relaxed-constexpr', '-Xptxas', '-v', '-cudart', 'shared']


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))


def _digest():
    h = hashlib.sha256()
    inc = os.path.join(os.path.dirname(HERE), 'include', 'mkidgpu.h')
    files = _sources() + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cuh')) + [inc]
    for f in files:
        h.update(f.encode())
        h.update(open(f, 'rb').read())
    h.update(' '.join(FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    dig = _digest()
    if not force and os.path.exists(OUT) and os.path.exists(STAMP) and open(STAMP).read().strip() == dig:
        return OUT
    objdir = os.path.join(HERE, 'build')
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for src in _sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + '.o')
        objs.append(obj)
        cmd = [NVCC] + FLAGS + ['-c', src, '-o', obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, pr in procs:
        out, _ = pr.communicate()
        log.append('== %s\n%s' % (os.path.basename(src), out))
        if pr.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError('nvcc failed on %s' % src)
    open(os.path.join(objdir, 'ptxas.log'), 'w').write('\n'.join(log))
    if verbose:
        print('\n'.join(log))
    # shared cudart: the library must not carry a private copy of the runtime (and of every
    # entry point it names); the rpath is the image's toolkit, torch's copy has the same SONAME
    subprocess.check_call([NVCC, '-shared', '-cudart', 'shared', '-o', OUT] + objs +
                          ['-gencode', 'arch=compute_100a,code=sm_100a',
                           '-Xlinker', '-rpath', '-Xlinker', '/usr/local/cuda/lib64', '-ldl'])
    open(STAMP, 'w').write(dig)
    return OUT


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
