"""Build a variant of libmkidgpu.so for kernel experiments: ONE source recompiled with extra -D flags, linked with the
objects of the regular build.

    python scripts/build_variant.py NAME channelize.cu -DK4_NAMED_BAR [...]
        -> mkids_sdr_b200/build/variants/libmkidgpu_NAME.so     (use with MKIDGPU_LIB=...)
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mkids_sdr_b200 import build as B  # noqa: E402


def main():
    name, src = sys.argv[1], sys.argv[2]
    extra = sys.argv[3:]
    B.build()
    objdir = os.path.join(B.HERE, 'build')
    vdir = os.path.join(objdir, 'variants')
    os.makedirs(vdir, exist_ok=True)
    obj = os.path.join(vdir, '%s_%s.o' % (src[:-3], name))
    r = subprocess.run([B.NVCC] + extra + B.FLAGS + ['-c', os.path.join(B.CSRC, src), '-o', obj],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    open(os.path.join(vdir, 'ptxas_%s.log' % name), 'w').write(r.stdout)
    if r.returncode:
        sys.stderr.write(r.stdout)
        raise SystemExit(1)
    objs = [obj if f == src[:-3] + '.o' else os.path.join(objdir, f) for f in sorted(os.listdir(objdir)) if f.endswith('.o')]
    out = os.path.join(vdir, 'libmkidgpu_%s.so' % name)
    subprocess.check_call([B.NVCC, '-shared', '-cudart', 'shared', '-o', out] + objs +
                          ['-gencode', 'arch=compute_100a,code=sm_100a', '-Xlinker', '-rpath', '-Xlinker', '/usr/local/cuda/lib64', '-ldl'])
    print(out)


if __name__ == '__main__':
    main()
