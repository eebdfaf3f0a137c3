"""Builds mdgat-matcher_b200/lib/libmdgat_b200.so from csrc/*.cu for sm_100a with nvcc.

In-tree build: the .so is git-ignored but travels to the GPU box with the gpurun snapshot.
nvcc cross-compiles without a GPU. Rebuilds only translation units whose sources changed.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIBDIR = os.path.join(HERE, 'lib')
OBJDIR = os.path.join(HERE, 'build')
LIB = os.path.join(LIBDIR, 'libmdgat_b200.so')
SOURCES = ['gemm_f64.cu', 'attention_f64.cu', 'attention_bwd.cu', 'sinkhorn.cu', 'sinkhorn_bwd.cu', 'match.cu', 'misc.cu', 'registration.cu', 'prepare.cu', 'ozaki_gemm.cu', 'attention_i8.cu', 'capi.cu']
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
         '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr']


def _digest(paths):
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, 'rb') as f:
            h.update(f.read())
    h.update(' '.join(FLAGS).encode())
    return h.hexdigest()


def build(verbose=False, force=False):
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.h', '.cuh'))]
    headers.append(os.path.join(os.path.dirname(HERE), 'include', 'mdgat_b200.h'))
    jobs, objs = [], []
    for src in SOURCES:
        sp = os.path.join(CSRC, src)
        obj = os.path.join(OBJDIR, src[:-3] + '.o')
        stamp = obj + '.sha'
        dig = _digest([sp] + headers)
        objs.append(obj)
        if not force and os.path.isfile(obj) and os.path.isfile(stamp) and open(stamp).read() == dig:
            continue
        jobs.append((sp, obj, stamp, dig))

    def compile_one(job):
        sp, obj, stamp, dig = job
        cmd = [NVCC] + FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', sp, '-o', obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('nvcc failed for %s:\n%s\n%s' % (sp, r.stdout, r.stderr))
        if verbose:
            print(r.stderr)
        with open(stamp, 'w') as f:
            f.write(dig)

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        list(ex.map(compile_one, jobs))
    if jobs or not os.path.isfile(LIB):
        cmd = [NVCC, '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a', '-lcudart_static',
                                                    '-Xcompiler', '-fPIC']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('link failed:\n%s\n%s' % (r.stdout, r.stderr))
    return LIB


if __name__ == '__main__':
    print(build(verbose='-v' in sys.argv, force='-f' in sys.argv))
