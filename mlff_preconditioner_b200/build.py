"""In-tree build of libmlffpc.so (sm_100a only): nvcc -> mlff_preconditioner_b200/libmlffpc.so."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libmlffpc.so')
SOURCES = ['core.cu', 'geometry.cu', 'matvec.cu', 'gemv.cu', 'symop.cu', 'symtma.cu', 'dense.cu', 'gramdd.cu', 'pchol.cu', 'precon.cu', 'pcg.cu', 'peer.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=default', '--use_fast_math=false']


def _nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError('nvcc not found')


STAMP = LIB + '.stamp'


def _source_digest():
    """Content hash of every source the library is built from (mtimes do not survive the copy to a GPU box)."""
    import hashlib

    h = hashlib.sha256(' '.join(NVCC_FLAGS).encode())
    deps = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cu', '.cuh', '.h')))
    deps.append(os.path.join(HERE, '..', 'include', 'mlffpc.h'))
    for d in deps:
        h.update(os.path.basename(d).encode())
        with open(d, 'rb') as f:
            h.update(f.read())
    return h.hexdigest()


def needs_build():
    if not os.path.exists(LIB) or not os.path.exists(STAMP):
        return True
    with open(STAMP) as f:
        return f.read().strip() != _source_digest()


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a into one shared library; returns its path."""
    if not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    objs = []
    flags = [f for f in NVCC_FLAGS if not f.startswith('--use_fast_math')]
    if verbose:
        flags = flags + ['-Xptxas', '-v']
    procs = []
    for src in SOURCES:
        obj = os.path.join(CSRC, src.replace('.cu', '.o'))
        objs.append(obj)
        cmd = [nvcc] + flags + ['-c', os.path.join(CSRC, src), '-o', obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write('--- %s\n%s\n' % (src, out.decode()))
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError('nvcc failed')
    subprocess.check_call([nvcc, '-gencode', 'arch=compute_100a,code=sm_100a', '-shared', '-o', LIB] + objs + ['-ldl'])
    with open(STAMP, 'w') as f:
        f.write(_source_digest())
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
