"""In-tree build of libsg_b200.so (nvcc, sm_100a only).  `python -m shapegan_b200.build` or __graft_entry__.build()."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB_DIR = os.path.join(HERE, 'lib')
LIB_PATH = os.path.join(LIB_DIR, 'libsg_b200.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
         '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr', '-Xptxas', '-v']


def sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + glob.glob(os.path.join(CSRC, '*.cuh')) + glob.glob(os.path.join(CSRC, '*.h')) + \
        glob.glob(os.path.join(HERE, '..', 'include', '*.h'))
    return any(os.path.getmtime(d) > t for d in deps)


def build_lib(force=False, verbose=False, trace=False):
    """Compile every .cu under csrc/ to objects (parallel), link the shared library.  Returns the .so path.
    trace: compile the clock64 stamps of tools/trace_igemm.py into the implicit-GEMM kernels (measurement builds only)."""
    if not force and not trace and not _stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    # one builder at a time (torchrun starts N ranks that may all find the library stale): the others wait, then find it fresh
    import fcntl
    lock = open(os.path.join(LIB_DIR, '.build.lock'), 'w')
    fcntl.flock(lock, fcntl.LOCK_EX)
    try:
        if not force and not trace and not _stale():
            return LIB_PATH
        return _build_locked(verbose, trace)
    finally:
        fcntl.flock(lock, fcntl.LOCK_UN)
        lock.close()


def _build_locked(verbose, trace):
    objdir = os.path.join(HERE, 'build')
    os.makedirs(objdir, exist_ok=True)
    procs = []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + '.o')
        cmd = [NVCC] + FLAGS + (['-DSG_IGEMM_TRACE'] if trace else []) + ['-c', src, '-o', obj]
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs, log = [], []
    for src, obj, pr in procs:
        out, _ = pr.communicate()
        log.append('== %s\n%s' % (os.path.basename(src), out))
        if pr.returncode != 0:
            sys.stderr.write('\n'.join(log))
            raise RuntimeError('nvcc failed on %s' % src)
        objs.append(obj)
    with open(os.path.join(objdir, 'ptxas.log'), 'w') as f:
        f.write('\n'.join(log))
    tmp = LIB_PATH + '.tmp.%d' % os.getpid()
    cmd = [NVCC, '-shared', '-o', tmp] + objs + ['-lcudart']
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if out.returncode != 0:
        sys.stderr.write(out.stdout)
        raise RuntimeError('link failed')
    os.replace(tmp, LIB_PATH)                      # atomic: a process that is loading the library never sees a half-written file
    if verbose:
        print('\n'.join(log))
    return LIB_PATH


if __name__ == '__main__':
    print(build_lib(force='--force' in sys.argv, verbose='-v' in sys.argv, trace='--trace' in sys.argv))
