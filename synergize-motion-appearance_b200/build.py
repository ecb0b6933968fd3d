"""In-tree nvcc build of libsma_b200.so for sm_100a (no torch headers: the library is plain C ABI).
Every .cu is compiled to an object file in parallel (csrc/build/), then linked."""
import glob
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(CSRC, 'build')
LIB = os.path.join(CSRC, 'libsma_b200.so')
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '-Xcompiler', '-fPIC']


def sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def _headers():
    return glob.glob(os.path.join(CSRC, '*.cuh')) + [os.path.join(os.path.dirname(HERE), 'include', 'sma_b200.h')]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def is_stale() -> bool:
    return _stale(LIB, sources() + _headers())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    os.makedirs(OBJ, exist_ok=True)
    hdrs = _headers()

    def compile_one(src):
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + '.o')
        if not force and not _stale(obj, [src] + hdrs):
            return obj, ''
        cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', src, '-o', obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('nvcc failed on ' + src + ':\n' + r.stdout + r.stderr)
        return obj, r.stderr
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        results = list(ex.map(compile_one, sources()))
    r = subprocess.run([nvcc, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', LIB] + [o for o, _ in results],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('link failed:\n' + r.stdout + r.stderr)
    if verbose:
        print(''.join(log for _, log in results))
    return LIB


if __name__ == '__main__':
    print(build(force=True, verbose=True))
