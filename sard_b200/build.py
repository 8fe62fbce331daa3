"""Build libsard_b200.so (the C-ABI library) in-tree with nvcc for sm_100a."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libsard_b200.so')
SOURCES = ['engine.cu', 'prof.cu', 'pack.cu', 'norm.cu', 'graph.cu', 'gemm.cu', 'gemm_fast.cu', 'gemm_skinny.cu', 'gemm_mma.cu', 'umma.cu', 'jsmlp_tc.cu', 'jsmlp_tc_bwd.cu', 'tower_tc.cu', 'critic_tc.cu', 'encoder.cu', 'loss.cu', 'gae.cu', 'optim.cu', 'p2p.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '--use_fast_math=false', '-Xcompiler', '-fPIC', '-Xptxas', '-v']


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, '..', 'include', 'sard_b200.h')]
    return any(os.path.getmtime(d) > t for d in deps)


HOSTPACK = os.path.join(HERE, '_hostpack.so')


def build_hostpack() -> str:
    """Native host packer (CPython extension, no CUDA): g++ only."""
    import sysconfig
    import numpy
    inc = sysconfig.get_paths()['include']
    subprocess.check_call(['g++', '-O3', '-std=c++17', '-shared', '-fPIC', '-pthread', f'-I{inc}', f'-I{numpy.get_include()}',
                           os.path.join(CSRC, 'hostpack.cpp'), '-o', HOSTPACK])
    return HOSTPACK


def build_variant(tag: str, defines) -> str:
    """A tuning build with extra -D flags into libsard_b200_<tag>.so (selected at run time with SARD_LIB)."""
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    out_dir = os.path.join(HERE, 'build', tag)
    os.makedirs(out_dir, exist_ok=True)
    procs, objs = [], []
    for src in SOURCES:
        obj = os.path.join(out_dir, src.replace('.cu', '.o'))
        objs.append(obj)
        cmd = [nvcc] + [f for f in NVCC_FLAGS if f not in ('--use_fast_math=false', '-Xptxas', '-v')] + [f'-D{d}' for d in defines] + ['-c', os.path.join(CSRC, src), '-o', obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f'nvcc failed on {src}:\n{out}')
    lib = os.path.join(HERE, f'libsard_b200_{tag}.so')
    subprocess.check_call([nvcc, '-shared', '-o', lib] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a'])
    return lib


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, 'build'), exist_ok=True)
    for src in SOURCES:
        obj = os.path.join(HERE, 'build', src.replace('.cu', '.o'))
        objs.append(obj)
        cmd = [nvcc] + [f for f in NVCC_FLAGS if f != '--use_fast_math=false'] + ['-c', os.path.join(CSRC, src), '-o', obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, p in procs:
        out, _ = p.communicate()
        log.append(f'== {src}\n{out}')
        if p.returncode != 0:
            raise RuntimeError(f'nvcc failed on {src}:\n{out}')
    subprocess.check_call([nvcc, '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a'])
    build_hostpack()
    with open(os.path.join(HERE, 'build', 'ptxas.log'), 'w') as f:
        f.write('\n'.join(log))
    if verbose:
        print('\n'.join(log))
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
