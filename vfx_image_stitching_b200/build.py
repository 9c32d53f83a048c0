"""Builds libb200sift.so in-tree with nvcc for sm_100a (no GPU needed).

    python -m vfx_image_stitching_b200.build [--force]

The .so is git-ignored but travels to the GPU box with the repo snapshot.
detect.cu is compiled with --fmad=false: numpy rounds every float32
operation once and the sparse stage mirrors that (see detect.cu header).
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OUT = os.path.join(HERE, 'libb200sift.so')
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
COMMON = ['-O3', '-lineinfo', '-std=c++17', '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden',
          '-DB200SIFT_BUILD', '--expt-relaxed-constexpr']
SOURCES = {
    'pyramid.cu': [],
    'detect.cu': ['--fmad=false'],
    'match.cu': [],
    'match_tc.cu': [],
    'stitch.cu': ['--fmad=false'],
    'api.cu': [],
}


def nvcc():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found: libb200sift cannot be built (there is no CPU fallback)')


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    cc = nvcc()
    extra_all = os.environ.get('B200SIFT_NVCC_FLAGS', '').split()   # build-time experiments (-D...)
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))]
    hdrs.append(os.path.join(HERE, '..', 'include', 'b200sift.h'))
    objs = []
    procs = []
    for src, extra in SOURCES.items():
        s = os.path.join(CSRC, src)
        if not os.path.exists(s):
            continue
        o = os.path.join(CSRC, src[:-3] + '.o')
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            cmd = [cc] + ARCH + COMMON + extra + extra_all + ['-c', s, '-o', o]
            if verbose:
                cmd.insert(1, '-Xptxas=-v')
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if out.strip() and (verbose or p.returncode != 0):
            print(f'--- {src}\n{out}', file=sys.stderr)
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError('nvcc failed')
    if force or procs or _stale(OUT, objs):
        cmd = [cc] + ARCH + ['-shared', '-o', OUT] + objs + ['-cudart', 'static', '-Xlinker', '--no-undefined']
        subprocess.check_call(cmd)
    return OUT


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
