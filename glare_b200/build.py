"""Builds libglare_b200.so (the C-ABI library declared in include/glare_b200.h) in-tree with nvcc for sm_100a.

    python -m glare_b200.build [--force] [--verbose]

The .so is git-ignored but travels to the GPU box with the repo snapshot.  nvcc cross-compiles without a GPU.
"""
import concurrent.futures
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libglare_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "--use_fast_math=false", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]


def _flags():
    # no fast-math: the VQ distance, the flow divisions and the bilinear weights must round like the reference
    return [f for f in FLAGS if not f.startswith("--use_fast_math")]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def source_hash(files=("conv_tc.cu", "tc.cuh", "common.cuh")):
    """sha256 (first 16 hex digits) over the sources of the bench's dominant kernel (conv_tc_kernel and the headers it includes): ties the
    committed ncu capture behind `roofline.traffic` to the code it measured"""
    import hashlib
    h = hashlib.sha256()
    for f in sorted(os.path.join(CSRC, n) for n in files):
        h.update(os.path.basename(f).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def build(force=False, verbose=False):
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    hdrs = sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + [os.path.join(os.path.dirname(HERE), "include", "glare_b200.h")]
    os.makedirs(OBJ, exist_ok=True)
    jobs = []
    for s in srcs:
        o = os.path.join(OBJ, os.path.basename(s)[:-3] + ".o")
        if force or _stale(o, [s] + hdrs):
            jobs.append((s, o))

    def cc(job):
        s, o = job
        cmd = [NVCC] + ARCH + _flags() + ["-Xptxas", "-v" if verbose else "-warn-spills", "-I", CSRC, "-c", s, "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (s, r.stdout, r.stderr))
        return s, r.stderr

    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for s, err in ex.map(cc, jobs):
            if verbose:
                print(err)
    objs = [os.path.join(OBJ, os.path.basename(s)[:-3] + ".o") for s in srcs]
    if force or jobs or _stale(LIB, objs):
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-Xcompiler", "-fPIC"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
