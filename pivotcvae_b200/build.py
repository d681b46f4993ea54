"""In-tree build of libpcv_b200.so (sm_100a only) with plain nvcc.

The .so lands next to this file so it travels with the repo snapshot to the GPU
box; it is git-ignored.  `python -m pivotcvae_b200.build` rebuilds it.
"""
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpcv_b200.so")
STAMP = os.path.join(HERE, "build", "sources.sha256")

SOURCES = ["pcv_api.cu", "score_select.cu", "score_select_tc.cu", "mlp.cu", "mlp_tc.cu", "ce.cu", "ce_tc2.cu", "urm.cu", "sampler.cu", "topk.cu", "gemm_tc.cu", "pretrain.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",  # every FMA in the library is an explicit fmaf(): see DESIGN.md "arithmetic contract"
    "-Xcompiler", "-fPIC",
]


def _nvcc():
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found: libpcv_b200.so cannot be built (there is no CPU fallback)")
    return cand


def _digest():
    h = hashlib.sha256()
    files = sorted(os.listdir(CSRC)) + ["../../include/pcv_b200.h"]
    for f in files:
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(f.encode())
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def kernel_digest(source):
    """Digest of one kernel source + the shared headers + the flags: ties an ncu capture to the build it was taken on."""
    h = hashlib.sha256()
    for f in (source, "tc_common.cuh", "pcv_common.cuh"):
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()[:16]


def build(force=False, verbose=False):
    """Compile every .cu for sm_100a and link the shared library. Returns its path."""
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP):
        with open(STAMP) as fh:
            if fh.read().strip() == dig:
                return LIB
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = []
    for src, obj, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s" % (src, out))
        if verbose:
            sys.stdout.write(out)
        objs.append(obj)
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    subprocess.run(cmd, check=True)
    with open(STAMP, "w") as fh:
        fh.write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
