"""Builds resnet_b200/libresnet_b200.so (in-tree) with nvcc for sm_100a.  No torch involved: the product is a
plain C-ABI shared library (include/resnet.h, include/resnet_b200.h)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libresnet_b200.so")
SOURCES = ["bw_kernels.cu", "simt_conv.cu", "igemm.cu", "model.cu", "capi.cu", "dp.cu", "prof.cu", "loader.cu", "dump.cu", "selfcheck.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--threads", "0",
         "-Xcompiler", "-fPIC,-fvisibility=default", "-shared"]


def needs_build():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", f) for f in ("resnet.h", "resnet_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return SO
    cmd = [NVCC] + FLAGS + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", SO, "-lcurand", "-ldl"]
    if verbose:
        print(" ".join(cmd))
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("nvcc failed building libresnet_b200.so")
    if verbose and r.stdout.strip():
        print(r.stdout)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
