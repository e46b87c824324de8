"""In-tree build of the native libraries (nvcc, sm_100a only).

    libsad_b200.so                     C-ABI kernels (include/sad_b200.h)
    libsad_exchange.so                 the gradient exchange: host C++ over NCCL (include/sad_exchange.h)
    libcaffe2_detectron_ops_gpu.so     operator-boundary library: shim runtime + operator classes,
                                       named after the reference module it replaces
                                       (caffe2/modules/detectron/CMakeLists.txt:7-13)

Both land next to this file so they travel with the gpurun snapshot.  Rebuilds only when a
source is newer than the library.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-std=c++17", "-O3", "-lineinfo", "--use_fast_math=false"][:3] + ARCH + [
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-fno-gnu-unique", "-w"]

LIB_KERNELS = os.path.join(HERE, "libsad_b200.so")
LIB_OPS = os.path.join(HERE, "libcaffe2_detectron_ops_gpu.so")
LIB_EXCHANGE = os.path.join(HERE, "libsad_exchange.so")   # host C++ over NCCL (include/sad_exchange.h); NCCL is resolved at run time


def _sources(*rel):
    return [os.path.join(CSRC, r) for r in rel]


def _walk_headers():
    out = []
    for base in (CSRC, os.path.join(ROOT, "include")):
        for d, _, files in os.walk(base):
            out += [os.path.join(d, f) for f in files if f.endswith((".h", ".cuh"))]
    return out


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("build failed: " + " ".join(cmd[:3]) + " ...")
    return r.stdout


def kernel_sources():
    kdir = os.path.join(CSRC, "kernels")
    return sorted(os.path.join(kdir, f) for f in os.listdir(kdir) if f.endswith(".cu"))


def build(force=False, verbose=False):
    nvcc = os.environ.get("NVCC", "nvcc")
    hdrs = _walk_headers()
    ksrc = kernel_sources()
    if force or _stale(LIB_KERNELS, ksrc + hdrs):
        _run([nvcc] + COMMON + ["-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(CSRC, "kernels"),
                                "-shared", "-o", LIB_KERNELS] + ksrc, verbose)
    osrc = _sources("caffe2_shim/shim_runtime.cc", "caffe2_shim/shim_c_api.cc") + sorted(
        os.path.join(CSRC, "ops", f) for f in os.listdir(os.path.join(CSRC, "ops")) if f.endswith((".cc", ".cu")))
    if force or _stale(LIB_OPS, osrc + hdrs + [LIB_KERNELS]):
        _run([nvcc] + COMMON + ["-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(CSRC, "caffe2_shim"), "-I" + CSRC,
                                "-x", "cu", "-shared", "-o", LIB_OPS] + osrc +
             ["-L" + HERE, "-l:libsad_b200.so", "-Xlinker", "-rpath=$ORIGIN"], verbose)
    xsrc = _sources("exchange/sad_exchange.cc", "exchange/slot_sum.cu")
    if force or _stale(LIB_EXCHANGE, xsrc + hdrs):
        _run([nvcc, "-std=c++17", "-O2", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC,-fvisibility=hidden,-fno-gnu-unique",
              "-w", "-I" + os.path.join(ROOT, "include"), "-shared", "-o", LIB_EXCHANGE] + xsrc + ["-ldl"], verbose)
    return LIB_KERNELS, LIB_OPS, LIB_EXCHANGE


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
