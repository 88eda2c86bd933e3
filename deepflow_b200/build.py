"""In-tree builds: libdfgpu.so (nvcc, sm_100a), libdfsynth.so (g++) and, for the
tests / cpu baseline only, oracle/liboracle.so."""
import glob
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
CXX = os.environ.get("CXX") or shutil.which("g++") or "g++"
NVCC = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd, cwd=None):
    r = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("build failed: %s\n%s" % (" ".join(cmd), r.stdout))
    return r.stdout


def build_synth(force=False):
    src = os.path.join(_HERE, "synth", "synth.cc")
    out = os.path.join(_HERE, "synth", "libdfsynth.so")
    if force or _newer(out, [src, os.path.join(ROOT, "include", "dfgpu.h")]):
        _run([CXX, "-O2", "-std=c++17", "-fPIC", "-shared", "-o", out, src])
    return out


def build_oracle(force=False):
    d = os.path.join(ROOT, "oracle")
    out = os.path.join(d, "liboracle.so")
    srcs = glob.glob(os.path.join(d, "*.cc")) + glob.glob(os.path.join(d, "*.h")) + \
        glob.glob(os.path.join(ROOT, "include", "*.h"))
    if force or _newer(out, srcs):
        _run(["make", "-C", d, "CXX=" + CXX] + (["-B"] if force else []))
    return out


NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def build_dfgpu(force=False, verbose=False):
    d = os.path.join(_HERE, "csrc")
    out = os.path.join(_HERE, "libdfgpu.so")
    cus = sorted(glob.glob(os.path.join(d, "*.cu")))
    ccs = sorted(glob.glob(os.path.join(d, "*.cc")))
    hdrs = glob.glob(os.path.join(d, "*.h")) + glob.glob(os.path.join(d, "*.cuh")) + \
        glob.glob(os.path.join(ROOT, "include", "*.h"))
    if not (force or _newer(out, cus + ccs + hdrs)):
        return out
    objs = []
    os.makedirs(os.path.join(d, "build"), exist_ok=True)
    log = []
    for s in cus + ccs:
        o = os.path.join(d, "build", os.path.basename(s) + ".o")
        if force or _newer(o, [s] + hdrs):
            cmd = [NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
                ["-x", "cu", "-I", os.path.join(ROOT, "include"), "-c", s, "-o", o]
            log.append(_run(cmd))
        objs.append(o)
    _run([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out] + objs +
         ["-lcudart"])
    return "\n".join(log) if verbose else out


def build_host(force=False):
    """libdfhost.so: the C++ mirror of the decoder goroutine (queue -> dfgpu_submit -> blocks)."""
    d = os.path.join(_HERE, "host")
    out = os.path.join(d, "libdfhost.so")
    srcs = glob.glob(os.path.join(d, "*.cc"))
    deps = srcs + glob.glob(os.path.join(d, "*.h")) + glob.glob(os.path.join(ROOT, "include", "*.h")) + \
        [os.path.join(_HERE, "libdfgpu.so")]
    if force or _newer(out, deps):
        _run([CXX, "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-o", out] + srcs +
             ["-I", os.path.join(ROOT, "include"), "-L", _HERE, "-ldfgpu", "-Wl,-rpath,$ORIGIN/.."])
    return out


def build_all(force=False):
    build_synth(force)
    build_dfgpu(force)
    build_host(force)
    return True
