"""Builds videomorphing_b200/libvmorph.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libvmorph.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-fmad=false",            # parity contract: no FMA contraction, IEEE div/sqrt (nvcc defaults)
         "-shared", "-Xcompiler", "-fPIC,-ffp-contract=off", "-ccbin", "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"]


def sources():
    return sorted(glob.glob(os.path.join(SRC, "*.cu")))


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = sources() + glob.glob(os.path.join(SRC, "*.h")) + glob.glob(os.path.join(SRC, "*.cuh")) + \
        [os.path.join(HERE, "..", "include", "vmorph.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + sources()
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libvmorph.so")
    return OUT


def build_headless(force=False):
    """tools/vmorph_headless: the headless C++ driver, plain g++ against include/vmorph.h + libvmorph.so (rpath'd in-tree)."""
    src = os.path.join(HERE, "..", "tools", "vmorph_headless.cpp")
    out = os.path.join(HERE, "vmorph_headless")
    if not force and os.path.exists(out) and os.path.getmtime(out) > max(os.path.getmtime(src), os.path.getmtime(OUT)):
        return out
    cmd = ["g++", "-O2", "-std=c++17", "-o", out, src, "-L" + HERE, "-l:libvmorph.so", "-Wl,-rpath,$ORIGIN"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("g++ failed building vmorph_headless")
    return out


def build_video_host(force=False):
    """tools/vmorph_video: the multi-GPU video host in C++ (threads + peer copies over the C ABI), plain g++."""
    src = os.path.join(HERE, "..", "tools", "vmorph_video.cpp")
    out = os.path.join(HERE, "vmorph_video")
    if not force and os.path.exists(out) and os.path.getmtime(out) > max(os.path.getmtime(src), os.path.getmtime(OUT)):
        return out
    cmd = ["g++", "-O2", "-std=c++17", "-pthread", "-o", out, src, "-L" + HERE, "-l:libvmorph.so", "-Wl,-rpath,$ORIGIN"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("g++ failed building vmorph_video")
    return out


def build_trace():
    """Development aid: libvmorph_trace.so = the same sources with -DVM_TRACE (per-phase cycle counters in the sweep)."""
    out = os.path.join(HERE, "libvmorph_trace.so")
    r = subprocess.run([NVCC] + FLAGS + ["-DVM_TRACE", "-o", out] + sources(), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("nvcc failed building libvmorph_trace.so")
    return out


if __name__ == "__main__":
    if "--trace" in sys.argv:
        print(build_trace())
    else:
        build(force=True, verbose="-v" in sys.argv)
        print(OUT)
        print(build_headless(force=True))
