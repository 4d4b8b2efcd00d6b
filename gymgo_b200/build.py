"""Build libgymgo_b200.so (hand-written sm_100a kernels + C ABI) IN-TREE with nvcc.

    python -m gymgo_b200.build [--force]

One nvcc job per board size (gg_size.cu -DGG_N=n) plus gg_api.cu, run in parallel, then one link.
Outputs: gymgo_b200/_lib/libgymgo_b200.so (+ objects under gymgo_b200/_lib/obj/).  *.so/*.o are
git-ignored but travel to the GPU box with the gpurun snapshot.
"""
import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "_lib")
OBJDIR = os.path.join(LIBDIR, "obj")
LIB = os.path.join(LIBDIR, "libgymgo_b200.so")
SIZES = list(range(2, 20))
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
HOST_FLAGS = ["-std=c++17", "-O3", "-fPIC", "-fvisibility=hidden", "-pthread"]
NVCC_FLAGS = ["-std=c++17", "-O3", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"] + ARCH


def nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)) + [os.path.join(HERE, "..", "include", "gymgo_b200.h")]


def source_digest():
    h = hashlib.sha256()
    for p in _sources():
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS + HOST_FLAGS).encode())
    return h.hexdigest()


def up_to_date():
    stamp = os.path.join(LIBDIR, "digest.txt")
    return os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == source_digest()


def _run(cmd):
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if p.returncode != 0:
        raise RuntimeError("command failed: %s\n%s" % (" ".join(cmd), p.stdout))
    return p.stdout


def build(force=False, verbose=False):
    if not force and up_to_date():
        return LIB
    os.makedirs(OBJDIR, exist_ok=True)
    for stale in (LIB, os.path.join(LIBDIR, "digest.txt")):      # never leave a stale library behind a failed build
        if os.path.exists(stale):
            os.remove(stale)
    cc = nvcc()
    jobs = []
    for n in SIZES:
        obj = os.path.join(OBJDIR, "gg_n%d.o" % n)
        jobs.append((obj, [cc] + NVCC_FLAGS + ["-DGG_N=%d" % n, "-c", os.path.join(CSRC, "gg_size.cu"), "-o", obj]))
    api = os.path.join(OBJDIR, "gg_api.o")
    jobs.append((api, [cc] + NVCC_FLAGS + ["-c", os.path.join(CSRC, "gg_api.cu"), "-o", api]))
    host = os.path.join(OBJDIR, "gg_host.o")                    # host codec: plain C++ (AVX-512 paths picked at run time)
    jobs.append((host, [os.environ.get("CXX") or shutil.which("g++") or "g++"] + HOST_FLAGS
                 + ["-c", os.path.join(CSRC, "gg_host.cpp"), "-o", host]))
    with concurrent.futures.ThreadPoolExecutor(max_workers=max(2, os.cpu_count() or 2)) as ex:
        for out in ex.map(lambda j: _run(j[1]), jobs):
            if verbose and out.strip():
                print(out)
    # cudart is linked statically (nvcc default): the .so only needs the driver on the GPU box
    _run([cc, "-shared", "-o", LIB] + ARCH + [j[0] for j in jobs])
    with open(os.path.join(LIBDIR, "digest.txt"), "w") as f:
        f.write(source_digest())
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
