"""Build libairfe.so (hand-written sm_100a CUDA + C ABI) in-tree with nvcc.  No torch extension machinery:
the product is a plain C-ABI shared library (include/airfe_c.h)."""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libairfe.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--use_fast_math=false",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "-Xcompiler", "-Wno-unused-function"]
FLAGS.remove("--use_fast_math=false")


STAMP = LIB + ".srchash"


def source_hash():
    """sha256 over every file the library is built from: a stamp next to the .so lets the loader detect a stale build
    independently of file times (which a snapshot copy does not preserve)."""
    h = hashlib.sha256()
    files = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cc", ".cuh", ".h")))
    files.append(os.path.join(os.path.dirname(HERE), "include", "airfe_c.h"))
    for f in files:
        h.update(os.path.basename(f).encode())
        h.update(open(f, "rb").read())
    return h.hexdigest()


def is_current():
    return os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == source_hash()


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cc")))


def _stale(objs_srcs):
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [s for _, s in objs_srcs] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "airfe_c.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    pairs = [(os.path.join(HERE, "build", os.path.basename(s) + ".o"), s) for s in sources()]
    if not force and not _stale(pairs) and is_current():
        return LIB
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(os.path.dirname(HERE), "include", "airfe_c.h"))
    hdr_t = max(os.path.getmtime(h) for h in hdrs)
    procs = []
    for obj, src in pairs:
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), hdr_t):
            continue
        cmd = [NVCC] + FLAGS + ["-x", "cu", "-c", src, "-o", obj]
        if verbose:
            print(" ".join(cmd))
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out.decode())
            raise RuntimeError("nvcc failed on " + src)
        if verbose and out:
            print(out.decode())
    # link to a temporary name and rename over the library: a process that already dlopen'ed the old file keeps its (now unlinked)
    # image instead of having its mapped pages rewritten under it
    tmp = LIB + ".tmp.%d" % os.getpid()
    cmd = [NVCC, "-shared", "-o", tmp] + [o for o, _ in pairs] + ["-lcudart"]
    subprocess.check_call(cmd)
    os.replace(tmp, LIB)
    with open(STAMP + ".tmp", "w") as f:
        f.write(source_hash() + "\n")
    os.replace(STAMP + ".tmp", STAMP)
    return LIB


def have_nvcc():
    return os.path.exists(NVCC)


def build_mock_caller():
    """g++: the C++ class surfaces (include/*.h + src/frontend/frontend.cc) against the compat shims, linked with libairfe.so."""
    root = os.path.dirname(HERE)
    out = os.path.join(root, "tests", "cpp", "mock_caller")
    srcs = [os.path.join(root, "src", "frontend", "frontend.cc"), os.path.join(root, "tests", "cpp", "mock_caller.cc")]
    deps = srcs + [os.path.join(root, "include", f) for f in os.listdir(os.path.join(root, "include"))] + [LIB]
    if os.path.exists(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in deps):
        return out
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-I" + os.path.join(root, "compat"), "-I" + os.path.join(root, "include")] + srcs + [
        "-o", out, "-L" + HERE, "-lairfe", "-Wl,-rpath," + HERE, "-Wl,-rpath,$ORIGIN/../../airslam_b200"]
    subprocess.check_call(cmd)
    return out


def build_ransac_test():
    """g++: CPU unit test of the RANSAC hook (tests/cpp/ransac_test.cc)."""
    root = os.path.dirname(HERE)
    out = os.path.join(root, "tests", "cpp", "ransac_test")
    srcs = [os.path.join(root, "src", "frontend", "frontend.cc"), os.path.join(root, "tests", "cpp", "ransac_test.cc")]
    deps = srcs + [os.path.join(root, "compat", "opencv2", "opencv.hpp"), LIB]
    if os.path.exists(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in deps):
        return out
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-I" + os.path.join(root, "compat"), "-I" + os.path.join(root, "include")] + srcs + [
        "-o", out, "-L" + HERE, "-lairfe", "-Wl,-rpath," + HERE, "-Wl,-rpath,$ORIGIN/../../airslam_b200"]
    subprocess.check_call(cmd)
    return out


def build_class_bench():
    """g++: tests/cpp/class_bench.cc (bench.py --mode class) against the same class surfaces."""
    root = os.path.dirname(HERE)
    out = os.path.join(root, "tests", "cpp", "class_bench")
    srcs = [os.path.join(root, "src", "frontend", "frontend.cc"), os.path.join(root, "tests", "cpp", "class_bench.cc")]
    deps = srcs + [os.path.join(root, "include", f) for f in os.listdir(os.path.join(root, "include"))] + [LIB]
    if os.path.exists(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in deps):
        return out
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-I" + os.path.join(root, "compat"), "-I" + os.path.join(root, "include")] + srcs + [
        "-o", out, "-L" + HERE, "-lairfe", "-Wl,-rpath," + HERE, "-Wl,-rpath,$ORIGIN/../../airslam_b200"]
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
