"""Build libb200render.so in-tree with nvcc for sm_100a (no JIT cache: the .so must travel with the repo)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libb200render.so")
SOURCES = ["api.cu", "api_train.cu", "field_fwd.cu", "field_fwd_train.cu", "field_bwd.cu", "wgrad.cu", "chain.cu", "prologue.cu", "pack.cu", "composite.cu", "compose.cu", "importance.cu", "match.cu", "losses.cu", "quat.cu"]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--expt-relaxed-constexpr",
         "-Xcompiler", "-fPIC", "-Xptxas", "-v", "-DB200R_CLUSTER=" + os.environ.get("B200R_CLUSTER", "2"), "-DB200R_WATCHDOG=" + os.environ.get("B200R_WATCHDOG", "1")]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "b200r.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    from concurrent.futures import ThreadPoolExecutor

    lib_t = os.path.getmtime(LIB) if os.path.exists(LIB) else 0.0
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))] + [os.path.join(HERE, "..", "include", "b200r.h")]
    hdr_t = max(os.path.getmtime(h) for h in headers)

    def compile_one(src):
        obj = os.path.join(CSRC, src.replace(".cu", ".o"))
        path = os.path.join(CSRC, src)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(path), hdr_t):
            return obj, 0, ""  # object is newer than its source and every header
        r = subprocess.run(["nvcc", *FLAGS, "-c", path, "-o", obj], capture_output=True, text=True)
        return obj, r.returncode, r.stdout + r.stderr

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:  # translation units compile in parallel
        results = list(ex.map(compile_one, SOURCES))
    objs = []
    for src, (obj, rc, log) in zip(SOURCES, results):
        if verbose or rc != 0:
            sys.stderr.write(log)
        if rc != 0:
            raise RuntimeError("nvcc failed on " + src)
        objs.append(obj)
    cmd = ["nvcc", "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
