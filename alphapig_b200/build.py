"""Build libalphapig_b200.so in-tree with nvcc for sm_100a (no torch dependency in the library)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libalphapig_b200.so")
SOURCES = ["engine.cu", "boards.cu", "tree.cu", "rollout.cu", "net.cu", "conv_tc.cu", "front_tc.cu", "heads_tc.cu", "replay.cu", "traj.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--extended-lambda", "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "-Xcompiler", "-Wno-unused-function",
              "-fmad=true"]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [
        os.path.join(HERE, "..", "include", "alphapig_b200.h"), os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every source of csrc/ for sm_100a and link libalphapig_b200.so in-tree.  Prints ONE line saying whether it
    compiled or re-used an up-to-date library (so a build log shows which it was)."""
    if not force and not _stale():
        print("[alphapig_b200.build] up to date, re-using %s (newer than every file of csrc/ and the header); "
              "--force recompiles" % os.path.relpath(LIB))
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("AP_NVCC_EXTRA", "").split()  # development A/B builds, e.g. -DAP_PURE_MINBLK=24
    objs, procs = [], []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for src in SOURCES:
        obj = os.path.join(HERE, "build", src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write("== %s ==\n%s\n" % (src, out))
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + objs +
                          ["-lcudart_static", "-lpthread", "-ldl", "-lrt"])
    print("[alphapig_b200.build] compiled %d sources with %s -gencode arch=compute_100a,code=sm_100a -> %s"
          % (len(SOURCES), nvcc, os.path.relpath(LIB)))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
