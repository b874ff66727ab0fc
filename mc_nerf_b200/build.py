"""Build libmcnerf.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m mc_nerf_b200.build [--force]

Each .cu under csrc/ is compiled to an object in csrc/build/ (parallel, only when stale) and linked
into mc_nerf_b200/csrc/libmcnerf.so.  The .so is git-ignored but travels to the GPU box.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
OUT = os.path.join(CSRC, "libmcnerf.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
# measurement / debugging variants (environment): MCNERF_SPIN (test_wait spinning), MCNERF_TRACE (clock64 traces of the
# tensor-core kernels, tools/trace_fwd.py), MCNERF_CHAOS (random delays at every mbarrier hand-off: shakes out protocol
# races - the parity and bit-exact determinism tests must still pass)
VARIANT = [f for f, e in (("-DMCNERF_SPIN_TEST_WAIT", "MCNERF_SPIN"), ("-DMCNERF_TC_TRACE", "MCNERF_TRACE"),
                          ("-DMCNERF_CHAOS", "MCNERF_CHAOS")) if os.environ.get(e)]
FLAGS = VARIANT + ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC",
         "-I", os.path.join(ROOT, "include"), "-I", CSRC]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    srcs = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(ROOT, "include", "mcnerf.h"))
    bdir = os.path.join(CSRC, "build")
    os.makedirs(bdir, exist_ok=True)
    stamp = os.path.join(bdir, "variant.txt")         # objects built with other variant flags are stale
    if not os.path.exists(stamp) or open(stamp).read() != " ".join(VARIANT):
        force = True
    jobs = []
    for s in srcs:
        src = os.path.join(CSRC, s)
        obj = os.path.join(bdir, s[:-3] + ".o")
        if force or _stale(obj, [src] + hdrs):
            cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            jobs.append((s, cmd))

    def run(job):
        name, cmd = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        return name, r

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for name, r in ex.map(run, jobs):
            if verbose or r.returncode != 0:
                sys.stderr.write(f"--- nvcc {name}\n{r.stdout}{r.stderr}\n")
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed on {name}")
    with open(stamp, "w") as f:
        f.write(" ".join(VARIANT))
    objs = [os.path.join(bdir, s[:-3] + ".o") for s in srcs]
    if force or jobs or _stale(OUT, objs):
        cmd = [NVCC, "-shared", "-o", OUT] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
