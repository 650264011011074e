"""Builds libmolclr_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m molclr_b200.build [--force] [--verbose] [--debug-switches]

``--debug-switches`` builds ``libmolclr_b200_dbg.so`` with ``-DMOLCLR_DEBUG_SWITCHES``: the only build in which the
``MOLCLR_GEMM_* / MOLCLR_AGG_* / MOLCLR_NTX_*`` tuning variables are read (tools/ point ``MOLCLR_B200_LIB`` at it).

The shared object has no torch / Python dependency: plain ``extern "C"`` entry points declared in
``include/molclr_b200.h``.  nvcc cross-compiles without a GPU.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmolclr_b200.so")
SOURCES = ["api.cu", "plan.cu", "rowwise.cu", "tables.cu", "gin_step.cu", "gemm.cu", "ntxent.cu", "ntxent_fused.cu", "augment.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
    "-I", os.path.join(ROOT, "include"), "-I", CSRC,
]


def _stale(lib=LIB):
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "molclr_b200.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, debug_switches=False):
    lib = LIB.replace(".so", "_dbg.so") if debug_switches else LIB
    if not force and not _stale(lib):
        return lib
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objdir = os.path.join(HERE, "build_dbg" if debug_switches else "build")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *(["-DMOLCLR_DEBUG_SWITCHES"] if debug_switches else []), "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs, failed = [], False
    for src, obj, pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0:
            failed = True
            sys.stderr.write(f"--- nvcc failed on {src}\n{out}\n")
        elif verbose or out.strip():
            sys.stderr.write(f"--- {src}\n{out}\n")
        objs.append(obj)
    if failed:
        raise RuntimeError("molclr_b200: CUDA build failed")
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib, *objs, "-lcudart_static", "-ldl", "-lrt", "-lpthread"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("molclr_b200: link failed\n" + r.stdout)
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, debug_switches="--debug-switches" in sys.argv))
