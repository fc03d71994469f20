"""Compiles ``csrc/*.cu`` into ``csrc/libltr_sm100.so`` with nvcc for sm_100a.

The library is built IN-TREE (it ships to the GPU box with the repository snapshot)
and has no dependency on torch or Python: only the CUDA runtime (linked statically).

    python -m pytorchltr_b200.build [--force] [--verbose]
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libltr_sm100.so")
PARSER_LIB = os.path.join(CSRC, "libltr_svmrank.so")
PARSER_SRC = os.path.join(CSRC, "svmrank_parser.cpp")
SOURCES = ["ltr_kernels.cu", "ltr_mlp.cu"]
# headers a source does NOT include (so editing them does not recompile it)
NOT_INCLUDED = {"ltr_kernels.cu": {"ltr_mlp_scorer.cuh"},
                "ltr_mlp.cu": {f for f in os.listdir(CSRC) if f.endswith(".cuh")} -
                              {"ltr_mlp_scorer.cuh", "ltr_common.cuh", "ltr_host.cuh", "ltr_p2p.cuh"}}
OBJ_DIR = os.path.join(ROOT, "build", "obj")
NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-cudart", "static",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: set NVCC or put /usr/local/cuda/bin on PATH")


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    deps.append(os.path.join(ROOT, "include", "ltr_sm100.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build_parser(force: bool = False) -> str:
    """The SVMrank text parser (host C++, no CUDA): csrc/libltr_svmrank.so."""
    if not force and os.path.exists(PARSER_LIB) and os.path.getmtime(PARSER_LIB) >= os.path.getmtime(PARSER_SRC):
        return PARSER_LIB
    cxx = shutil.which("g++") or "/usr/bin/g++"
    cmd = [cxx, "-O3", "-std=c++17", "-fPIC", "-shared", "-pthread", "-Wall", "-o", PARSER_LIB, PARSER_SRC]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout)
        raise RuntimeError("g++ failed:\n" + " ".join(cmd))
    return PARSER_LIB


def build(force: bool = False, verbose: bool = False, defines=(), out: str = None) -> str:
    """Builds the shared library if it is missing or older than its sources.  `defines` / `out`
    build an A-B variant (e.g. defines=["LTR_RING_MIN_CTAS=4"], out="build/ltr_variant.so"; select it
    at run time with LTR_SM100_LIB)."""
    target = out or LIB
    if not out:
        build_parser(force)
    if not force and not out and not _stale():
        return LIB
    base = [_nvcc()] + NVCC_FLAGS + ["-I", os.path.join(ROOT, "include"), "-I", CSRC]
    base += ["-D" + d for d in defines]
    if verbose:
        base += ["-Xptxas", "-v"]
    # g++ from PATH: the image's CC/CXX wrappers are not needed for nvcc's host pass
    env = dict(os.environ)

    def run(cmd):
        res = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if verbose or res.returncode != 0:
            sys.stderr.write(res.stdout)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd))

    # one object per source, recompiled only when the source or a header it includes is newer
    tag = "" if not defines else "_" + "_".join(sorted(defines)).replace("=", "-")
    os.makedirs(OBJ_DIR, exist_ok=True)
    headers = [f for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    objs, procs = [], []
    for src in SOURCES:
        obj = os.path.join(OBJ_DIR, src.replace(".cu", tag + ".o"))
        deps = [os.path.join(CSRC, src), os.path.join(ROOT, "include", "ltr_sm100.h")]
        deps += [os.path.join(CSRC, h) for h in headers if h not in NOT_INCLUDED.get(src, ())]
        objs.append(obj)
        if force or verbose or not os.path.exists(obj) or any(os.path.getmtime(d) > os.path.getmtime(obj) for d in deps):
            cmd = base + ["-c", "-o", obj, os.path.join(CSRC, src)]
            procs.append((cmd, subprocess.Popen(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                                                text=True)))
    failed = None
    for cmd, pr in procs:
        text, _ = pr.communicate()
        if verbose or pr.returncode != 0:
            sys.stderr.write(text)
        if pr.returncode != 0:
            failed = cmd
    if failed:
        raise RuntimeError("nvcc failed:\n" + " ".join(failed))
    run(base + ["-shared", "-o", target] + objs)
    return target


if __name__ == "__main__":
    _defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    _out = next((a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--out=")), None)
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, defines=_defs, out=_out))
