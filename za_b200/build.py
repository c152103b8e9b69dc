"""Builds za_b200/libza_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels to the GPU box)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libza_b200.so")
OUT_ZA2C = os.path.join(HERE, "libza2c.so")
FRONTEND_SOURCES = ["parser.cpp", "bincode.cpp", "eval.cpp"]      # za's front-end (host C++), part of libza2c.so
SOURCES = ["capi.cu", "ntt.cu", "msm.cu", "prove.cu", "format.cu", "verify.cu", "setup.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
              "-Xcompiler", "-O2", "--expt-relaxed-constexpr", "-rdc=false"]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "za_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_variant(tag, defines, verbose=False):
    """Development only: a second copy of the library compiled with extra -D flags into za_b200/variants/ (selected at
    run time with ZA_B200_SO=<path>) so that one GPU call can time several code variants side by side."""
    out_dir = os.path.join(HERE, "variants")
    os.makedirs(os.path.join(out_dir, "build_" + tag), exist_ok=True)
    out = os.path.join(out_dir, f"libza_b200_{tag}.so")
    procs, objs = [], []
    for s in SOURCES:
        o = os.path.join(out_dir, "build_" + tag, s + ".o")
        cmd = [_nvcc()] + NVCC_FLAGS + ["-D" + d for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, s), "-o", o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    for s, p in procs:
        o, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed for {s} ({tag}):\n{o}")
        if verbose:
            sys.stderr.write(o)
    subprocess.check_call([_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out] + objs + ["-lcudart"])
    return out


def build(force=False, verbose=False):
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    if not force and not needs_build():
        fe = [os.path.join(CSRC, "za2c.cpp")] + [os.path.join(CSRC, "frontend", f) for f in os.listdir(os.path.join(CSRC, "frontend"))]
        if not os.path.exists(OUT_ZA2C) or any(os.path.getmtime(OUT_ZA2C) < os.path.getmtime(f) for f in fe):
            build_za2c()
        return OUT
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for s in srcs:
        o = os.path.join(HERE, "build", os.path.basename(s) + ".o")
        extra = os.environ.get("ZA_NVCC_EXTRA", "").split()
        cmd = [_nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"nvcc failed for {s}:\n{out}\n")
        elif verbose:
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("nvcc compilation of za_b200 failed")
    subprocess.check_call([_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", OUT] + objs + ["-lcudart"])
    build_za2c()
    return OUT


def build_za2c():
    """libza2c.so: the outer C ABI of za's bindings (include/za2c.h) on top of libza_b200.so (host code only)."""
    srcs = [os.path.join(CSRC, "za2c.cpp")] + [os.path.join(CSRC, "frontend", f) for f in FRONTEND_SOURCES]
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-o", OUT_ZA2C] + srcs + ["-L" + HERE, "-lza_b200",
                           "-Wl,-rpath,$ORIGIN"])
    return OUT_ZA2C


if __name__ == "__main__":
    if "--variant" in sys.argv:       # python -m za_b200.build --variant TAG DEFINE [DEFINE ...]
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], [a for a in sys.argv[i + 2:] if not a.startswith("-")], verbose="-v" in sys.argv))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
