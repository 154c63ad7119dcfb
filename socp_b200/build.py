"""Build libsocp_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libsocp_b200.so")
# the parity build of SURVEY.md section 8(d): the same sources without FMA contraction (-fmad=false), to report the
# trajectory parity of both builds side by side (selected at run time with SOCP_LIB=<path>, tools/parity_builds.py)
SO_NOFMA = os.path.join(HERE, "libsocp_b200_nofma.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]


def sources():
    out = [os.path.join(HERE, "..", "include", "socp_b200.h")]
    for f in sorted(os.listdir(CSRC)):
        out.append(os.path.join(CSRC, f))
    return out


def stale(so=SO):
    if not os.path.exists(so):
        return True
    t = os.path.getmtime(so)
    return any(os.path.getmtime(s) > t for s in sources())


def build_variant(name, extra):
    """An A/B build with extra nvcc flags -> socp_b200/variants/<name>.so (select with SOCP_LIB)."""
    d = os.path.join(HERE, "variants")
    os.makedirs(d, exist_ok=True)
    so = os.path.join(d, name + ".so")
    cmd = [os.environ.get("NVCC", "nvcc")] + NVCC_FLAGS + list(extra) + ["-o", so, os.path.join(CSRC, "api.cu")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building " + so)
    return so


def build(force=False, verbose=False, extra=(), nofma=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ... -> socp_b200/libsocp_b200.so
    (nofma=True: the -fmad=false parity build -> socp_b200/libsocp_b200_nofma.so)"""
    so = SO_NOFMA if nofma else SO
    if not force and not stale(so):
        return so
    nvcc = os.environ.get("NVCC", "nvcc")
    extra = list(extra) + os.environ.get("SOCP_NVCC_EXTRA", "").split()      # experiments: -DSOCP_...=...
    if nofma:
        extra = extra + ["-fmad=false"]
    cmd = [nvcc] + NVCC_FLAGS + list(extra) + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", so, os.path.join(CSRC, "api.cu")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libsocp_b200.so")
    if verbose:
        print(r.stderr)
    return so


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, nofma="--nofma" in sys.argv))
