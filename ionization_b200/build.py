"""Build the CUDA engine in-tree:  python -m ionization_b200.build

nvcc cross-compiles for sm_100a without a GPU.  Output: ionization_b200/_lib/libionization_b200.so (git-ignored,
travels to the GPU box with the repo snapshot).
"""
import os
import shutil
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = [os.path.join(_HERE, "csrc", "engine.cu")]
DEPS = SRC + [os.path.join(_HERE, "csrc", n) for n in sorted(os.listdir(os.path.join(_HERE, "csrc"))) if n.endswith(".cuh")] + [
    os.path.join(os.path.dirname(_HERE), "include", "ionization_b200.h")
]
OUT = os.path.join(_HERE, "_lib", "libionization_b200.so")

NVCC_FLAGS = [
    "-O3",
    "-std=c++17",
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler",
    "-fPIC",
    "-shared",
]


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def up_to_date():
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(d) <= t for d in DEPS)


def build(force=False, verbose=False):
    if not force and up_to_date():
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = [find_nvcc(), *NVCC_FLAGS, *(["-Xptxas", "-v"] if verbose else []), "-o", OUT, *SRC]
    env = dict(os.environ)
    # the image's default CC wrapper lacks some spec files; nvcc is happiest with the system g++
    if os.path.exists("/usr/bin/g++"):
        cmd[1:1] = ["-ccbin", "/usr/bin/g++"]
    res = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stderr[-4000:])
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
