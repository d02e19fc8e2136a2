"""Build libmp3b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmp3b200.so")
SOURCES = ["mp3_encoder.cu", "mp3_config.cpp", "mp3_tag.cpp", "mp3_id3.cpp"]
DEPS = SOURCES + ["mp3_config.h", "mp3_device.cuh", "mp3_math.cuh", "mp3_tables.h", "k_filterbank.cuh", "k_psy.cuh",
                  "k_quant.cuh", "k_tag.cuh", "mp3_tag.h", "mp3_handle.inc", "../../include/mp3b200.h"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    # bit-exactness contract: no FMA contraction, IEEE div/sqrt, no flush-to-zero (DESIGN.md "numerics")
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math", "-diag-suppress", "222", "-shared",
]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force=False, verbose=False, variant=None, defines=()):
    """variant/defines: tuning builds (libmp3b200_<variant>.so with -D flags), selected at run time by MP3B200_LIB."""
    out = LIB if variant is None else os.path.join(HERE, "libmp3b200_%s.so" % variant)
    if variant is None and not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-D" + d for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-o", out] + SOURCES
    subprocess.check_call(cmd, cwd=CSRC)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
