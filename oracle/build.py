"""Build the CPU oracle (test infrastructure).

  oracle/_build/liboracle.so   <- oracle/jpeg_oracle.c           (our restatement)
  oracle/_build/libljt.so      <- oracle/ljt_shim.c              (libjpeg-turbo harness)
  oracle/_ref/libref_parser.so <- /root/reference/src/rocjpeg_parser.cpp + oracle/ref_parser_shim.cpp
  oracle/_ref/libref_kernels.so<- /root/reference/src/rocjpeg_hip_kernels.cpp compiled for the CPU
                                  through oracle/ref_hip_emul/ (a HIP-on-CPU emulation header)

The two _ref libraries are compiled from the reference sources WHERE THEY LIE
(nothing is copied into this repository); they exist only when /root/reference
is present (this container). On the GPU box the prebuilt files travel with the
snapshot. Run: python oracle/build.py
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("ROCJPEG_REFERENCE", "/root/reference")
BUILD = os.path.join(HERE, "_build")
REFOUT = os.path.join(HERE, "_ref")


def _newer(src_list, out):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(s) > t for s in src_list if os.path.exists(s))


def _run(cmd, **kw):
    r = subprocess.run(cmd, capture_output=True, text=True, **kw)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd if isinstance(cmd, list) else [cmd]) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("oracle build failed")


def build(verbose=False):
    os.makedirs(BUILD, exist_ok=True)
    src = os.path.join(HERE, "jpeg_oracle.c")
    out = os.path.join(BUILD, "liboracle.so")
    if _newer([src], out):
        _run(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", out, src, "-lm"])
    src = os.path.join(HERE, "ljt_shim.c")
    out = os.path.join(BUILD, "libljt.so")
    if _newer([src], out):
        _run(["gcc", "-O2", "-shared", "-fPIC", "-o", out, src, "-ldl", "-lpthread"])
    if os.path.isdir(os.path.join(REF, "src")):
        os.makedirs(REFOUT, exist_ok=True)
        # (1) the reference parser, unmodified, behind a C shim
        shim = os.path.join(HERE, "ref_parser_shim.cpp")
        rsrc = os.path.join(REF, "src", "rocjpeg_parser.cpp")
        out = os.path.join(REFOUT, "libref_parser.so")
        if _newer([shim, rsrc], out):
            _run(["g++", "-O2", "-std=c++17", "-w", "-shared", "-fPIC", "-I", os.path.join(REF, "src"),
                  "-o", out, rsrc, shim])
        # (2) the reference HIP kernels, compiled for the CPU. g++ cannot parse the
        # <<<grid, block, shmem, stream>>> launch syntax, so the source is piped
        # through sed (launch -> EMUL_LAUNCH macro) straight into the compiler;
        # no transformed copy is written to disk.
        shim = os.path.join(HERE, "ref_kernels_shim.cpp")
        emul = os.path.join(HERE, "ref_hip_emul")
        rsrc = os.path.join(REF, "src", "rocjpeg_hip_kernels.cpp")
        out = os.path.join(REFOUT, "libref_kernels.so")
        if _newer([shim, rsrc, os.path.join(emul, "hip", "hip_runtime.h")], out):
            obj = os.path.join(REFOUT, "ref_kernels.o")
            sed = (r"sed -E ':a;N;$!ba;s/([A-Za-z0-9_]+)<<<([^;]*)>>>\(/EMUL_LAUNCH(\1, \2)(/g' " + rsrc)
            cmd = (sed + " | g++ -O2 -std=c++17 -w -fPIC -ffp-contract=off -x c++ -c -I " + emul + " -I " +
                   os.path.join(REF, "src") + " -o " + obj + " -")
            _run(cmd, shell=True)
            _run(["g++", "-O2", "-std=c++17", "-w", "-shared", "-fPIC", "-ffp-contract=off", "-I", emul, "-I",
                  os.path.join(REF, "src"), "-o", out, obj, shim])
            os.remove(obj)
    if verbose:
        print("oracle built:", os.listdir(BUILD), os.listdir(REFOUT) if os.path.isdir(REFOUT) else [])


if __name__ == "__main__":
    build(verbose=True)
