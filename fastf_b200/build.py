"""Builds the native pieces in-tree (no JIT cache): the sm_100a CUDA library behind include/fastf_gpu.h,
the synthetic-data generator and (test infrastructure only) the oracle.  Used by __graft_entry__.build()."""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "fastf_b200")
BUILD = os.path.join(PKG, "_build")
LIB = os.path.join(BUILD, "libfastf_gpu.so")
SYNTH_LIB = os.path.join(BUILD, "libfastf_synth.so")
SYNTH_BIN = os.path.join(BUILD, "fastf_synth")
CLI_BIN = os.path.join(BUILD, "fastF")
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC", "-shared"]


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("build step failed: " + " ".join(cmd))
    return r.stdout


def _sources(d, exts):
    return [os.path.join(d, f) for f in sorted(os.listdir(d)) if f.endswith(exts)]


def source_hash():
    """sha256 (first 16 hex digits) over the kernel sources and the C-ABI header: baked into the library (fastf_build_info) so that a
    stale prebuilt binary is noticed at load time"""
    import hashlib
    csrc = os.path.join(PKG, "csrc")
    h = hashlib.sha256()
    for f in _sources(csrc, (".cu", ".cuh", ".h")) + [os.path.join(ROOT, "include", "fastf_gpu.h")]:
        h.update(os.path.basename(f).encode() + b"\0" + open(f, "rb").read())
    return h.hexdigest()[:16]


def build_cuda(force=False):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    csrc = os.path.join(PKG, "csrc")
    deps = _sources(csrc, (".cu", ".cuh", ".h")) + [os.path.join(ROOT, "include", "fastf_gpu.h")]
    if not force and _newer(LIB, deps):
        return LIB
    if not os.path.exists(nvcc):
        if os.path.exists(LIB):
            sys.stderr.write("fastf_b200.build: nvcc not found, keeping the prebuilt libfastf_gpu.so although sources are newer (the source hash is checked at load time)\n")
            return LIB
        raise RuntimeError("nvcc not found and no prebuilt libfastf_gpu.so")
    os.makedirs(BUILD, exist_ok=True)
    _run([nvcc] + NVCC_FLAGS + ['-DFASTF_SRC_HASH="%s"' % source_hash(), "-o", LIB, os.path.join(csrc, "capi.cu"), os.path.join(csrc, "sharded.cu"), "-ldl"])
    return LIB


def build_synth(force=False):
    src = os.path.join(PKG, "synth", "synth.c")
    deps = [src, os.path.join(PKG, "synth", "synth.h")]
    os.makedirs(BUILD, exist_ok=True)
    if force or not _newer(SYNTH_LIB, deps):
        _run(["gcc", "-O2", "-fPIC", "-shared", "-o", SYNTH_LIB, src, "-lz", "-lpthread", "-lm"])
    if force or not _newer(SYNTH_BIN, deps):
        _run(["gcc", "-O2", "-DFASTF_SYNTH_MAIN", "-o", SYNTH_BIN, src, "-lz", "-lpthread", "-lm"])
    return SYNTH_LIB


def build_cli(force=False, emu=False):
    """The C host (`fastF bam2db|freq`).  emu=True links the SIMT-emulator build of the kernels instead (tests only)."""
    host = os.path.join(PKG, "host")
    srcs = _sources(host, (".c",))
    if not srcs:
        return None
    deps = srcs + _sources(host, (".h",)) + [os.path.join(ROOT, "include", "fastf_gpu.h")]
    if emu:
        libdir, libname, out = os.path.join(ROOT, "tests", "emu", "_build"), "fastf_emu", os.path.join(ROOT, "tests", "emu", "_build", "fastF_emu")
        deps.append(build_emu())
    else:
        libdir, libname, out = BUILD, "fastf_gpu", CLI_BIN
        deps.append(LIB)
    if force or not _newer(out, deps):
        _run(["gcc", "-O2", "-std=gnu11", "-Wall", "-Wno-unused-result", "-I" + os.path.join(ROOT, "include"), "-o", out] + srcs +
             ["-L" + libdir, "-l" + libname, "-Wl,-rpath,$ORIGIN", "-lz", "-l:libsqlite3.so.0", "-lm", "-ldl"])
    return out


def build_oracle():
    """Test infrastructure: the CPU restatement and, when /root/reference exists, the unmodified reference."""
    _run(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    # INTEGRATION.md section B as a test target: the reference's own main.c linked against this repository's bam2db()
    _run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref_gpu"])


def build_ref_main_emu():
    """Test infrastructure: reference main.c + our host, linked against the SIMT-emulator build (boxes without a GPU).  Returns the path or None."""
    build_emu()
    _run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref_emu"])
    p = os.path.join(ROOT, "oracle", "_ref", "fastF_gpu_emu")
    return p if os.path.exists(p) else None


def build_emu():
    """Test infrastructure: the same kernel sources compiled for the host under the SIMT emulator."""
    out = os.path.join(ROOT, "tests", "emu", "_build", "libfastf_emu.so")
    csrc = os.path.join(PKG, "csrc")
    deps = _sources(csrc, (".cu", ".cuh", ".h")) + [os.path.join(ROOT, "include", "fastf_gpu.h"), os.path.join(ROOT, "tests", "emu", "cuda_emu.h")]
    if _newer(out, deps):
        return out
    os.makedirs(os.path.dirname(out), exist_ok=True)
    _run(["g++", "-O2", "-std=c++17", "-DFASTF_EMU", '-DFASTF_SRC_HASH="%s"' % source_hash(), "-I" + os.path.join(ROOT, "tests", "emu"), "-x", "c++", "-fPIC", "-shared", "-o", out,
          os.path.join(csrc, "capi.cu"), os.path.join(csrc, "sharded.cu"), "-lz", "-lpthread"])
    return out


def build_all(force=False):
    build_cuda(force)
    build_synth(force)
    build_cli(force)
    build_oracle()


if __name__ == "__main__":
    build_all("--force" in sys.argv)
    print("built", LIB)
