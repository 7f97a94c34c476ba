"""Builds libfqgrep_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libfqgrep_b200.so")
CLI = os.path.join(HERE, "fqgrep-b200")      # the reference's command line over the device path (csrc/cli.cc)
SOURCES = ["api.cu", "fastq_pipeline.cu", "stream.cu", "match_soa.cu", "match_fastq.cu", "synth.cu", "regex_compile.cc", "ac_build.cc", "gram_filter.cc", "names.cc",
           "fastq_host.cc", "input_host.cc", "reader_host.cc"]
HOST_SOURCES = ["fastq_cut.cc"]      # plain g++: per-function target attributes (AVX2 newline counting with a run-time check)
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--use_fast_math",
              "-Xcompiler", "-fPIC,-O3,-Wall", "-Xptxas", "-v", "--expt-relaxed-constexpr"]


def _stale():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    if not os.path.exists(CLI):
        return True
    t = min(t, os.path.getmtime(CLI))
    inc = os.path.join(HERE, "..", "include")
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(inc, "fqgrep_b200.h"), os.path.join(inc, "fqgrep_b200.hpp"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not _stale():
        return SO
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src + ".o")
        objs.append(obj)
        cmd = [nvcc] + NVCC_FLAGS + ["-x", "cu", "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    cxx = os.environ.get("CXX", "g++")
    for src in HOST_SOURCES:
        obj = os.path.join(objdir, src + ".o")
        objs.append(obj)
        cmd = [cxx, "-O3", "-std=c++17", "-fPIC", "-Wall", "-pthread", "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, p in procs:
        out, _ = p.communicate()
        log.append("== %s\n%s" % (src, out))
        if p.returncode != 0:
            sys.stderr.write("\n".join(log))
            raise RuntimeError("nvcc failed on %s" % src)
    cmd = [nvcc, "-shared", "-o", SO] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-lz"]
    subprocess.check_call(cmd)
    # host-only C++ above the C ABI: links the library by relative rpath so the pair travels together
    subprocess.check_call([cxx, "-O2", "-std=c++17", "-Wall", "-I", os.path.join(HERE, "..", "include"), os.path.join(CSRC, "cli.cc"),
                           "-o", CLI, "-L", HERE, "-lfqgrep_b200", "-Wl,-rpath,$ORIGIN"])
    with open(os.path.join(objdir, "build.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    return SO


def build_variant(name, defines, sources=("match_fastq.cu",)):
    """Measurement hook: the same library with `sources` recompiled under extra -D flags, as libfqgrep_b200_<name>.so
    (loaded through FQG_LIB_VARIANT).  Used for A/B runs of kernel parameters inside one GPU call."""
    build()
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objdir = os.path.join(HERE, "build")
    objs = []
    for src in SOURCES + HOST_SOURCES:
        obj = os.path.join(objdir, src + ".o")
        if src in sources:
            obj = os.path.join(objdir, "%s.%s.o" % (src, name))
            subprocess.check_call([nvcc] + NVCC_FLAGS + ["-D" + d for d in defines] + ["-x", "cu", "-c", os.path.join(CSRC, src), "-o", obj],
                                  stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        objs.append(obj)
    so = os.path.join(HERE, "libfqgrep_b200_%s.so" % name)
    subprocess.check_call([nvcc, "-shared", "-o", so] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-lz"])
    return so


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print(SO)
