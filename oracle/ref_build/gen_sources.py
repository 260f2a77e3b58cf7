#!/usr/bin/env python3
"""Emit sources.mk for oracle/ref_build/Makefile.

TEST INFRASTRUCTURE ONLY.  Reads the *lists of source files* that make up the
reference's CPU libraries (core + reference + omp + device glue + the stub
hooks for the disabled cuda/hip/dpcpp backends) out of the reference's
CMakeLists.txt files where they lie under /root/reference.  No reference file
is copied; cmake is never run.  The MPI / PAPI / METIS conditional blocks are
dropped (those dependencies are absent in this image).
"""
import re, sys, os

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"

def strip_blocks(txt, conds):
    for c in conds:
        txt = re.sub(r"if\(%s\)(.*?)endif\(\)" % re.escape(c), "", txt, flags=re.S)
    return txt

def cpp_tokens(txt):
    return [t for t in re.findall(r"[\w./${}]+\.cpp", txt)]

def listed(cmakelists, base, conds=()):
    txt = open(os.path.join(REF, cmakelists)).read()
    txt = strip_blocks(txt, conds)
    # only the first target_sources(...) / set(UNIFIED_SOURCES ...) list matters
    out = []
    for t in cpp_tokens(txt):
        if "$" in t:
            continue
        p = os.path.normpath(os.path.join(base, t))
        if os.path.exists(os.path.join(REF, p)) and p not in out:
            out.append(p)
    return out

core = listed("core/CMakeLists.txt", "core",
              ("GINKGO_BUILD_MPI", "GINKGO_HAVE_PAPI_SDE", "GINKGO_HAVE_METIS"))
core = [c for c in core if "/test/" not in c]
reference = listed("reference/CMakeLists.txt", "reference")
omp = listed("omp/CMakeLists.txt", "omp")
unified = listed("common/unified/CMakeLists.txt", "common/unified")
unified.append("common/unified/matrix/dense_kernels.instantiate.cpp")
hooks = ["core/device_hooks/cuda_hooks.cpp", "core/device_hooks/hip_hooks.cpp",
         "core/device_hooks/dpcpp_hooks.cpp"]
devices = ["devices/machine_topology.cpp", "devices/device.cpp",
           "devices/cuda/executor.cpp", "devices/hip/executor.cpp",
           "devices/dpcpp/executor.cpp", "devices/omp/executor.cpp",
           "devices/reference/dummy.cpp"]

def emit(name, lst):
    print("%s := \\\n  %s\n" % (name, " \\\n  ".join(lst)))

emit("CORE_SRCS", core + hooks + devices)
emit("REFERENCE_SRCS", reference)
emit("OMP_SRCS", omp + unified)
