#!/usr/bin/env python3
"""Fill in the reference's include/ginkgo/config.hpp.in the way its cmake configure step
would for: CPU-only (reference+omp), no MPI/HWLOC/PAPI/TAU/VTUNE/METIS/ROCTX, half and
bfloat16 disabled, mixed precision off.  TEST INFRASTRUCTURE ONLY; writes to stdout."""
import re, sys
txt = open(sys.argv[1]).read()
subst = {
    "Ginkgo_VERSION_MAJOR": "1", "Ginkgo_VERSION_MINOR": "12", "Ginkgo_VERSION_PATCH": "0",
    "Ginkgo_VERSION_TAG": "develop", "GINKGO_VERBOSE_LEVEL": "1",
    "GINKGO_DPCPP_MAJOR_VERSION": "0", "GINKGO_DPCPP_MINOR_VERSION": "0",
    "GINKGO_HAVE_PAPI_SDE": "0", "GINKGO_HAVE_TAU": "0", "GINKGO_HAVE_VTUNE": "0",
    "GINKGO_HAVE_METIS": "0", "GINKGO_HAVE_ROCTX": "0", "GINKGO_HAVE_HWLOC": "0",
    "METIS_HEADER": "metis.h",
}
defined = {"GKO_HAVE_CXXABI_H", "GKO_SIZE_T_IS_UINT64_T"}
for k, v in subst.items():
    txt = txt.replace("@%s@" % k, v)
txt = re.sub(r"#cmakedefine01 (\w+)", lambda m: "#define %s %d" % (m.group(1), 1 if m.group(1) in defined else 0), txt)
txt = re.sub(r"#cmakedefine (\w+)", lambda m: ("#define %s" % m.group(1)) if m.group(1) in defined else "/* #undef %s */" % m.group(1), txt)
assert "@" not in re.sub(r"//.*", "", txt), "unsubstituted cmake variable left"
sys.stdout.write(txt)
