"""the adaptor a maintainer drops into the reference tree (include/B200MatrixSolver.h, a
solver::AbstractSolver) is compiled here against the reference's OWN headers (srcAna/Solver.h, Run.h, Geometry.h,
Excitation.h, Tools.h ...) behind the stand-ins of oracle/stub (Types.h, the Eigen container, a three-typedef hdf5.h),
and linked against the C ABI: the binding is proven to match both sides' declarations.  Skipped where /root/reference
does not exist (the GPU box)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/srcAna"

pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "Solver.h")), reason="reference sources absent")

TU = r'''
#include "B200MatrixSolver.h"
// what srcAna/Solver.cpp:30-34 becomes in a serial build
std::shared_ptr<optimet::solver::AbstractSolver> b200_factory(optimet::Run const &run) {
  return std::make_shared<optimet::solver::B200Matrix>(run, std::vector<int>{0});
}
'''


def test_adaptor_compiles_against_the_reference_headers(tmp_path):
    src = tmp_path / "adaptor_tu.cpp"
    src.write_text(TU)
    obj = tmp_path / "adaptor_tu.o"
    cmd = ["g++", "-std=c++11", "-Wall", "-Wno-unused-variable", "-c", str(src), "-o", str(obj),
           "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "oracle", "stub"),
           "-I" + os.path.join(ROOT, "oracle", "stub", "adaptor"), "-I" + REF]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]
    # every C-ABI symbol the adaptor references is exported by the built library
    nm = subprocess.run(["nm", "-u", str(obj)], capture_output=True, text=True).stdout
    used = sorted({ln.split()[-1] for ln in nm.splitlines() if ln.split() and ln.split()[-1].startswith("ob_")})
    assert {"ob_create_multi", "ob_multi_run", "ob_multi_set_cluster", "ob_multi_set_frequency", "ob_multi_set_incident",
            "ob_multi_set_option", "ob_destroy_multi"} <= set(used)
    lib = os.path.join(ROOT, "optimet_b200", "liboptimet_b200.so")
    exported = subprocess.run(["nm", "-D", "--defined-only", lib], capture_output=True, text=True).stdout
    for sym in used:
        assert (" T " + sym) in exported, sym
