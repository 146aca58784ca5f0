"""The C-ABI shared library builds for sm_100a, loads on a CPU-only box, exports every symbol that
include/coalgpu.h declares, and refuses to compute without a GPU (no CPU fallback)."""
import os
import re
import subprocess

import pytest

from conftest import ROOT, has_gpu, load_golden


def declared_functions():
    text = open(os.path.join(ROOT, "include", "coalgpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(coalgpu_[a-z0-9_]+)\s*\(", text)))


def test_exports_match_header(cuda_lib):
    names = declared_functions()
    assert len(names) >= 12
    for n in names:
        assert hasattr(cuda_lib, n), f"libcoalgpu.so does not export {n}"
    from coalescenator_b200.sampler import EXPORTS
    assert sorted(EXPORTS) == names


def test_sass_is_sm100a_only():
    import subprocess
    from coalescenator_b200.build import LIB, build_cuda
    build_cuda()
    out = subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_gpu(cuda_lib):
    import coalescenator_b200 as cb
    p, _, _ = load_golden("cfg1_pair0")
    with pytest.raises(cb.CoalGpuError) as e:
        cb.ImportanceSampler().run_sampler(p, 10)
    assert e.value.code == -2          # COALGPU_ENODEV
    with pytest.raises(cb.CoalGpuError):
        cb.DeviceSampler(p)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under coalescenator_b200/ may reference it."""
    pkg = os.path.join(ROOT, "coalescenator_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle_py" not in src and "coal_oracle" not in src and "liboracle" not in src, os.path.join(dirpath, f)


BINDING = os.path.join(ROOT, "oracle", "_ref", "coalref_gpu")


@pytest.mark.skipif(not os.path.exists(BINDING), reason="oracle/_ref/coalref_gpu not built (needs /root/reference)")
@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU failure mode")
def test_reference_side_binding_links_and_fails_loudly_without_gpu():
    """include/GpuImportanceSampler.hpp compiled against the reference's own headers and linked into the reference
    driver in place of ImportanceSampler (oracle/Makefile): it must link, and without a GPU it must raise, not fall
    back to the CPU sampler that is linked into the same binary."""
    r = subprocess.run([BINDING, os.path.join(ROOT, "tests", "golden", "cfg1_pair0.coalprob"), "10", "1", "1"],
                       capture_output=True, text=True)
    assert r.returncode != 0
    assert "coalgpu: no CUDA device" in r.stderr
    assert "likelihood" not in r.stdout
