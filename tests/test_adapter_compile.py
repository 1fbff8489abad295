"""CPU tier: the gdcs adapter (adapters/gdcs_cuda.cpp) compiles against the REFERENCE'S OWN, unmodified gdcs.h and
godot-cpp, and so do the reference's callers of that class.  godot-cpp's class headers are generated offline from the
extension_api.json the reference vendors; nothing is linked or run (there is no Godot binary here):
the check is that one file swapped in -- gdcs.cpp -> gdcs_cuda.cpp -- leaves every other source of the extension as is.
Skipped where /root/reference is absent (the GPU box)."""
import os
import subprocess
import sys

import pytest

REPO = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
REF = os.environ.get("GDPT_REFERENCE", "/root/reference")
GODOT_CPP = os.path.join(REF, "godot-cpp")

pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(GODOT_CPP, "gdextension", "extension_api.json")),
                                reason="the reference tree (with its vendored godot-cpp) is not on this machine")


@pytest.fixture(scope="module")
def godot_headers(tmp_path_factory):
    out = tmp_path_factory.mktemp("godot_cpp_gen")
    gen = ("import sys; sys.path.insert(0, %r); import binding_generator as bg; "
           "bg.generate_bindings(%r, True, '64', 'single', %r)" % (GODOT_CPP, os.path.join(GODOT_CPP, "gdextension", "extension_api.json"), str(out)))
    subprocess.run([sys.executable, "-c", gen], check=True, capture_output=True, timeout=300)
    assert os.path.exists(out / "gen" / "include" / "godot_cpp" / "classes" / "rendering_device.hpp")
    return [f"-I{GODOT_CPP}/include", f"-I{out}/gen/include", f"-I{GODOT_CPP}/gdextension"]


def syntax_check(source, includes):
    cmd = ["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Wno-unused-parameter", "-Wno-unused-variable", "-Wno-sign-compare",
           "-Wno-reorder", "-Wno-unused-but-set-variable"] + includes + [source]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=600)


def test_adapter_implements_the_unmodified_gdcs_header(godot_headers):
    inc = godot_headers + [f"-I{REF}/src/gdcs/include", f"-I{REPO}/include"]
    r = syntax_check(os.path.join(REPO, "adapters", "gdcs_cuda.cpp"), inc)
    assert r.returncode == 0, r.stderr[-4000:]


def test_adapter_defines_every_method_gdcs_h_declares():
    """Every ComputeShader member function of gdcs.h has a definition in the adapter (a missing one would only show at link time)."""
    import re
    header = open(os.path.join(REF, "src", "gdcs", "include", "gdcs.h")).read()
    body = re.sub(r"//[^\n]*", "", header[header.index("class ComputeShader"):])  # commented-out declarations do not count
    declared = set(re.findall(r"\b(~?\w+)\s*\([^;{]*\)\s*(?:const)?\s*;", body))
    declared.discard("static_assert")
    adapter = open(os.path.join(REPO, "adapters", "gdcs_cuda.cpp")).read()
    defined = set(re.findall(r"ComputeShader::(~?\w+)\s*\(", adapter))
    assert declared and declared <= defined, f"not defined by the adapter: {sorted(declared - defined)}"


@pytest.mark.parametrize("source", ["src/path_tracing/path_tracing_camera.cpp", "src/path_tracing/post_processing/progressive_rendering.cpp",
                                    "src/path_tracing/post_processing/temporal_reprojection.cpp"])
def test_reference_callers_compile_unmodified_beside_it(godot_headers, source):
    """The reference's users of ComputeShader need nothing but gdcs.h: they compile as they are."""
    inc = godot_headers + [f"-I{REF}/src/gdcs/include", f"-I{REF}/src", f"-I{REF}/src/path_tracing", f"-I{REF}/src/bvh",
                           f"-I{REF}/src/path_tracing/post_processing"]
    r = syntax_check(os.path.join(REF, source), inc)
    assert r.returncode == 0, r.stderr[-4000:]
