"""The C-ABI libraries load on a CPU-only box and export every symbol the headers declare;
wire structs have the reference's sizes; without a GPU the backend fails loudly (no fallback)."""
import ctypes
import os
import re

import pytest

from conftest import REPO, has_gpu

INCLUDE = os.path.join(REPO, "include")


def declared(header):
    text = open(os.path.join(INCLUDE, header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"GDPT_API\s+[\w\s\*]+?\b(gdpt_\w+)\s*\(", text)))


def test_cuda_library_exports_every_declared_symbol():
    names = declared("gdpt.h")
    assert len(names) >= 25
    lib = ctypes.CDLL(os.path.join(REPO, "gdpathtracing_b200", "libgdpt_cuda.so"))
    for n in names:
        assert hasattr(lib, n), f"libgdpt_cuda.so does not export {n}"
    from gdpathtracing_b200 import _lib
    assert sorted(n for n, _, _ in _lib.CUDA_API) == names, "ctypes table and gdpt.h disagree"


def test_host_library_exports_every_declared_symbol():
    names = declared("gdpt_host.h")
    lib = ctypes.CDLL(os.path.join(REPO, "gdpathtracing_b200", "libgdpt_host.so"))
    for n in names:
        assert hasattr(lib, n), f"libgdpt_host.so does not export {n}"
    from gdpathtracing_b200 import _lib
    assert sorted(n for n, _, _ in _lib.HOST_API) == names


def test_wire_struct_sizes_match_reference_layout(tmp_path):
    """Compile a C probe against include/gdpt_wire.h (its static asserts are the check)."""
    import subprocess
    src = tmp_path / "probe.c"
    src.write_text('#include "gdpt_wire.h"\n#include "gdpt.h"\n#include "gdpt_host.h"\n#include <stdio.h>\n'
                   "int main(void){printf(\"%zu %zu %zu %zu %zu %zu %zu %zu\\n\", sizeof(gdpt_render_params), sizeof(gdpt_camera),"
                   "sizeof(gdpt_bvh_node), sizeof(gdpt_tlas_node), sizeof(gdpt_blas_instance), sizeof(gdpt_triangle_geometry),"
                   "sizeof(gdpt_triangle_data), sizeof(gdpt_material));return 0;}\n")
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-std=c11", "-I", INCLUDE, "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert [int(x) for x in out] == [36, 160, 48, 32, 176, 48, 80, 64]


def test_reference_struct_sizes_agree():
    from oracle import oracle
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    r = oracle.ref()
    assert [r.refbvh_sizeof(i) for i in range(4)] == [48, 144, 176, 32]


@pytest.mark.skipif(has_gpu(), reason="CPU-only behaviour")
def test_no_gpu_means_loud_failure_not_fallback():
    from gdpathtracing_b200 import PathTracingCamera, _lib, scenes
    dev = ctypes.c_void_p()
    rc = _lib.cuda.gdpt_device_create(0, ctypes.byref(dev))
    assert rc == -1 and not dev.value
    assert b"no CPU path" in _lib.cuda.gdpt_last_error(None)
    cam = PathTracingCamera()
    cam.geometry_group = scenes.populate(scenes.cornell32())
    cam.set_window_size(64, 64)
    with pytest.raises(_lib.GdptError):
        cam.init()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(REPO, "gdpathtracing_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(root, f), errors="ignore").read()
                assert "oracle" not in text.replace("the oracle", "").replace("an oracle", "") or f in ("pt_math.cuh",), \
                    f"{f} mentions the oracle package"


def test_reference_arm_of_bench_runs_without_a_gpu():
    """`bench.py --impl reference` (the oracle on host threads) is what the driver times beside our arm: it must run on
    a machine without a GPU and print exactly one JSON line with the contract's keys."""
    import json
    import subprocess
    import sys
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(repo, "bench.py"), "--impl", "reference", "--scene", "cornell32", "--width", "64",
                          "--height", "64", "--depth", "4", "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "Mrays/s" and line["value"] > 0
    from oracle import oracle
    # the reference's own shader text compiled as C++ where that library exists, our restatement of it otherwise
    assert line["cpu_baseline"]["kind"] == ("reference" if oracle.ref_shader_available() else "port")
    assert line["cpu_baseline"]["cores"] >= 1 and set(line["config"]) == {"workload", "partition"}
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
