"""The reference-side shim (integration/CudaLogicRenderer.cpp) is compiled against stub engine headers
and linked against the C-ABI library: syntax, types and every gk_* symbol it uses are checked without Vulkan or a GPU."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INT = os.path.join(ROOT, "integration")
LIB = os.path.join(ROOT, "gknextrenderer_b200", "lib")


def test_shim_compiles_and_links_against_the_c_abi(built, tmp_path):
    exe = str(tmp_path / "shim_link_check")
    cmd = ["g++", "-std=c++20", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(INT, "stubs"), "-I", os.path.join(ROOT, "include"),
           os.path.join(INT, "CudaLogicRenderer.cpp"), os.path.join(INT, "shim_link_check.cpp"), "-o", exe, "-L", LIB, "-lgknext_cuda", f"-Wl,-rpath,{LIB}"]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-3000:]
    run = subprocess.run([exe], capture_output=True, text=True)  # constructs the renderer object only (no GPU call)
    assert run.returncode == 0 and "shim linked" in run.stdout, run.stderr[-2000:]
