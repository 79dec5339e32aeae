"""The reference-side adapter of INTEGRATION.md (integration/CudaMultiplexRenderer.h) is real code:
it compiles against the reference's own headers, binds only exported C-ABI symbols, links and runs.
Needs the reference checkout (headers + oracle/_ref/libxnref_model.so for Grid's constructor), so it
runs in the build container and is skipped on the GPU box."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/src"
REF_MODEL = os.path.join(ROOT, "oracle", "_ref", "libxnref_model.so")
LIB_DIR = os.path.join(ROOT, "xenodon_b200")

pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")


def _compile(tmp_path):
    obj = tmp_path / "adapter_check.o"
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Wno-unused-parameter",
           "-I", os.path.join(ROOT, "integration", "shim"), "-I", os.path.join(ROOT, "integration"),
           "-I", os.path.join(ROOT, "include"), "-isystem", REF,
           "-c", os.path.join(ROOT, "integration", "adapter_check.cpp"), "-o", str(obj)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]
    return obj


def test_adapter_compiles_against_reference_headers_and_binds_exported_symbols(tmp_path):
    import xenodon_b200 as xb

    obj = _compile(tmp_path)
    undefined = subprocess.run(["nm", "-u", str(obj)], capture_output=True, text=True, check=True).stdout.split()
    bound = sorted(s for s in undefined if s.startswith("xn_"))
    assert {"xn_ctx_create", "xn_upload_grid", "xn_upload_svo", "xn_set_target", "xn_set_params", "xn_render",
            "xn_sync", "xn_frame_gather", "xn_ctx_destroy", "xn_traversal_from_name"} <= set(bound)
    lib = xb.lib()
    for name in bound:
        assert hasattr(lib, name), f"adapter binds {name}, which the library does not export"


def test_adapter_links_and_fails_loudly_without_a_device(tmp_path):
    if not os.path.exists(REF_MODEL):
        pytest.skip("oracle/_ref/libxnref_model.so not built (make -C oracle ref)")
    obj = _compile(tmp_path)
    exe = tmp_path / "adapter_check"
    r = subprocess.run(["/usr/bin/g++", "-o", str(exe), str(obj), "-L", LIB_DIR, "-lxenodon_b200", REF_MODEL,
                        f"-Wl,-rpath,{LIB_DIR}", f"-Wl,-rpath,{os.path.dirname(REF_MODEL)}",
                        "-Wl,--allow-shlib-undefined"],  # libtiff's own dependencies resolve through its rpath
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]
    import PIL  # libxnref_model.so uses Pillow's bundled libtiff; its dependencies sit beside it
    pillow_libs = os.path.join(os.path.dirname(os.path.dirname(PIL.__file__)), "pillow.libs")
    env = dict(os.environ, LD_LIBRARY_PATH=pillow_libs + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))
    run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120, env=env)
    import xenodon_b200 as xb
    try:
        have_gpu = xb.device_count() > 0
    except xb.XenodonError:  # no driver at all
        have_gpu = False
    if have_gpu:
        assert run.returncode == 0 and run.stdout.startswith("rendered 2304 rays"), run.stdout + run.stderr
    else:  # no CPU fallback: the library's own error comes back through the adapter's Error
        assert run.returncode == 3 and run.stdout.startswith("error: "), run.stdout + run.stderr
        assert "CUDA" in run.stdout or "device" in run.stdout, run.stdout
