"""CPU tests of the boundary: libsdf_b200.so loads, exports exactly what include/sdf_b200.h declares, and refuses to
run without a GPU (no CPU fallback).  No compute calls."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "sdf_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sdf_b200_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_what_capi_binds():
    from sdf_tools_b200 import _capi
    assert sorted(n for n, _, _ in _capi.SYMBOLS) == declared_symbols()


def test_library_exports_every_declared_symbol():
    from sdf_tools_b200 import _capi
    lib = _capi.load()
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in sdf_b200.h but not exported"
    out = subprocess.run(["nm", "-D", "--defined-only", _capi.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r"\bT (sdf_b200_[a-z0-9_]+)", out)))
    assert exported == declared_symbols(), "exported C symbols and header differ"
    assert b"sm_100a" in lib.sdf_b200_version()


def test_no_cpu_fallback():
    from sdf_tools_b200 import _capi
    lib = _capi.load()
    if lib.sdf_b200_device_count() > 0:
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    rc = lib.sdf_b200_create(C.byref(h), 0)
    assert rc == _capi.ERR_CUDA and not h.value
    assert b"no CUDA device" in lib.sdf_b200_last_error(None)
    from sdf_tools_b200 import Context, SdfB200Error
    with pytest.raises(SdfB200Error):
        Context(0)


def test_library_has_sm100a_code_and_no_torch_dependency():
    from sdf_tools_b200 import _capi
    ldd = subprocess.run(["ldd", _capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in ldd and "libcudart" not in ldd          # static cudart, plain C ABI
    sass = subprocess.run(["cuobjdump", "-lelf", _capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass


def test_cpp_face_compiles_and_fails_loudly_without_gpu(tmp_path):
    """The C++ headers with the reference's signatures compile against the C ABI with plain g++."""
    src = tmp_path / "t.cpp"
    src.write_text(r'''
#include <sdf_tools/collision_map.hpp>
#include <cstdio>
int main() {
    using namespace sdf_tools;
    CollisionMapGrid g(Eigen::Isometry3d::Identity(), "world", 0.05, (int64_t)20, (int64_t)40, (int64_t)1, COLLISION_CELL(-10000.f));
    g.SetValue((int64_t)3, (int64_t)1, (int64_t)0, COLLISION_CELL(1.0f));
    const std::function<bool(const VoxelGrid::GRID_INDEX&)> fn = [&](const VoxelGrid::GRID_INDEX& i) { return g.GetImmutable(i).first.occupancy > 0.5f; };
    // non-uniform cells must throw std::invalid_argument before anything touches the GPU (sdf_generation.hpp:277-281)
    VoxelGrid::VoxelGrid<int> bad(Eigen::Isometry3d::Identity(), 1.0, 2.0, 1.0, 2, 2, 2, 0, 0);
    try { sdf_generation::ExtractSignedDistanceField(bad, fn, 0.f, "f", false); return 2; }
    catch (const std::invalid_argument& e) { std::printf("invalid_argument: %s\n", e.what()); }
    try {
        auto r = g.ExtractSignedDistanceField(-10000.f, false, false);
        std::printf("value %.6f max %.6f min %.6f\n", r.first.GetImmutable((int64_t)3, (int64_t)1, (int64_t)0).first, r.second.first, r.second.second);
    } catch (const std::runtime_error& e) { std::printf("runtime_error: %s\n", e.what()); }
    return 0;
}
''')
    exe = tmp_path / "t"
    libdir = os.path.join(ROOT, "sdf_tools_b200", "lib")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-Wshadow", "-I", os.path.join(ROOT, "include"),
                           str(src), "-o", str(exe), "-L", libdir, "-lsdf_b200", f"-Wl,-rpath,{libdir}"])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "invalid_argument: Grid must have uniform resolution" in out.stdout
    from sdf_tools_b200 import _capi
    if _capi.load().sdf_b200_device_count() == 0:
        assert "runtime_error: sdf_b200" in out.stdout and "no CUDA device" in out.stdout
    else:
        assert "value -0.050000" in out.stdout
