"""The C-ABI shared library loads, exports every symbol include/b200atmo.h declares, and the ctypes mirrors agree
with the compiled struct sizes. No compute calls (no GPU needed)."""
import ctypes as C
import os
import re

import pytest

from godot_atmosphere_shader_b200 import abi, context

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_header_symbols():
    lib = context.lib()
    header = open(os.path.join(ROOT, "include", "b200atmo.h")).read()
    declared = set(re.findall(r"\b(b200atmo_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/b200atmo.h but not exported"
    assert declared == set(context.EXPORTS)
    assert lib.b200atmo_version() == 2


def test_struct_sizes_match():
    lib = context.lib()
    assert lib.b200atmo_sizeof_params() == C.sizeof(abi.B200AtmoParams)
    assert lib.b200atmo_sizeof_frame() == C.sizeof(abi.B200AtmoFrame)
    assert lib.b200atmo_sizeof_camera() == C.sizeof(abi.B200AtmoCamera)
    assert lib.b200atmo_sizeof_peer_targets() == C.sizeof(abi.B200AtmoPeerTargets)


def test_default_params_match_shader_defaults():
    lib = context.lib()
    p = abi.B200AtmoParams()
    lib.b200atmo_default_params(C.byref(p))
    assert bytes(p) == bytes(abi.default_params())
    # shader-source defaults
    assert (p.planet_radius, p.atmosphere_height, p.density) == (1.0, pytest.approx(0.1), pytest.approx(0.2))
    assert p.scattering_strength == 20.0 and list(p.scattering_wavelengths) == [700.0, 530.0, 440.0]
    assert list(p.atmosphere_ambient_color) == pytest.approx([0.0, 0.0, 0.002])
    assert (p.cloud_density_scale, p.cloud_bottom, p.cloud_top, p.cloud_blend) == (50.0, pytest.approx(0.2), 0.5, 0.5)
    assert (p.cloud_shape_factor, p.cloud_shape_scale, p.cloud_shape_invert) == (pytest.approx(0.8), 1.0, 0.0)


def test_no_cpu_fallback_without_a_device():
    """Without a CUDA device the product refuses to create a context (it never falls back to the oracle or the CPU)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(context.B200AtmoError) as e:
        context.AtmosphereContext(0)
    assert e.value.code == abi.E_CUDA and "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    """Neither the package, nor the GDExtension wrapper, nor the C example reference the oracle, the compiled reference
    shaders or the host simulation; the shipped libraries link neither."""
    import subprocess
    for top in ("godot_atmosphere_shader_b200", "gdextension", "examples", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".c", ".cpp", ".sh", ".inc")):
                    text = open(os.path.join(dirpath, f)).read().replace("tests/hostsim", "")
                    for word in ("pyoracle", "liboracle", "hostsim", "pyref", "libatmo_ref", "atmo_oracle", "from oracle", "import oracle"):
                        assert word not in text, f"{top}/{f} references test infrastructure ({word})"
    for so in ("libb200atmo.so", "libb200atmo_node.so"):
        needed = subprocess.run(["readelf", "-d", os.path.join(ROOT, "godot_atmosphere_shader_b200", so)], capture_output=True, text=True).stdout
        assert "oracle" not in needed and "atmo_ref" not in needed


@pytest.mark.gpu
def test_c_client_example_builds_and_runs(tmp_path):
    """examples/minimal.c: a plain C program against include/b200atmo.h (what a GDExtension would link)."""
    import subprocess
    exe = str(tmp_path / "minimal")
    lib_dir = os.path.join(ROOT, "godot_atmosphere_shader_b200")
    subprocess.check_call(["gcc", "-Wall", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "minimal.c"),
                           "-L" + lib_dir, "-lb200atmo", "-Wl,-rpath," + lib_dir, "-lm", "-o", exe])
    out = subprocess.check_output([exe], text=True)
    vals = [float(x) for x in out.split("=")[1].split(";")[0].split()]
    assert len(vals) == 4 and all(0.0 <= v <= 1.0 for v in vals) and vals[3] > 0.1   # looking at the planet: alpha well above 0
    assert "kernels launched" in out
