#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY — builds oracle/_ref/libatmo_ref.so: the REFERENCE'S OWN shader sources compiled as C++.

The reference (Zylann/godot_atmosphere_shader) is GDShader text; Godot is not available here, but GDShader is a GLSL
dialect and GLSL is close enough to C++ that a purely SYNTACTIC rewrite plus a header of GLSL types/built-ins
(oracle/ref/glsl_compat.hpp) lets g++ compile the sources where they lie under /root/reference. Every entry shader
(shaders/planet_atmosphere_*.gdshader) becomes one translation unit containing its #defines, the include tree
(planet_atmosphere_main.gdshaderinc and what it includes) and its own vertex()/fragment() trampolines; the LUT bake shader
(shaders/optical_depth.gdshader) another. A small driver (ref_driver.inc / ref_bake_driver.inc) plays the engine: sets
uniforms and built-ins, runs vertex() once and fragment() per pixel.

The rewrite touches syntax only (no expression is reordered, no constant changed):
  * `#include "x"`               -> the file's text, recursively (its own #ifndef guards stay in charge)
  * `shader_type` / `render_mode` lines removed; `uniform T n : hints = v;` -> `static T n = v;`; `varying T n;` -> `static T n;`
  * `out T n` / `inout T n` parameters -> `T& n`; `in T n` -> `T n`; a trailing comma before `)` removed
  * float literals get an `f` suffix (GLSL literals are fp32; C++ ones would be double), and the keyword `float` is a macro
    for the library's real type: fp32 in libatmo_ref.so (THE reference), double in its twin libatmo_ref64.so (literals unsuffixed)
  * swizzles of 2-4 components (`.xyz`, `.rgb`, `.xz` ...) -> accessor calls (`.xyz()`); only reads occur in the sources
  * `discard;` -> sets the driver's flag and returns
  * `#define ATMOSPHERE_RAYMARCH_STEPS n` / `CLOUDS_MAX_RAYMARCH_STEPS n` -> a runtime variable initialised to n
    (so the BASELINE scale-ups 32 / 128 run through the same code); `#ifdef DOUBLE_PRECISION ... #endif` -> a runtime `if`
  * one C++-only name clash: cloud_funcs.gdshaderinc:39 declares a local `height_curve` initialised by a call to the
    function `height_curve` (legal GLSL scoping, ill-formed C++): the local is renamed
The generated C++ lives in a temporary directory during the build and is deleted; only libatmo_ref.so lands in
oracle/_ref/ (git-ignored; it travels to the GPU box, where /root/reference does not exist). Nothing of the reference is
copied into the repo tree.

usage: python oracle/ref/build_ref.py [--reference /root/reference] [--keep-sources /tmp/dir]
"""
import argparse
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE = os.path.dirname(HERE)
OUT = os.path.join(ORACLE, "_ref")
SHADERS = "addons/zylann.atmosphere/shaders"

ENTRY_SHADERS = [
    "planet_atmosphere_no_clouds", "planet_atmosphere_clouds", "planet_atmosphere_clouds_high",
    "planet_atmosphere_clouds_high_rm", "planet_atmosphere_v1_no_clouds", "planet_atmosphere_v1_clouds",
    "planet_atmosphere_v1_clouds_high",
]
STEP_MACROS = ("ATMOSPHERE_RAYMARCH_STEPS", "CLOUDS_MAX_RAYMARCH_STEPS")


def inline_includes(path, root):
    out = []
    for line in open(path, encoding="utf-8").read().splitlines():
        m = re.match(r'\s*#include\s+"([^"]+)"', line)
        if m:
            inc = os.path.normpath(os.path.join(os.path.dirname(path), m.group(1)))
            out.append(f"// >>> {os.path.relpath(inc, root)}")
            out.append(inline_includes(inc, root))
            out.append(f"// <<< {os.path.relpath(inc, root)}")
        else:
            out.append(line)
    return "\n".join(out)


def rewrite(text, real64=False):
    """GDShader -> C++ (syntax only; see the module docstring). real64: literals stay double (the fp64 twin)."""
    n_subs = {}

    def sub(pattern, repl, s, key, flags=0):
        s2, n = re.subn(pattern, repl, s, flags=flags)
        n_subs[key] = n_subs.get(key, 0) + n
        return s2

    text = sub(r"^\s*(shader_type|render_mode)\b[^\n]*$", "", text, "shader_type/render_mode", re.M)
    text = sub(r"\buniform\s+(\w+)\s+(\w+)\s*(?::[^=;]*)?(=[^;]*)?;", r"static \1 \2 \3;", text, "uniform")
    text = sub(r"\bvarying\s+(\w+)\s+(\w+)\s*;", r"static \1 \2;", text, "varying")
    text = sub(r"\b(?:inout|out)\s+(\w+)\s+(\w+)", r"\1& \2", text, "out/inout")
    text = sub(r"\bin\s+(?=(?:vec[234]|float|int|bool|mat[234])\b)", "", text, "in")
    text = sub(r",(\s*)\)", r"\1)", text, "trailing comma")
    if not real64:
        text = sub(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][+-]?\d+)?)(?![\w.])", r"\1f", text, "float literal")
    text = sub(r"\.([xyzw]{2,4}|[rgba]{2,4})\b(?!\s*\()", r".\1()", text, "swizzle")
    text = sub(r"\bdiscard\s*;", "{ ref_discarded = true; return; }", text, "discard")
    for macro in STEP_MACROS:
        text = sub(rf"^#define\s+{macro}\s+(\d+)\s*$",
                   rf"static int ref_{macro} = \1;  // the entry shader's #define, made a runtime variable\n#define {macro} ref_{macro}",
                   text, macro, re.M)
    text = sub(r"#ifdef DOUBLE_PRECISION(.*?)#endif", r"if (ref_double_precision) {\1}", text, "DOUBLE_PRECISION", re.S)
    # cloud_funcs.gdshaderinc:39 — local variable named like the function its initialiser calls
    if "float height_curve = max(height_curve(" in text:
        text = text.replace("float height_curve = max(height_curve(", "float height_curve_value = max(height_curve(")
        assert text.count("* height_curve;") == 1, "cloud_funcs.gdshaderinc changed: revisit the height_curve rename"
        text = text.replace("* height_curve;", "* height_curve_value;")
        n_subs["height_curve rename"] = 1
    return text, n_subs


def gen_entry(root, name, real64=False):
    path = os.path.join(root, SHADERS, name + ".gdshader")
    body, subs = rewrite(inline_includes(path, root), real64)
    src = f"""// GENERATED by oracle/ref/build_ref.py from {SHADERS}/{name}.gdshader — do not edit, do not commit.
#include "glsl_compat.hpp"
#define REF_ENTRY {name}
#define REF_ENTRY_STR "{name}"
namespace ref_{name} {{
GLSL_USING
#include "ref_builtins.inc"
// ------------------------------------------------ reference shader text (syntactic rewrite) ------------------------
#define float real   /* the shader's `float`: fp32 in libatmo_ref.so, double in the fp64 twin libatmo_ref64.so */
{body}
#undef float
// ------------------------------------------------ end of reference shader text --------------------------------------
#include "ref_driver.inc"
}}  // namespace
"""
    return src, subs


def gen_bake(root, real64=False):
    path = os.path.join(root, SHADERS, "optical_depth.gdshader")
    body, subs = rewrite(inline_includes(path, root), real64)
    src = f"""// GENERATED by oracle/ref/build_ref.py from {SHADERS}/optical_depth.gdshader — do not edit, do not commit.
#include "glsl_compat.hpp"
namespace ref_optical_depth {{
GLSL_USING
#include "ref_builtins.inc"
// ------------------------------------------------ reference shader text (syntactic rewrite) ------------------------
#define float real   /* the shader's `float`: fp32 in libatmo_ref.so, double in the fp64 twin libatmo_ref64.so */
{body}
#undef float
// ------------------------------------------------ end of reference shader text --------------------------------------
#include "ref_bake_driver.inc"
}}  // namespace
"""
    return src, subs


def build(reference="/root/reference", verbose=True, out_dir=None, keep_sources=None):
    """Generates the C++ sources in a temporary directory, compiles them and leaves ONLY libatmo_ref.so in `out_dir`
    (default oracle/_ref): the rewritten shader text is a derived copy of the reference and is not kept in the repo tree
    (`keep_sources=<dir>` keeps it there for debugging)."""
    import shutil
    import tempfile
    if not os.path.isdir(os.path.join(reference, SHADERS)):
        raise FileNotFoundError(f"{reference}/{SHADERS} not found: the reference tree is needed to build oracle/_ref")
    out_root = out_dir or OUT
    os.makedirs(out_root, exist_ok=True)
    work = keep_sources or tempfile.mkdtemp(prefix="b200atmo_ref_")
    os.makedirs(work, exist_ok=True)
    try:
        libs = []
        for real64 in (False, True):
            tag = "64" if real64 else ""
            sources = []
            for name in ENTRY_SHADERS:
                src, subs = gen_entry(reference, name, real64)
                p = os.path.join(work, f"gen{tag}_{name}.cpp")
                open(p, "w").write(src)
                sources.append(p)
                if verbose and not real64:
                    print(f"  {name}: " + ", ".join(f"{k} x{v}" for k, v in subs.items() if v))
            src, subs = gen_bake(reference, real64)
            p = os.path.join(work, f"gen{tag}_optical_depth.cpp")
            open(p, "w").write(src)
            sources.append(p)
            sources.append(os.path.join(HERE, "ref_dispatch.cpp"))
            lib = os.path.join(out_root, f"libatmo_ref{tag}.so")
            cxx = os.environ.get("CXX", "g++")
            flags = ["-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-pthread", "-Wall", "-Wno-unused-variable",
                     "-Wno-unused-function", "-Wno-unused-but-set-variable", "-I" + HERE] + (["-DGLSL_REAL=double"] if real64 else [])
            objs, procs = [], []
            for s in sources:
                o = os.path.join(work, f"o{tag}_" + os.path.basename(s)[:-4] + ".o")
                objs.append(o)
                procs.append((s, subprocess.Popen([cxx] + flags + ["-c", s, "-o", o], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
            failed = False
            for s, pr in procs:
                out, _ = pr.communicate()
                if pr.returncode != 0:
                    failed = True
                    sys.stderr.write(f"--- {s}\n{out[-6000:]}\n")
            if failed:
                raise RuntimeError("oracle/_ref: compiling the rewritten reference shaders failed")
            subprocess.check_call([cxx, "-shared", "-pthread", "-o", lib] + objs)
            if verbose:
                print(f"built {lib}")
            libs.append(lib)
        return libs[0]
    finally:
        if not keep_sources:
            shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default=os.environ.get("B200ATMO_REFERENCE", "/root/reference"))
    ap.add_argument("--keep-sources", default=None, help="directory to keep the generated C++ in (debugging; keep it outside the repo)")
    a = ap.parse_args()
    build(a.reference, keep_sources=a.keep_sources)
