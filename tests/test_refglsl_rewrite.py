"""Unit tests of the syntactic GLSL -> C++ rewrite (oracle/refglsl/build_ref.py) on synthetic snippets: the rules that the
reference-pinning harness relies on must not silently change meaning."""
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "oracle", "refglsl"))
import build_ref  # noqa: E402


def rw(src, repl=()):
    return build_ref.rewrite(src.splitlines(), list(repl), None)


def test_float_literals_get_f_suffix_and_integers_do_not():
    out = rw("float a = 1. / 0.;\nfloat b = .5 + 3.14 * 2 + 1e-3 + 0x3FFu + 10u + 7;\nint c = 1 << 24;\n")
    assert "1.f / 0.f" in out and ".5f" in out and "3.14f" in out and "1e-3f" in out
    assert "0x3FFu" in out and "10u" in out and " 7;" in out and "1 << 24" in out and "24f" not in out


def test_array_declarators_and_uninitialised_locals():
    out = rw("void main()\n{\n    ivec2[4] offsets = { {0,0}, {0,1} };\n    float[4] depths;\n    float maxDepth;\n    float a, b;\n}\n")
    assert "ivec2 offsets[4] =" in out and "float depths[4]" in out
    assert "float maxDepth{};" in out and "a{}, b{}" in out and "void shader_main()" in out


def test_interface_declarations_become_variables():
    src = """#version 430
#extension GL_ARB_shading_language_include : require
layout (local_size_x = 8, local_size_y = 8) in;
layout (rgba32f, binding = 1) restrict writeonly uniform image2D img_output;
layout (binding = 0) uniform usampler2D tex;
layout (std140, binding = 1) buffer block_
{
    vec4 items[];
};
layout (shared, binding = 0) buffer counters_
{
    uint[1024] counter;
};
uniform int level = 3;
uniform float zFar;
float scale = 19 / log2(zFar);
shared vec4[16] cache;
void f(out vec3 a, inout float b, const in vec2 c, in float d) {}
"""
    out = rw(src)
    assert "#version" not in out and "#extension" not in out
    assert "static const wgsize<8, 8, 1> gl_WorkGroupSize;" in out
    assert "image2D img_output;" in out and "usampler2D tex;" in out and "layout" not in out
    assert "vec4* items;" in out and "uint* counter;" in out
    assert "int level = 3;" in out                       # uniform default: static initialisation
    assert "float scale;" in out and "scale = 19 / log2(zFar);" in out and "shader_globals_init" in out
    assert "static vec4 cache[16];" in out
    assert "void f(vec3& a, float& b, const vec2 c, float d)" in out


def test_first_local_size_x_wins_when_repeated():
    assert "wgsize<64, 1, 1>" in rw("layout (local_size_x = 64, local_size_y = 1, local_size_x = 1) in;\n")


def test_host_replacements_must_match():
    import pytest
    assert "#undef X" in rw("#define X\n", [("#define X", "#undef X")])
    with pytest.raises(AssertionError):
        rw("#define Y\n", [("#define X", "#undef X")])


def test_rewritten_snippet_compiles_and_runs_against_the_runtime():
    """a small synthetic compute shader (swizzles, implicit int->float conversion, image store, barrier + shared atomics)
    through the same path as the reference units"""
    shader = """
layout (local_size_x = 4) in;
layout (rgba32f, binding = 0) uniform image2D img;
uniform float gain = 2.0;
shared uint hits;
void main()
{
    if (gl_LocalInvocationID.x == 0) hits = 0;
    barrier();
    atomicAdd(hits, 1);
    barrier();
    ivec2 p = ivec2(gl_GlobalInvocationID.xy);
    vec4 v = vec4(p + 0.5, 1, hits);
    v.xy *= gain;
    v.z = min(v.z, 15) / 4;
    imageStore(img, p, v);
}
"""
    body = rw(shader)
    drv = """
extern "C" int run(float* out, int n) {
    img.data = out; img.w = n; img.h = 1; img.depth = 1; img.fmt = F_RGBA32F;
    shader_globals_init();
    dispatch_phased(shader_main, gl_WorkGroupSize, n / 4, 1, 1);
    return 0;
}
"""
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.cpp")
        with open(src, "w") as f:
            f.write('#include "%s"\nnamespace glsl_rt { namespace u_t {\n%s\n%s\n}}\n' % (os.path.join(ROOT, "oracle", "refglsl", "glsl_rt.h"), body, drv))
        lib = os.path.join(d, "t.so")
        subprocess.check_call([build_ref.CXX, "-std=c++20", "-O1", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-o", lib, src])
        import ctypes as C
        import numpy as np
        out = np.zeros((8, 4), np.float32)
        C.CDLL(lib).run(out.ctypes.data_as(C.c_void_p), 8)
        for x in range(8):
            assert out[x].tolist() == [(x + 0.5) * 2.0, 0.5 * 2.0, 0.25, 4.0]
