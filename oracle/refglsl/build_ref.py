#!/usr/bin/env python3
"""Build oracle/_ref/libref_glsl.so: the reference's own GLSL sources (read from /root/reference/data/shaders where
they lie, never copied into the repo) rewritten just enough to be C++20, compiled with g++ against glsl_rt.h and
wrapped by the small drivers in drivers/*.inc. TEST INFRASTRUCTURE: used only to pin oracle/oracle.cpp.

The rewrite is purely syntactic (what a GLSL front end does before code generation):
  * `#include </data/shaders/...>` is expanded; `#version` / `#extension` lines are dropped;
  * the textual replacements the reference's host code applies before compiling a shader (`Shader::globalReplace`,
    e.g. `#define LEVEL_ZERO` -> `#undef LEVEL_ZERO`, ImperfectShadowmap.cpp:52-60) are applied per variant;
  * floating-point literals get an `f` suffix (GLSL literals are 32-bit; C++'s would be double);
  * `T[4] name` array declarators become `T name[4]`; `in` / `out` / `inout` parameter qualifiers become value /
    reference parameters; `void main()` becomes `void shader_main()`;
  * interface declarations (`layout(...) uniform image2D x;`, `uniform T x;`, SSBO / UBO blocks, `shared`) become
    namespace-scope variables the driver fills in;
  * locals declared without an initialiser are value-initialised (reading them is undefined in GLSL; NVIDIA's
    compiler yields 0, which is also what the oracle assumes).
The library goes to oracle/_ref/ (git-ignored); the generated sources are deleted after the build unless --keep is given."""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE = os.path.dirname(HERE)
REF_ROOT = os.environ.get("MLGI_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(ORACLE, "_ref")
GEN = os.path.join(OUT, "gen")
CXX = "/usr/bin/g++"

# (unit name, shader path, host-side textual replacements, driver include)
UNITS = [
    ("pull0", "ism/pull.comp", [], "pullpush.inc"),
    ("pull", "ism/pull.comp", [("#define LEVEL_ZERO", "#undef LEVEL_ZERO")], "pullpush.inc"),
    ("push0", "ism/push.comp", [], "pullpush.inc"),
    ("push", "ism/push.comp", [("#define LEVEL_ZERO", "#undef LEVEL_ZERO")], "pullpush.inc"),
    ("vplproc", "gi/vpl_processor.comp", [], "vplproc.inc"),
    ("gather", "gi/final_gathering.comp", [], "gather.inc"),
    ("gather_noshadow", "gi/final_gathering.comp", [("#define ENABLE_SHADOWING true", "#define ENABLE_SHADOWING false")], "gather.inc"),
    ("gather_scaled", "gi/final_gathering.comp", [("#define SCALE_ISMS false", "#define SCALE_ISMS true")], "gather.inc"),
    ("clustering", "clustered_shading/clustering.comp", [], "clustering.inc"),
    ("lightlists", "clustered_shading/light_lists.comp", [], "lightlists.inc"),
    ("splat", "ism/ism.comp", [], "splat.inc"),
    ("tesc", "ism/ism.tesc", [], "points.inc"),
    ("tese", "ism/ism.tese", [], "points.inc"),
    ("geom", "ism/ism.geom", [], "points.inc"),
    ("blur_h", "gi/gi_blur.frag", [("#define DIRECTION ivec2(0,0)", "#define DIRECTION ivec2(1,0)")], "blur.inc"),
    ("blur_v", "gi/gi_blur.frag", [("#define DIRECTION ivec2(0,0)", "#define DIRECTION ivec2(0,1)")], "blur.inc"),
    ("composite", "deferredshading.frag", [], "composite.inc"),
]

SCALARS = r"(?:float|int|uint|bool|vec[234]|ivec[234]|uvec[234]|mat[34])"


def expand_includes(path, seen=None):
    seen = seen if seen is not None else set()
    out = []
    for line in open(path, encoding="utf-8", errors="replace").read().splitlines():
        m = re.match(r"\s*#\s*include\s*<(/data/shaders/[^>]+)>", line)
        if m:
            inc = os.path.join(REF_ROOT, m.group(1).lstrip("/"))
            if inc not in seen:
                seen.add(inc)
                out.extend(expand_includes(inc, seen))
            continue
        out.append(line)
    return out


def strip_comments(src):
    src = re.sub(r"/\*.*?\*/", lambda m: "\n" * m.group(0).count("\n"), src, flags=re.S)
    return re.sub(r"//[^\n]*", "", src)


FLOAT_LIT = re.compile(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][+-]?\d+)?|\d+[eE][+-]?\d+)(?![\w.])")


def rewrite(lines, replacements, total_vpls):
    src = "\n".join(lines)
    for a, b in replacements:
        assert a in src, "host replacement %r not found" % a
        src = src.replace(a, b)
    src = strip_comments(src)
    src = re.sub(r"^\s*#\s*(version|extension)[^\n]*$", "", src, flags=re.M)
    src = FLOAT_LIT.sub(lambda m: m.group(1) + "f", src)
    if total_vpls is not None:          # the reference hard-codes 1024 VPLs (final_gathering.comp:18 etc.)
        src = re.sub(r"(const\s+int\s+totalVplCount\s*=\s*)1024", r"\g<1>%d" % total_vpls, src)
    decls = []     # namespace-scope declarations produced from interface blocks

    # work group size
    def local_size(m):
        d = {}
        for k, v in re.findall(r"local_size_(\w)\s*=\s*(\d+)", m.group(0)):
            d.setdefault(k, v)          # light_lists.comp:6 names local_size_x twice (64, then 1): the intended 64
        return "static const wgsize<%s, %s, %s> gl_WorkGroupSize;" % (d.get("x", "1"), d.get("y", "1"), d.get("z", "1"))
    src = re.sub(r"layout\s*\([^)]*local_size_[^)]*\)\s*in\s*;", local_size, src)

    # tessellation / geometry stage interface: layout-only declarations, unsized per-vertex arrays, `flat`
    src = re.sub(r"^\s*layout\s*\([^)]*\)\s*(?:in|out)\s*;", "", src, flags=re.M)
    src = re.sub(r"^\s*(?:flat\s+)?(?:in|out)\s+(" + SCALARS + r")\s*\[\s*\]\s+(\w+)\s*;", r"\1 \2[3];", src, flags=re.M)
    src = re.sub(r"^\s*flat\s+(in|out)\s+", r"\1 ", src, flags=re.M)
    # buffer / uniform blocks: `layout(...) buffer name_ { T member[...]; };`  ->  `T* member;`
    def block(m):
        body = m.group("body")
        out = []
        body = re.sub(r"(\w+)\s*\[([^\]]*)\]\s+(\w+)\s*;", r"\1 \3[\2];", body)      # `T[n] name;` -> `T name[n];`
        for t, name, dim in re.findall(r"(\w+)\s+(\w+)\s*\[([^\]]*)\]\s*;", body):
            out.append("%s* %s;" % (t, name))
        for t, name in re.findall(r"(\w+)\s+(\w+)\s*;", body):
            out.append("%s %s;" % (t, name))
        return "\n".join(out)
    src = re.sub(r"layout\s*\([^)]*\)\s*(?:restrict\s+|readonly\s+|writeonly\s+|coherent\s+)*(?:buffer|uniform)\s+\w+\s*\{(?P<body>[^}]*)\}\s*\w*\s*;",
                 block, src)
    # opaque uniforms with a layout qualifier
    src = re.sub(r"layout\s*\([^)]*\)\s*((?:restrict\s+|readonly\s+|writeonly\s+|coherent\s+)*)uniform\s+", "uniform ", src)
    src = re.sub(r"\buniform\s+((?:restrict\s+|readonly\s+|writeonly\s+|coherent\s+)+)", "uniform ", src)
    src = re.sub(r"^\s*uniform\s+", "UNIFORM_DECL ", src, flags=re.M)
    # namespace-scope variables with a non-constant initialiser (`float scaleFactor = ... / log2(zFar);`): GLSL evaluates
    # them per invocation after the uniforms are set; here they are assigned by shader_globals_init(), which the driver
    # calls once the uniforms are in place
    inits = []
    def global_init(m):
        inits.append("%s = %s;" % (m.group(2), m.group(3)))
        return "%s %s;" % (m.group(1), m.group(2))
    src = re.sub(r"^(" + SCALARS + r")\s+(\w+)\s*=\s*([^;]+);", global_init, src, flags=re.M)
    src = src.replace("UNIFORM_DECL ", "")
    src += "\nstatic void shader_globals_init() { %s }\n" % " ".join(inits)
    src = re.sub(r"\bshared\s+", "static ", src)
    # fragment-shader interface
    src = re.sub(r"^\s*(?:layout\s*\([^)]*\)\s*)?in\s+(" + SCALARS + r")\s+(\w+)\s*;", r"\1 \2;", src, flags=re.M)
    src = re.sub(r"^\s*(?:layout\s*\([^)]*\)\s*)?out\s+(" + SCALARS + r")\s+(\w+)\s*;", r"\1 \2;", src, flags=re.M)
    # parameter qualifiers
    src = re.sub(r"\bconst\s+in\s+", "const ", src)
    src = re.sub(r"([(,]\s*)in\s+(?=" + SCALARS + r"|sampler)", r"\1", src)
    src = re.sub(r"([(,]\s*)(?:out|inout)\s+(" + SCALARS + r")\s+(\w+)", r"\1\2& \3", src)
    # array declarators: `T[4] name` -> `T name[4]`
    src = re.sub(r"\b(" + SCALARS + r")\s*\[\s*([^\]\n]+?)\s*\]\s+(\w+)", r"\1 \3[\2]", src)
    # uninitialised scalar / vector locals (indented declarations only: namespace-scope ones are zero anyway)
    src = re.sub(r"^([ \t]+(?:" + SCALARS + r")\s+\w+)\s*;", r"\1{};", src, flags=re.M)
    def multi(m):                       # `float a, b, c;`
        names = [n.strip() for n in m.group(3).split(",")]
        return m.group(1) + m.group(2) + " " + ", ".join(n + "{}" for n in names) + ";"
    src = re.sub(r"^([ \t]+)(float|int|uint)\s+(\w+(?:\s*,\s*\w+)+)\s*;", multi, src, flags=re.M)
    src = re.sub(r"\bvoid\s+main\s*\(\s*(?:void)?\s*\)", "void shader_main()", src)
    return src


def main():
    """usage: build_ref.py [--keep] [unit ...]   (--keep leaves the generated C++ under oracle/_ref/gen for inspection)"""
    import shutil
    # --vpls N: the reference hard-codes `const int totalVplCount = 1024;` in every shader; to exercise other VPL counts the
    # literal is substituted (the same kind of textual replacement its host code applies to the #defines) and the library
    # is written to libref_glsl_v<N>.so. light_lists.comp additionally hard-codes 64-entry sub-lists and is only meaningful
    # at 1024.
    total_vpls = None
    argv = sys.argv[1:]
    if "--vpls" in argv:
        i = argv.index("--vpls"); total_vpls = int(argv[i + 1]); del argv[i:i + 2]
    sys.argv[1:] = argv
    keep = "--keep" in sys.argv[1:]
    shutil.rmtree(GEN, ignore_errors=True)
    os.makedirs(GEN, exist_ok=True)
    units = UNITS
    only = [a for a in sys.argv[1:] if not a.startswith("-")]
    if only:
        units = [u for u in UNITS if u[0] in only]
    srcs = []
    try:
        for name, rel, repl, driver in units:
            drv = os.path.join(HERE, "drivers", driver)
            if not os.path.exists(drv):
                continue
            body = rewrite(expand_includes(os.path.join(REF_ROOT, "data/shaders", rel)), repl, total_vpls)
            path = os.path.join(GEN, name + ".cpp")
            with open(path, "w") as f:
                f.write("// GENERATED by oracle/refglsl/build_ref.py from %s -- never commit\n" % os.path.join(REF_ROOT, "data/shaders", rel))
                f.write('#include "../../refglsl/glsl_rt.h"\n#define REF_UNIT %s\n#define REF_UNIT_%s 1\nnamespace glsl_rt { namespace u_%s {\n' % (name, name, name))
                f.write(body)
                f.write('\n#include "../../refglsl/drivers/%s"\n}}\n' % driver)
            srcs.append(path)
        lib = os.path.join(OUT, "libref_glsl.so" if total_vpls is None else "libref_glsl_v%d.so" % total_vpls)
        cmd = [CXX, "-std=c++20", "-O2", "-march=x86-64-v3", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-w", "-o", lib] + srcs
        print(" ".join(cmd))
        subprocess.check_call(cmd)
        print("built", lib)
    finally:
        if not keep:
            shutil.rmtree(GEN, ignore_errors=True)      # the rewritten reference text does not stay in the tree


if __name__ == "__main__":
    main()
