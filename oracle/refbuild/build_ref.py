#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY.  Builds oracle/_ref/libref.so from the reference's own shader sources.

The reference (f1shel/Asuna) runs its per-pixel program as GLSL on the Vulkan ray-tracing pipeline and
cannot be built here (no Vulkan SDK / glslang / RT device).  Its shader text, however, is
C-like, and the reference tree vendors GLM (src/ext/nvpro_core/third_party/tinygltf/examples/common/glm),
whose types and functions carry GLSL semantics.  This recipe

  1. reads src/shared/*.h, src/shaders/utils/*.glsl, all twelve src/shaders/bxdf/*.rchit, the two miss
     shaders and the ray-generation shader FROM /root/reference AT BUILD TIME (nothing is copied into this
     repository; the generated unit lands in oracle/_ref/, which is git-ignored),
  2. rewrites only what C++ cannot parse: `in/out/inout` parameter qualifiers, unsuffixed floating
     literals (GLSL literals are fp32), rvalue swizzles, `layout(...)` declarations and #include lines,
  3. wraps each shader stage in its own namespace (they all define eval / pdf / sampleBsdf / main),
  4. compiles it against GLM with oracle.cpp (-DASUNA_REF_SHADERS) supplying the scene store, the BVH
     and the C ABI, so that libref.so exposes the same oracle_* entry points as liboracle.so while every
     line of shading, sampling, RNG, camera and accumulation arithmetic is the reference's own.

tests/test_ref_pins.py holds liboracle.so (the hand restatement) to libref.so.  The product never loads either.
Usage: python oracle/refbuild/build_ref.py [--reference /root/reference] [--keep-going]
"""
import argparse
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE = os.path.dirname(HERE)
OUT = os.path.join(ORACLE, "_ref")

SHARED = ["binding.h", "light.h", "material.h", "pushconstant.h", "instance.h", "sun_and_sky.h", "vertex.h", "camera.h"]
UTILS = ["structs.glsl", "math.glsl", "sun_and_sky.glsl", "sample_light.glsl", "tonemapping.glsl"]
RCHIT = ["brdf_lambertian", "brdf_kang18", "brdf_emissive", "brdf_pbr_metalness_roughness", "brdf_plastic",
         "brdf_rough_plastic", "brdf_conductor", "brdf_rough_conductor", "brdf_mirror", "brdf_disney",
         "bsdf_dielectric", "brdf_phong"]

FLOAT_LIT = re.compile(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][+-]?\d+)?|\d+[eE][+-]?\d+)(?![\w.])")
QUAL_REF = re.compile(r"([(,]\s*)(?:inout|out)\s+((?:const\s+)?\w+)\s+(?=\w)")
QUAL_IN = re.compile(r"([(,]\s*)in\s+(?=(?:const\s+)?\w+\s+\w)")
SWIZZLE = re.compile(r"\.(xyz|xy|rgba|rgb|rg)\b(?!\s*\()")


def strip_comments(src):
    """Drop /* */ and // comments (they would otherwise be hit by the rewrites below)."""
    src = re.sub(r"/\*.*?\*/", lambda m: "\n" * m.group(0).count("\n"), src, flags=re.S)
    return re.sub(r"//[^\n]*", "", src)


def rewrite(src, name):
    src = strip_comments(src)
    out = []
    for line in src.split("\n"):
        s = line.strip()
        if s.startswith("#version") or s.startswith("#extension") or s.startswith("#include"):
            continue
        if s.startswith("layout") or s.startswith("hitAttributeEXT"):
            if not s.endswith(";"):
                raise SystemExit(f"{name}: multi-line layout declaration not handled: {s}")
            continue
        out.append(line)
    src = "\n".join(out)
    src = src.replace("__cplusplus", "REFGLSL_NEVER_DEFINED")  # take the GLSL branch of the shared headers
    src = FLOAT_LIT.sub(lambda m: m.group(1) + "f", src)
    src = QUAL_REF.sub(lambda m: f"{m.group(1)}{m.group(2)}& ", src)
    src = QUAL_IN.sub(lambda m: m.group(1), src)
    src = SWIZZLE.sub(lambda m: "->*SW_" + m.group(1), src)
    # GLSL evaluates function arguments left to right; C++ leaves the order open (g++ goes right to left)
    # except inside braces.  The only calls with two side-effecting arguments are rand2 / rand3
    # (utils/math.glsl:43-44); everything else is checked to have at most one RNG draw per statement.
    src = re.sub(r"\bvec([23])\((rand\((\w+)\)(?:, rand\(\3\))+)\)", r"vec\1{\2}", src)
    for stmt in src.split(";"):
        body = stmt.split("{")[-1]
        if len(re.findall(r"\b(?:rand[23]?|pcg)\s*\(", body)) > 1 and "vec2{" not in stmt and "vec3{" not in stmt:
            raise SystemExit(f"{name}: two RNG draws in one statement, evaluation order would be unspecified: {body.strip()[:120]}")
    return f"// ======== {name} ========\n{src}\n"


def generate(ref):
    sh = os.path.join(ref, "src", "shaders")
    rd = lambda *p: open(os.path.join(*p)).read()
    inc = lambda f: open(os.path.join(HERE, f)).read()
    parts = ['#include "ref_bridge.h"\n#include "shim_pre.h"\n']  # shim_pre.h opens namespace refglsl
    for f in SHARED:
        parts.append(rewrite(rd(ref, "src", "shared", f), "src/shared/" + f))
    for f in UTILS[:4]:
        parts.append(rewrite(rd(sh, "utils", f), "src/shaders/utils/" + f))
    parts.append("namespace tonemapping {\n" + rewrite(rd(sh, "utils", UTILS[4]), "src/shaders/utils/" + UTILS[4]) + "}\n")
    parts.append(inc("shim_bindings.inc"))
    parts.append(rewrite(rd(sh, "utils", "rchit_layouts.glsl"), "src/shaders/utils/rchit_layouts.glsl"))
    for m in RCHIT:
        f = f"raytrace.{m}.rchit"
        parts.append(f"namespace rchit_{m} {{\n" + rewrite(rd(sh, "bxdf", f), "src/shaders/bxdf/" + f) + "}\n")
    parts.append("namespace rmiss_default {\n" + rewrite(rd(sh, "raytrace.default.rmiss"), "src/shaders/raytrace.default.rmiss") + "}\n")
    parts.append("namespace rmiss_shadow {\n" + rewrite(rd(sh, "raytrace.shadow.rmiss"), "src/shaders/raytrace.shadow.rmiss") + "}\n")
    parts.append(inc("shim_trace.inc"))
    parts.append("namespace rgen {\n" + rewrite(rd(sh, "raytrace.projective.rgen"), "src/shaders/raytrace.projective.rgen") + "}\n")
    # post-process stage (LDR output): post.idle.frag with its own copies of tonemapping.glsl (it defines
    # TONEMAP_UNCHARTED before the include) and random.glsl
    post = rd(sh, "post.idle.frag")
    post = re.sub(r"(\w+)\.rgb\s*=([^;]*);", r"assign_rgb(\1, \2);", strip_comments(post))  # lvalue swizzles
    post = re.sub(r"=\s*float\[\d+\]\(([^;]*)\);", r"= {\1};", post)                       # GLSL array constructor
    tm = rd(sh, "utils", "tonemapping.glsl").replace("TONEMAPPING_GLSL", "TONEMAPPING_GLSL_POST")
    parts.append("namespace post_idle {\n" + inc("shim_post.inc") + "#define TONEMAP_UNCHARTED\n" +
                 rewrite(tm, "src/shaders/utils/tonemapping.glsl (post)") +
                 rewrite(rd(sh, "utils", "random.glsl"), "src/shaders/utils/random.glsl") +
                 rewrite(post.replace("#define TONEMAP_UNCHARTED", ""), "src/shaders/post.idle.frag") + "}\n")
    parts.append(inc("shim_export.inc"))
    parts.append(inc("shim_hooks.inc") if os.path.exists(os.path.join(HERE, "shim_hooks.inc")) else "")
    parts.append("}  // namespace refglsl\n")
    return "".join(parts)


def extract_block(src, start_pat, name):
    """Text from the first match of start_pat to the brace (or brace + ';') that closes its first '{'."""
    m = re.search(start_pat, src)
    if not m:
        raise SystemExit(f"{name}: pattern {start_pat!r} not found in the reference")
    i = src.index("{", m.start())
    depth, j = 0, i
    while True:
        depth += {"{": 1, "}": -1}.get(src[j], 0)
        j += 1
        if depth == 0:
            break
    if src[j:j + 1] == ";":
        j += 1
    return src[m.start():j] + "\n"


def generate_host(ref):
    """Host-side reference arithmetic the scene loaders must reproduce, compiled from the reference's C++:
    computeDiffuseFresnel + the complex-IOR table (src/loader/material.cpp), the camera matrices
    (src/core/camera.cpp with the vendored header-only nvmath) and the env-map importance tables
    (the body of EnvMap::EnvMap, src/core/texture.cpp)."""
    rd = lambda *p: open(os.path.join(ref, "src", *p)).read()
    mat, cam, tex = rd("loader", "material.cpp"), rd("core", "camera.cpp"), rd("core", "texture.cpp")
    ctor = extract_block(tex, r"EnvMap::EnvMap\(const std::string& envmapPath\)", "core/texture.cpp")
    body = ctor[ctor.index("{") + 1:ctor.rindex("}")]
    keep = []
    for line in body.split("\n"):
        st = line.strip()
        if st.startswith("m_data = readImage") or st.startswith("int32_t width, height") or st.startswith("m_shape ="):
            continue
        if st.startswith("m_marginal = malloc") or st.startswith("m_conditional = malloc"):
            continue
        keep.append(line)
    parts = [open(os.path.join(HERE, "shim_host_pre.h")).read(),
             "// ======== src/loader/material.cpp ========\n",
             extract_block(mat, r"static float dielectricReflectance\(", "loader/material.cpp"),
             extract_block(mat, r"static float computeDiffuseFresnel\(", "loader/material.cpp"),
             extract_block(mat, r"struct ComplexIor\b", "loader/material.cpp"),
             extract_block(mat, r"static const ComplexIor complexIorList\[\]", "loader/material.cpp"),
             re.search(r"static const int complexIorCount = \d+;", mat).group(0) + "\n",
             extract_block(mat, r"bool complexIorListLookup\(", "loader/material.cpp"),
             "// ======== src/core/camera.cpp ========\n",
             extract_block(cam, r"mat4 perspectiveTransform\(", "core/camera.cpp"),
             extract_block(cam, r"mat4 cameraToRasterTransform\(", "core/camera.cpp"),
             "// ======== src/core/texture.cpp: body of EnvMap::EnvMap ========\n",
             "static void envMapTables(void* m_data, int32_t width, int32_t height, void* m_marginal, void* m_conditional) {\n",
             "\n".join(keep), "\n}\n",
             open(os.path.join(HERE, "shim_host_post.inc")).read()]
    return "".join(parts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default=os.environ.get("ASUNA_REFERENCE", "/root/reference"))
    args = ap.parse_args()
    ref = args.reference
    if not os.path.isdir(os.path.join(ref, "src", "shaders")):
        print(f"build_ref: no reference tree at {ref}; oracle/_ref is left as it is", file=sys.stderr)
        return 0
    os.makedirs(OUT, exist_ok=True)
    gen = os.path.join(OUT, "ref_glsl_gen.cpp")
    with open(gen, "w") as f:
        f.write(generate(ref))
    glm = os.path.join(ref, "src", "ext", "nvpro_core", "third_party", "tinygltf", "examples", "common", "glm")
    inc = os.path.join(os.path.dirname(ORACLE), "include")
    # -ffp-contract=off: no FMA contraction, like the oracle (GLSL `precise`-free code may contract on a GPU;
    #   the CPU pin is about the expression structure, and both sides must round alike to be compared in ulps).
    # -ftrivial-auto-var-init=zero: GLSL locals read before assignment (A.3-5, A.3-7) become zeros instead of
    #   stack garbage, so the build is deterministic.
    common = ["-O2", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-ftrivial-auto-var-init=zero", "-pthread"]
    cmd1 = ["g++", "-std=c++20", *common, "-Werror=float-conversion", "-Werror=double-promotion",
            "-I", HERE, "-I", glm, "-c", gen, "-o", os.path.join(OUT, "ref_glsl_gen.o")]
    cmd2 = ["g++", "-std=c++17", *common, "-DASUNA_REF_SHADERS", "-I", HERE, "-I", inc, "-I", ORACLE,
            "-c", os.path.join(ORACLE, "oracle.cpp"), "-o", os.path.join(OUT, "oracle_ref.o")]
    hgen = os.path.join(OUT, "ref_host_gen.cpp")
    with open(hgen, "w") as f:
        f.write(generate_host(ref))
    cmdh = ["g++", "-std=c++17", *common, "-w", "-I", os.path.join(ref, "src", "ext", "nvpro_core"),
            "-c", hgen, "-o", os.path.join(OUT, "ref_host_gen.o")]
    cmd3 = ["g++", "-shared", "-pthread", "-o", os.path.join(OUT, "libref.so"), os.path.join(OUT, "ref_glsl_gen.o"),
            os.path.join(OUT, "oracle_ref.o"), os.path.join(OUT, "ref_host_gen.o")]
    for cmd in (cmd1, cmd2, cmdh, cmd3):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:  # the compilers' warnings about the reference's text are noise unless the build fails
            sys.stderr.write(r.stdout + r.stderr)
            raise SystemExit(f"build_ref: {' '.join(cmd[:3])} ... failed with {r.returncode}")
    print("built", os.path.join(OUT, "libref.so"))
    return 0


if __name__ == "__main__":
    sys.exit(main())
