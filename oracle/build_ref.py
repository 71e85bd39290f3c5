#!/usr/bin/env python3
"""Builds oracle/_ref/libfloor_ref_minify.so: the REFERENCE's own Host-Compute image sampler, compiled with g++.

TEST INFRASTRUCTURE ONLY (the checker of the checker): it pins oracle/minify_oracle.c, the C restatement, against the
arithmetic of the reference's own sources.  Nothing of the product path may load it.

What is compiled: all 17 kernels `libfloor_mip_map_minify_<IMAGE>_<SAMPLE>` of include/floor/device/backend/mip_map_minify.hpp,
`fl::image<>::read_lod_linear / write_lod` (image.hpp), `fl::host_device_image<...>::read_linear / write`
(host_image.hpp), the vector / const_math / soft_f16 / image_types headers they pull in (see NEEDED) and the reference's own
explicit vector instantiations (REF_SOURCES); oracle/ref_harness.cpp (ours) drives the kernels in the level / layer loop of
src/device/device_image.cpp:290-327.

libfloor only supports clang >= 19 (include/floor/floor_version.hpp:96-129) and these headers use clang-only extensions that
g++ 13 cannot parse.  The headers are therefore read where they lie under /root/reference, a SMALL LIST OF MECHANICAL,
arithmetic-neutral substitutions (PATCHES below; each one says what and why) is applied in a temporary directory, and only
the resulting shared object is kept, in oracle/_ref/ (git-ignored).  No reference source is copied into the repository.
None of the substitutions changes what a line computes: coordinate handling, texel fetch, format decode, the lerp
(`const_math::interpolate`), format encode and the level-size math are the reference's own code (the 1D sampler and the depth
read get an implicit vector1 -> scalar conversion spelled out, nothing else on the texel path is touched), compiled as strict IEEE
(-fno-fast-math -ffp-contract=off), which is the canonical numeric mode of SURVEY.md 8c.

/root/reference does not exist on the GPU box: the .so is built here (by __graft_entry__.build()) and travels with the repo
snapshot; tests that use it skip when it is absent.
"""
import os
import re
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("FLOOR_REFERENCE", "/root/reference")
OUT_DIR = os.path.join(HERE, "_ref")
OUT = os.path.join(OUT_DIR, "libfloor_ref_minify.so")
OUT_NO_DOUBLE = os.path.join(OUT_DIR, "libfloor_ref_minify_nodouble.so")
# timing build for bench.py's reference arm: the reference's own release flags (build.sh:730-750, CMakeLists.txt:31:
# -O3 -funroll-loops -ffast-math -fstrict-aliasing -march=corei7-avx -mf16c) WITHOUT -ffast-math: with g++ 13 the combination
# -funsafe-math-optimizations + -ffinite-math-only + -fno-signed-zeros rewrites the sampler's weight computation and the
# texels come out wrong (bilinear weights of 0.75 / 0.6 instead of 0.5), so a fast-math g++ build is not the reference's
# arithmetic.  tests/test_reference_pin.py checks that this build computes the same bytes as the strict one.
OUT_FAST = os.path.join(OUT_DIR, "libfloor_ref_minify_fast.so")
STRICT_FLAGS = ["-O2", "-fno-fast-math", "-ffp-contract=off"]
FAST_FLAGS = ["-O3", "-funroll-loops", "-fstrict-aliasing", "-ffp-contract=off", "-march=corei7-avx", "-mf16c"]

# transitive include set of host_image.hpp + image_types.hpp + vector_lib.hpp (g++ -M)
NEEDED = [
    "floor/floor_conf.hpp",
    "floor/core/essentials.hpp",
    "floor/core/enum_helpers.hpp",
    "floor/constexpr/ext_traits.hpp",
    "floor/constexpr/const_math.hpp",
    "floor/constexpr/const_array.hpp",
    "floor/constexpr/soft_f16.hpp",
    "floor/math/constants.hpp",
    "floor/math/rt_math.hpp",
    "floor/math/matrix4.hpp",
    "floor/math/vector.hpp",
    "floor/math/vector_helper.hpp",
    "floor/math/vector_lib.hpp",
    "floor/math/vector_ops.hpp",
    "floor/math/vector_ops_cleanup.hpp",
    "floor/device/backend/host_limits.hpp",
    "floor/device/backend/host_pre.hpp",
    "floor/device/backend/device_info.hpp",
    "floor/device/backend/sampler.hpp",
    "floor/device/backend/image_types.hpp",
    "floor/device/backend/host_image.hpp",
    "floor/device/backend/image.hpp",
    "floor/device/backend/mip_map_minify.hpp",
]

# reference translation units compiled as they are (paths relative to the reference root)
REF_SOURCES = [
    "src/math/vector.cpp",
    "src/math/vector_1d.cpp",
    "src/math/vector_2d.cpp",
    "src/math/vector_3d.cpp",
    "src/math/vector_4d.cpp",
]

# (file, pattern, replacement, count, why) -- regex, re.S
PATCHES = [
    # clang's __attribute__((enable_if(...))) overload sets (compile-time-index bounds diagnostics): g++ cannot parse them.
    # The run-time overload that remains is the one every caller here would resolve to anyway.
    ("floor/constexpr/const_array.hpp", r"#if !defined\(_MSC_VER\) // duplicate name mangling issues", "#if 0 /* g++: no enable_if attribute */", 4,
     "drop the constant-index overloads of const_array::operator[] / at"),
    ("floor/math/vector.hpp",
     r"\tconstexpr (?:const )?scalar_type& operator\[\]\(const uint32_t& index\)(?: const)?\n\t__attribute__\(\(enable_if\([^\n]*unavailable\(\"index out of bounds\"\)\)\);\n",
     "", 2, "drop the 'unavailable' out-of-bounds declarations of vector::operator[]"),
    ("floor/math/matrix4.hpp",
     r"\tconstexpr (?:const )?scalar_type& operator\[\]\(const size_t& index\)(?: const)?\n\t__attribute__\(\(enable_if\([^\n]*unavailable\(\"index out of bounds\"\)\)\);\n",
     "", 2, "drop the 'unavailable' out-of-bounds declarations of matrix4::operator[]"),
    ("floor/math/matrix4.hpp",
     r"static constexpr matrix4 perspective\(const scalar_type fov, const scalar_type aspect,\n(\s*)const scalar_type z_near, const scalar_type z_far\)\n\t__attribute__\(\(enable_if\(fov == fov, \"perspective with constant field-of-view\"\)\)\) \{",
     r"static constexpr matrix4 perspective_constant_fov(const scalar_type fov, const scalar_type aspect,\n\1const scalar_type z_near, const scalar_type z_far) {",
     1, "constant-fov overload of matrix4::perspective (unused here) gets its own name"),
    ("floor/constexpr/const_math.hpp",
     r"\n\t__attribute__\(\(enable_if\(!__builtin_constant_p\(&n\) \|\| \(__builtin_constant_p\(&n\) && n <= 67\), \"64-bit range\"\)\)\) \{",
     " {", 1, "binomial(): range diagnostic attribute (unused here)"),
    # math::<fn> = compile-time or run-time implementation chosen by clang enable_if + asm-label forwarding.  g++ spelling
    # of the same selection: std::is_constant_evaluated().  (Only math::floor / math::min / max / clamp level helpers are on
    # the image path; the lerp itself is const_math::interpolate, untouched.)
    ("floor/constexpr/const_math.hpp",
     r"#define FLOOR_CONST_SELECT\(ARG_EXPANDER, ENABLE_IF_EXPANDER, func_name, ce_func, rt_func, type, overload_suffix\) \\\n.*?return rt_func \(ARG_EXPANDER\(, FLOOR_COMMA\)\); \\\n\t\}\n\t\n#define FLOOR_CONST_SELECT_1",
     "#define FLOOR_CONST_SELECT(ARG_EXPANDER, ENABLE_IF_EXPANDER, func_name, ce_func, rt_func, type, overload_suffix) \\\n"
     "\tstatic __attribute__((always_inline)) inline constexpr auto func_name (ARG_EXPANDER(const type, FLOOR_COMMA)) { \\\n"
     "\t\tif (std::is_constant_evaluated()) { return ce_func (ARG_EXPANDER(, FLOOR_COMMA)); } \\\n"
     "\t\telse { return rt_func (ARG_EXPANDER(, FLOOR_COMMA)); } \\\n"
     "\t} \\\n"
     "\tstatic __attribute__((always_inline)) inline constexpr auto __ ## func_name (ARG_EXPANDER(const type, FLOOR_COMMA)) { \\\n"
     "\t\tif (std::is_constant_evaluated()) { return ce_func (ARG_EXPANDER(, FLOOR_COMMA)); } \\\n"
     "\t\telse { return rt_func (ARG_EXPANDER(, FLOOR_COMMA)); } \\\n"
     "\t}\n\t\n#define FLOOR_CONST_SELECT_1",
     1, "const-select macro: is_constant_evaluated() instead of enable_if/asm-label overloads"),
    # clang converts vector1<T> <-> T implicitly where g++ sees an ambiguity (vector1's own operator< through the converting
    # constructor vs the built-in one through the conversion operator; fit(T) vs fit(vector4<T>)).  The 1D sampler
    # (host_image.hpp:859-865) and the 1-channel depth read (image.hpp:519) get the conversion spelled out: same value.
    ("floor/device/backend/host_image.hpp", r"\(frac_coord < 0\.5f \? frac_coord \+ 0\.5f : 1\.5f - frac_coord\)",
     "(frac_coord.x < 0.5f ? frac_coord.x + 0.5f : 1.5f - frac_coord.x)", 2,
     "1D read_linear / compare_linear: the weight as an explicit scalar (frac_coord is a vector1<float>)"),
    ("floor/device/backend/host_image.hpp", r"frac_coord < 0\.5f", "frac_coord.x < 0.5f", 2,
     "1D read_linear / compare_linear: explicit vector1<float> -> float in the neighbour-offset comparison"),
    ("floor/device/backend/image.hpp", r"(#else\n\t+)return output_type::fit\(color\);",
     r"\1if constexpr (std::is_same_v<std::decay_t<decltype(color)>, vector1<sample_type>>) { return output_type::fit(color.x); } else { return output_type::fit(color); }",
     1, "depth read: explicit vector1<float> -> float before image_vec_ret_type::fit"),
    # FLOOR_DEVICE_NO_DOUBLE (device-run Host-Compute builds: all-float encoder scale, host_image.hpp:398-402) cannot be
    # defined for the whole host-mode header set (vector_helper.hpp:983 drops double while vector_lib.hpp:85-100 still lists
    # it), so the one #if on the path gets its own switch; the second build defines it.
    ("floor/device/backend/host_image.hpp", r"#if !defined\(FLOOR_DEVICE_NO_DOUBLE\)\n(\t+using fp_scale_type = )",
     r"#if !defined(FLOOR_REF_NO_DOUBLE)\n\1", 1, "own switch for the reference's all-float encoder branch"),
]


def apply_patches(root):
    for rel, pat, repl, count, why in PATCHES:
        path = os.path.join(root, rel)
        with open(path, "r", encoding="utf-8") as f:
            src = f.read()
        new, n = re.subn(pat, repl, src, flags=re.S | re.M)
        if n != count:
            raise RuntimeError(f"patch for {rel} ({why}) matched {n} times, expected {count}")
        with open(path, "w", encoding="utf-8") as f:
            f.write(new)


def build(verbose=False, keep=None):
    """Returns the path of the double-scale build (the canonical mode); the FLOOR_DEVICE_NO_DOUBLE build sits beside it."""
    if not os.path.isdir(os.path.join(REF, "include", "floor")):
        return OUT if os.path.exists(OUT) else None  # GPU box: use the prebuilt files
    harness = os.path.join(HERE, "ref_harness.cpp")
    srcs = [os.path.join(REF, "include", r) for r in NEEDED] + [os.path.join(REF, r) for r in REF_SOURCES]
    stamp = max(os.path.getmtime(p) for p in srcs + [harness, os.path.join(HERE, "ref_shim.hpp"), os.path.abspath(__file__)])
    if all(os.path.exists(o) and os.path.getmtime(o) >= stamp for o in (OUT, OUT_NO_DOUBLE, OUT_FAST)):
        return OUT
    tmp = keep or tempfile.mkdtemp(prefix="floor_ref_")
    try:
        for rel in NEEDED:
            s = os.path.join(REF, "include", rel)
            d = os.path.join(tmp, "include", rel)
            os.makedirs(os.path.dirname(d), exist_ok=True)
            shutil.copyfile(s, d)
        apply_patches(os.path.join(tmp, "include"))
        os.makedirs(OUT_DIR, exist_ok=True)
        # -D substitutions: x86 g++ spells the IEEE half type _Float16 (clang: __fp16); three clang builtins; GCC turns a
        # failed always_inline (recursive shmwrap<bool>, rt_math.hpp) into an error, so the attribute becomes a no-op
        # (inlining cannot change results under -fno-fast-math -ffp-contract=off)
        base = ["g++", "-std=gnu++2b", "-fPIC", "-pthread",
                "-w", "-fpermissive", "-D__fp16=_Float16", "-D__builtin_clzs=floor_ref_clzs", "-D__builtin_ctzs=floor_ref_ctzs",
                "-D__builtin_assume(x)=((void)0)", "-Dalways_inline=unused", "-DFLOOR_DEVICE_HOST_COMPUTE=1",
                "-include", os.path.join(HERE, "ref_shim.hpp"), "-I", os.path.join(tmp, "include")]
        jobs = []
        objs = {"strict": [], "fast": []}
        for mode, flags in (("strict", STRICT_FLAGS), ("fast", FAST_FLAGS)):
            for src in REF_SOURCES:  # the reference's own explicit instantiations of fl::vectorN<T>, compiled where they lie
                o = os.path.join(tmp, mode + "_" + os.path.basename(src)[:-4] + ".o")
                objs[mode].append(o)
                jobs.append((o, base + flags + ["-c", os.path.join(REF, src), "-o", o]))
        for name, flags in (("harness.o", STRICT_FLAGS), ("harness_nd.o", STRICT_FLAGS + ["-DFLOOR_REF_NO_DOUBLE=1"]),
                            ("harness_fast.o", FAST_FLAGS)):
            jobs.append((os.path.join(tmp, name), base + flags + ["-c", harness, "-o", os.path.join(tmp, name)]))
        procs = [(o, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)) for o, cmd in jobs]
        for o, p in procs:
            out, _ = p.communicate()
            if p.returncode != 0:
                if verbose:
                    sys.stderr.write(out)
                raise RuntimeError("g++ failed on the patched reference sources (" + o + "):\n" + out[-4000:])
        for out_so, h, mode in ((OUT, "harness.o", "strict"), (OUT_NO_DOUBLE, "harness_nd.o", "strict"), (OUT_FAST, "harness_fast.o", "fast")):
            subprocess.check_call(["g++", "-shared", "-pthread", "-o", out_so, os.path.join(tmp, h)] + objs[mode])
        return OUT
    finally:
        if keep is None:
            shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    print(build(verbose=True, keep=os.environ.get("FLOOR_REF_KEEP")))
