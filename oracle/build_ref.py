#!/usr/bin/env python3
"""Build oracle/_ref/: the reference's OWN World::tick / World::tickTemperature, headless.

TEST INFRASTRUCTURE ONLY (see oracle/README.md).  Nothing under fallingsandsurvival_b200/ may
import or load what this produces.

The reference cannot be built with its own build system here (CMake + Conan downloads SDL2,
SDL_gpu, Box2D, easy_profiler, enet, ... at configure time; there is no network), but the hot
path needs none of that.  Recipe:

  * compiled unmodified, straight from /root/reference: Material.cpp, MaterialInstance.cpp,
    Materials.cpp, Tiles.cpp, Particle.cpp and lib/CTPL-ctpl_v.0.0.2/ctpl_stl.h;
  * FallingSandSurvival/world.cpp lines 1053-1062 (getTile/setTile), 1075-1088
    (CalculateVerticalFlowValue), 1091-2018 + 2027-2041 (World::tick minus the Box2D physicsCheck
    block), 2043-2143 (World::tickTemperature) and 2171-2338 (World::tickParticles) are extracted into
    oracle/_ref/gen/tick_lines.inc (and Game.cpp:2584-2650, the dirty -> pixels loop of Game::tick, into
    oracle/_ref/gen/dirty_lines.inc) behind #line directives (so __LINE__, our RNG site id, keeps the
    reference's numbering) with three single-line instrumentation edits listed in PATCHES;
  * shims (oracle/ref_shim/): SDL typedefs, a cut-down `World` declaration, the C API.

libfss_ref_chunk.so (separate, because it needs the system's liblz4): Chunk.cpp lines 47-187 (Chunk::read) and 189-281
(Chunk::write) behind a cut-down `Chunk` declaration (oracle/ref_shim/fss_ref_chunk.cpp), for tests/test_chunkio.py.

No reference SOURCE is written into the repository: the generated file and the libraries live in
oracle/_ref/, which is git-ignored (but travels to the GPU box with gpurun).

Variants (each one .so):
  c4x16    CHUNK 4x16, keyed RNG   -- THE PARITY ORACLE (the schedule the CUDA kernels run)
  c16      CHUNK 16x16, keyed RNG  -- SURVEY's probe schedule, kept for schedule-sensitivity tests
  c128     CHUNK 128x128, keyed RNG
  shipped  CHUNK 128x128, libc rand() -- the reference as shipped (CPU baseline only; not deterministic)
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("FSS_REFERENCE", "/root/reference")
F = os.path.join(REF, "FallingSandSurvival")
OUT = os.path.join(HERE, "_ref")
GEN = os.path.join(OUT, "gen")

GAME_DIRTY_LOOP = (2584, 2650)
RANGES = [(1053, 1062), (1075, 1088), (1091, 2018), (2027, 2041), (2043, 2143), (2171, 2338)]

# (line number, exact text that must be present, replacement) -- all keep the line count
PATCHES = [
    (1159, "int index = x + y * width;", "int index = x + y * width; FSS_REF_CTX(x, y);"),
    (1674, "int index = x + y * width;", "int index = x + y * width; FSS_REF_CTX(x, y);"),
    (1910, "int index = x + y * width;", "int index = x + y * width; FSS_REF_CTX(x, y);"),
    (1288, "getTile(x, y + 4).mat->physicsType == PhysicsType::AIR)",
     "getTile(x, y + 4).mat->physicsType == PhysicsType::AIR && FSS_REF_PARTICLES)"),
    (1352, "getTile(x, y + 4).mat->physicsType == PhysicsType::AIR)",
     "getTile(x, y + 4).mat->physicsType == PhysicsType::AIR && FSS_REF_PARTICLES)"),
]

VARIANTS = {
    "c4x16": ["-DCHUNK_W=4", "-DCHUNK_H=16"],
    "c16": ["-DCHUNK_W=16", "-DCHUNK_H=16"],
    "c128": ["-DCHUNK_W=128", "-DCHUNK_H=128"],
    "shipped": ["-DCHUNK_W=128", "-DCHUNK_H=128", "-DFSS_REF_LIBC_RAND"],
}


def generate():
    os.makedirs(GEN, exist_ok=True)
    with open(os.path.join(F, "world.cpp"), encoding="utf-8", errors="replace") as f:
        lines = f.read().split("\n")
    patched = {}
    for ln, needle, repl in PATCHES:
        src = patched.get(ln, lines[ln - 1])
        if needle not in src:
            raise SystemExit(f"world.cpp:{ln} does not contain {needle!r}; reference changed?")
        patched[ln] = src.replace(needle, repl, 1)
    out = ["/* GENERATED from the reference by oracle/build_ref.py -- never commit */"]
    for a, b in RANGES:
        out.append(f'#line {a} "world.cpp"')
        for ln in range(a, b + 1):
            out.append(patched.get(ln, lines[ln - 1]))
    with open(os.path.join(GEN, "tick_lines.inc"), "w") as f:
        f.write("\n".join(out) + "\n")
    # the `dirty` job of Game::tick (Game.cpp:2584-2650): the loop that turns dirty cells into pixels, as it stands
    with open(os.path.join(F, "Game.cpp"), encoding="utf-8", errors="replace") as f:
        glines = f.read().split("\n")
    a, b = GAME_DIRTY_LOOP
    if "world->dirty[i]" not in glines[a + 2] or "for(int i = 0; i < world->width * world->height; i++)" not in glines[a - 1]:
        raise SystemExit("Game.cpp: the dirty loop is not where it was; reference changed?")
    out = ["/* GENERATED from the reference by oracle/build_ref.py -- never commit */", f'#line {a} "Game.cpp"'] + glines[a - 1:b]
    with open(os.path.join(GEN, "dirty_lines.inc"), "w") as f:
        f.write("\n".join(out) + "\n")


CHUNK_RANGES = [(47, 187), (189, 281)]  # Chunk::read, Chunk::write


def generate_chunk():
    with open(os.path.join(F, "Chunk.cpp"), encoding="utf-8", errors="replace") as f:
        lines = f.read().split("\n")
    if not lines[46].startswith("void Chunk::read()") or not lines[188].startswith("void Chunk::write("):
        raise SystemExit("Chunk.cpp: Chunk::read / Chunk::write are not where they were; reference changed?")
    out = ["/* GENERATED from the reference by oracle/build_ref.py -- never commit */"]
    for a, b in CHUNK_RANGES:
        out.append(f'#line {a} "Chunk.cpp"')
        out += lines[a - 1:b]
    with open(os.path.join(GEN, "chunk_lines.inc"), "w") as f:
        f.write("\n".join(out) + "\n")


def build_chunk(objs):
    """libfss_ref_chunk.so: the reference's Chunk::read / Chunk::write lines + the system liblz4 (skipped where that is missing)"""
    lz4 = next((p for p in ("/usr/lib/x86_64-linux-gnu/liblz4.so.1", "/usr/lib64/liblz4.so.1", "/usr/lib/liblz4.so.1") if os.path.exists(p)), None)
    if lz4 is None:
        print("liblz4.so.1 not found: libfss_ref_chunk.so not built", file=sys.stderr)
        return
    shim = os.path.join(HERE, "ref_shim")
    so = os.path.join(OUT, "libfss_ref_chunk.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-w", "-shared", "-I", shim, "-I", F, "-I", GEN, "-include", "vector", "-include", "cstdio",
                           "-o", so, os.path.join(shim, "fss_ref_chunk.cpp")] + [o for o in objs if not o.endswith("Particle.o")] + [lz4])
    print("built", so)


def build(variants):
    shim = os.path.join(HERE, "ref_shim")
    common = ["g++", "-O2", "-std=c++17", "-pthread", "-fPIC", "-ffp-contract=off", "-w",
              "-I", shim, "-I", F, "-I", os.path.join(F, "lib", "CTPL-ctpl_v.0.0.2"), "-I", GEN]
    objs = []
    # the reference's own translation units, unmodified
    for name in ["Material", "MaterialInstance", "Materials", "Particle"]:
        o = os.path.join(GEN, name + ".o")
        subprocess.check_call(common + ["-include", "vector", "-include", "cstdio", "-c",
                                        os.path.join(F, name + ".cpp"), "-o", o])
        objs.append(o)
    pre = os.path.join(shim, "fss_ref_pre.h")
    for v in variants:
        keyed = "-DFSS_REF_LIBC_RAND" not in VARIANTS[v]
        tiles_o = os.path.join(GEN, f"Tiles_{'keyed' if keyed else 'libc'}.o")
        tiles_flags = ["-include", pre, "-DFSS_REF_SITE_BASE=10000"] + ([] if keyed else ["-DFSS_REF_LIBC_RAND"])
        subprocess.check_call(common + tiles_flags + ["-c", os.path.join(F, "Tiles.cpp"), "-o", tiles_o])
        api_o = os.path.join(GEN, f"fss_ref_{v}.o")
        subprocess.check_call(common + VARIANTS[v] + ["-include", pre, "-c",
                                                     os.path.join(shim, "fss_ref.cpp"), "-o", api_o])
        so = os.path.join(OUT, f"libfss_ref_{v}.so")
        subprocess.check_call(["g++", "-shared", "-pthread", "-o", so, api_o, tiles_o] + objs)
        print("built", so)
    build_chunk(objs)


def main():
    if not os.path.isdir(F):
        print(f"reference not found at {REF}; keeping prebuilt oracle/_ref as is", file=sys.stderr)
        return 0
    variants = sys.argv[1:] or list(VARIANTS)
    generate()
    generate_chunk()
    build(variants)
    return 0


if __name__ == "__main__":
    sys.exit(main())
