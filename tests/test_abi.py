"""The C-ABI boundary without a GPU: the library loads, exports every symbol include/fss.h declares, and the
ctypes / numpy mirrors in abi.py have the layouts the C compiler gives the header's structs."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from fallingsandsurvival_b200 import abi, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "fss.h")


@pytest.fixture(scope="module")
def lib():
    return C.CDLL(build.build())


def declared_functions():
    with open(HEADER) as f:
        return re.findall(r"^(?:int|const char\*) (fss_\w+)\(", f.read(), flags=re.M)


def test_library_exports_every_declared_symbol(lib):
    names = declared_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), n


def test_python_binding_declares_every_symbol():
    from fallingsandsurvival_b200 import world
    src = open(os.path.join(ROOT, "fallingsandsurvival_b200", "world.py")).read()
    for n in declared_functions():
        assert n in src, f"world.py does not bind {n}"
    assert world.LIB_PATH.endswith("libfss_b200.so")


def test_abi_version_and_no_device_error(lib):
    assert lib.fss_abi_version() == abi.FSS_ABI_VERSION
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    cfg = abi.Config()
    cfg.struct_size = C.sizeof(abi.Config)
    cfg.width = cfg.height = 256
    h = C.c_void_p()
    rc = lib.fss_create(C.byref(cfg), C.byref(h))
    assert rc == -4  # FSS_ERR_NO_DEVICE: the product has no CPU fallback
    lib.fss_last_error.restype = C.c_char_p
    assert b"no CPU fallback" in lib.fss_last_error()


def test_world_class_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from fallingsandsurvival_b200 import world
    with pytest.raises(world.FssError) as e:
        world.World(None, None, size=(256, 256))
    assert e.value.code == -4


def test_struct_layouts_match_the_header(tmp_path):
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "fss.h"
#define S(t) printf(#t " %zu\n", sizeof(t))
#define O(t, f) printf(#t "." #f " %zu\n", offsetof(t, f))
int main(void) {
  S(fss_cell); O(fss_cell, color); O(fss_cell, temperature); O(fss_cell, fluid_amount); O(fss_cell, fluid_amount_diff);
  S(fss_reaction); S(fss_material); O(fss_material, reactions); O(fss_material, create_kind); O(fss_material, create_temperature);
  S(fss_config); O(fss_config, strip_y0); O(fss_config, condense_material);
  S(fss_particle_spawn); O(fss_particle_spawn, cell);
  S(fss_particle); O(fss_particle, target_force); O(fss_particle, lifetime); O(fss_particle, in_object_state); O(fss_particle, phase); O(fss_particle, temporary); O(fss_particle, tile);
  S(fss_render_info); O(fss_render_info, had_flow);
  S(fss_fill_spec); O(fss_fill_spec, cumulative_percent); O(fss_fill_spec, solid_from_row); O(fss_fill_spec, sparse_per_mille);
  S(fss_stats); O(fss_stats, storage_pitch);
  S(fss_tile_record); O(fss_tile_record, color); O(fss_tile_record, temperature);
  printf("CHUNK %d %d %d %d %d\n", FSS_CHUNK_W, FSS_CHUNK_H, FSS_ZONE_ALIGN_X, FSS_ZONE_ALIGN_Y, FSS_ZONE_MARGIN);
  return 0;
}'''
    src = tmp_path / "layout.c"
    src.write_text(prog)
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    out = dict(line.rsplit(" ", 1) for line in subprocess.check_output([str(exe)], text=True).splitlines() if not line.startswith("CHUNK"))
    out = {k: int(v) for k, v in out.items()}
    py = {"fss_cell": abi.CellStruct, "fss_reaction": abi.Reaction, "fss_material": abi.Material, "fss_config": abi.Config,
          "fss_particle_spawn": abi.ParticleSpawn, "fss_fill_spec": abi.FillSpec, "fss_stats": abi.Stats,
          "fss_particle": abi.Particle, "fss_render_info": abi.RenderInfo}
    assert abi.RECORD_DTYPE.itemsize == out.pop("fss_tile_record") == 12
    assert abi.RECORD_DTYPE.fields["color"][1] == out.pop("fss_tile_record.color")
    assert abi.RECORD_DTYPE.fields["temperature"][1] == out.pop("fss_tile_record.temperature")
    for key, val in out.items():
        if "." in key:
            t, f = key.split(".")
            assert getattr(py[t], f).offset == val, key
        else:
            assert C.sizeof(py[key]) == val, key
    assert abi.CELL_DTYPE.itemsize == out["fss_cell"]
    assert abi.CELL_DTYPE.fields["fluid_amount_diff"][1] == out["fss_cell.fluid_amount_diff"]
    assert abi.PARTICLE_DTYPE.itemsize == out["fss_particle_spawn"]
    assert abi.LIVE_PARTICLE_DTYPE.itemsize == out["fss_particle"] == 68
    for f in ("target_force", "lifetime", "in_object_state", "phase", "temporary", "tile"):
        assert abi.LIVE_PARTICLE_DTYPE.fields[f][1] == out["fss_particle." + f], f
    consts = subprocess.check_output([str(exe)], text=True).splitlines()[-1].split()[1:]
    assert [int(c) for c in consts] == [abi.FSS_CHUNK_W, abi.FSS_CHUNK_H, abi.FSS_ZONE_ALIGN_X, abi.FSS_ZONE_ALIGN_Y, abi.FSS_ZONE_MARGIN]


def test_spawn_order_key_is_the_same_in_c_and_python(tmp_path):
    """fss_spawn_order_key (fss.h) is what the device sorts by; abi.spawn_order is what the Python side merges strips with"""
    rng = np.random.default_rng(5)
    n = 4000
    zone = (24, 40)
    recs = np.zeros(n, dtype=abi.PARTICLE_DTYPE)
    recs["tick_iter"] = rng.integers(0, 6, n) + 8 * rng.integers(0, 3, n)
    recs["x"] = rng.integers(zone[0], zone[0] + 512, n)
    recs["y"] = rng.integers(zone[1], zone[1] + 512, n)
    prog = r'''
#include <stdio.h>
#include "fss.h"
int main(void) {
  unsigned t, x, y;
  while(scanf("%u %u %u", &t, &x, &y) == 3) printf("%llu\n", (unsigned long long)fss_spawn_order_key(t, x, y, 24, 40));
  return 0;
}'''
    src = tmp_path / "key.c"
    src.write_text(prog)
    exe = tmp_path / "key"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    text = "".join(f"{int(r['tick_iter'])} {int(r['x'])} {int(r['y'])}\n" for r in recs)
    keys = np.array([int(v) for v in subprocess.check_output([str(exe)], input=text, text=True).split()], dtype=np.uint64)
    assert len(keys) == n
    order = abi.spawn_order(recs, zone)
    assert np.array_equal(keys[order], np.sort(keys, kind="stable"))
    # equal keys only for equal (tick_iter, x, y)
    uniq = {}
    for k, r in zip(keys, recs):
        assert uniq.setdefault(int(k), (int(r["tick_iter"]), int(r["x"]), int(r["y"]))) == (int(r["tick_iter"]), int(r["x"]), int(r["y"]))


def test_product_never_imports_the_oracle():
    """the oracle is test infrastructure: nothing under fallingsandsurvival_b200/ may reference it"""
    pkg = os.path.join(ROOT, "fallingsandsurvival_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oraclesim" not in text and "refsim" not in text and "fss_oracle" not in text, f


def build_c_caller(tmp_path):
    exe = tmp_path / "abi_smoke"
    pkg = os.path.join(ROOT, "fallingsandsurvival_b200")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-I", os.path.join(ROOT, "include"), "-o", str(exe),
                           os.path.join(ROOT, "tests", "c", "abi_smoke.c"), "-L", pkg, "-lfss_b200", f"-Wl,-rpath,{pkg}"])
    return exe


def test_plain_c_program_links_against_the_library(lib, tmp_path):
    """the header is C (not C++), and a C caller links: the reference-side shim needs nothing else"""
    exe = build_c_caller(tmp_path)
    import torch
    if torch.cuda.is_available():
        return
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 3 and "no CPU fallback" in r.stderr  # FSS_ERR_NO_DEVICE, loudly


def build_cpp_shim_test(tmp_path):
    exe = tmp_path / "world_shim_test"
    pkg = os.path.join(ROOT, "fallingsandsurvival_b200")
    subprocess.check_call(["g++", "-std=c++14", "-Wall", "-I", os.path.join(ROOT, "include"), "-o", str(exe),
                           os.path.join(ROOT, "tests", "cpp", "world_shim_test.cpp"), "-L", pkg, "-lfss_b200", f"-Wl,-rpath,{pkg}"])
    return exe


def test_cpp_world_mirror_compiles_as_cxx14(lib, tmp_path):
    """include/fss_world.hpp (the C++ mirror of the reference's World for this path) builds with the reference's own
    language level (C++14, CMakeLists.txt:325) and fails loudly without a GPU"""
    exe = build_cpp_shim_test(tmp_path)
    import torch
    if torch.cuda.is_available():
        return
    r = subprocess.run([str(exe), str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 3 and "no CPU fallback" in r.stderr
