"""The reference's chunk FILE (Chunk::read / Chunk::write, Chunk.cpp:47-281) on the host side of the C ABI: the LZ4 block codec
of csrc/fss_chunkio.h and the framing around it, checked in both directions against
  * the system's liblz4 (the very library the game links: LZ4_compress_fast(.., 10) / LZ4_decompress_safe), and
  * the reference's OWN Chunk::write / Chunk::read lines compiled into oracle/_ref/libfss_ref_chunk.so.
Host code only: runs without a GPU."""
import ctypes as C
import os

import numpy as np
import pytest

from helpers import full_table, make_world
from fallingsandsurvival_b200 import abi, world

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_CHUNK = os.path.join(ROOT, "oracle", "_ref", "libfss_ref_chunk.so")


def _liblz4():
    try:
        L = C.CDLL("liblz4.so.1")
    except OSError:
        pytest.skip("no system liblz4")
    L.LZ4_compress_fast.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int]
    L.LZ4_decompress_safe.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int]
    L.LZ4_compressBound.argtypes = [C.c_int]
    return L


def _samples():
    rng = np.random.default_rng(5)
    cells, _ = make_world((128, 128), "c3", seed=12)
    recs = np.zeros((128, 128), dtype=abi.RECORD_DTYPE)
    recs["index"], recs["color"], recs["temperature"] = cells["material"], cells["color"], cells["temperature"]
    return [b"", b"a", b"abc", b"a" * 100, b"abcd" * 5000, bytes(rng.integers(0, 256, 50000, dtype=np.uint8)),
            bytes(rng.integers(0, 4, 200000, dtype=np.uint8)), b"x" * 70000 + b"y" * 300 + b"x" * 70000,
            bytes(range(13)), b"q" * 12 + b"r", recs.tobytes(), np.zeros(128 * 128, dtype=np.uint32).tobytes()]


def test_lz4_blocks_are_interchangeable_with_liblz4():
    lz = _liblz4()
    for data in _samples():
        mine = world.lz4_compress(data)
        assert world.lz4_decompress(mine, len(data)) == data
        out = C.create_string_buffer(max(len(data), 1))
        assert lz.LZ4_decompress_safe(mine, out, len(mine), len(data)) == len(data) and out.raw[:len(data)] == data  # liblz4 reads ours
        if data:
            cap = lz.LZ4_compressBound(len(data))
            buf = C.create_string_buffer(cap)
            for acceleration in (1, 10):  # Chunk.cpp:226 uses 10
                n = lz.LZ4_compress_fast(data, buf, len(data), cap, acceleration)
                assert world.lz4_decompress(buf.raw[:n], len(data)) == data                                          # we read liblz4's


def test_lz4_known_blocks_and_malformed_input():
    """hand-assembled blocks per the block format: literals only; an overlapping match (run-length); long lengths with 255-bytes"""
    assert world.lz4_decompress(bytes([0x50]) + b"hello", 5) == b"hello"
    # token 0x1F: 1 literal 'a', match offset 1, length 15 + 4 + extra byte 5 = 24 -> 'a' * 25, then a last sequence of 5 literals
    blk = bytes([0x1F]) + b"a" + bytes([0x01, 0x00, 0x05]) + bytes([0x50]) + b"bcdef"
    assert world.lz4_decompress(blk, 30) == b"a" * 25 + b"bcdef"
    # 300 literals: length nibble 15, then 255 and 30
    lit = bytes(range(256)) + bytes(44)
    assert world.lz4_decompress(bytes([0xF0, 255, 30]) + lit, 300) == lit
    for bad in (b"", bytes([0x10]), bytes([0x1F]) + b"a" + bytes([0x00, 0x00, 0x05]),   # empty, literal missing, offset 0
                bytes([0x1F]) + b"a" + bytes([0x02, 0x00, 0x05]),                           # offset beyond the output
                bytes([0x50]) + b"hello!"):                                                 # does not fit (cap 5)
        with pytest.raises(world.FssError):
            world.lz4_decompress(bad, 5)


def _chunk_layers(seed):
    rng = np.random.default_rng(seed)
    out = []
    for k, mix in enumerate(("c3", "sand")):
        cells, _ = make_world((128, 128), mix, seed=seed + k)
        recs = np.zeros((128, 128), dtype=abi.RECORD_DTYPE)
        recs["index"], recs["color"], recs["temperature"] = cells["material"], cells["color"], cells["temperature"]
        out.append(recs)
    return out[0], out[1], rng.integers(0, 1 << 32, (128, 128), dtype=np.uint32) & np.uint32(0xFF0000FF)


def _same(a, b):
    return all(np.array_equal(a[f], b[f]) for f in ("index", "color", "temperature"))


def test_chunk_file_round_trip_and_errors():
    tiles, layer2, bg = _chunk_layers(3)
    data = world.chunk_file_encode(5, tiles, layer2, bg)
    phase, t2, l2, b2 = world.chunk_file_decode(data)
    assert phase == 5 and _same(t2, tiles) and _same(l2, layer2) and np.array_equal(b2, bg)
    assert len(data) < 17 + 128 * 128 * (2 * 12 + 4)
    for bad in (data[:10], data[:100], data[:5] + b"\x00\x00\x00\x00" + data[9:], data[:-1]):
        with pytest.raises(world.FssError):
            world.chunk_file_decode(bad)


@pytest.fixture(scope="module")
def ref_chunk():
    if not os.path.exists(REF_CHUNK):
        pytest.skip("oracle/_ref/libfss_ref_chunk.so not built (python oracle/build_ref.py; needs the system liblz4)")
    L = C.CDLL(REF_CHUNK)
    L.fssrefchunk_init.argtypes = [C.c_uint]
    L.fssrefchunk_write.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.fssrefchunk_read.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.c_void_p, C.c_void_p, C.c_void_p]
    L.fssrefchunk_init(1234)
    return L


def test_chunk_files_interchange_with_the_reference_build(ref_chunk, tmp_path):
    """a file written by the reference's own Chunk::write lines decodes here to the same records; a file encoded here is read
    by the reference's own Chunk::read lines without complaint and yields the same records"""
    for seed in (1, 2):
        tiles, layer2, bg = _chunk_layers(seed)
        path = str(tmp_path / f"ref_{seed}.chunk").encode()
        assert ref_chunk.fssrefchunk_write(path, 3, tiles.ctypes.data, layer2.ctypes.data, bg.ctypes.data) == 0
        with open(path, "rb") as f:
            phase, t2, l2, b2 = world.chunk_file_decode(f.read())
        assert phase == 3 and _same(t2, tiles) and _same(l2, layer2) and np.array_equal(b2, bg)

        mine = str(tmp_path / f"mine_{seed}.chunk").encode()
        with open(mine, "wb") as f:
            f.write(world.chunk_file_encode(-2, tiles, layer2, bg))
        t3, l3 = np.zeros_like(tiles), np.zeros_like(layer2)
        b3 = np.zeros_like(bg)
        ph = C.c_int()
        assert ref_chunk.fssrefchunk_read(mine, C.byref(ph), t3.ctypes.data, l3.ctypes.data, b3.ctypes.data) == 0
        assert ph.value == -2 and _same(t3, tiles) and _same(l3, layer2) and np.array_equal(b3, bg)
