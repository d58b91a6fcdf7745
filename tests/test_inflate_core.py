"""The DEFLATE decoder and the CRC-32 the GPU runs per warp (indelope_b200/csrc/inflate_core.cuh), built here as plain C++ with one lane
(tests/native/inflate_host.cpp) and checked against zlib streams of every block type -- no GPU needed.  The 32-lane execution of the same
source is covered by tests/test_gpu_bam.py."""
import ctypes as C
import os
import random
import subprocess
import zlib

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module", params=[0, 2048], ids=["no-ring", "ring2048"])
def L(request):
    """both build variants of the decoder: matches read back from the output, or from a ring of the last 2 KiB where they reach no further"""
    ring = request.param
    src = os.path.join(HERE, "native", "inflate_host.cpp")
    so = os.path.join(HERE, "native", "libinflate_host_r%d.so" % ring)
    core = os.path.join(ROOT, "indelope_b200", "csrc", "inflate_core.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(core)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-x", "c++", "-DIDL_INF_RING=%d" % ring, "-I", os.path.join(ROOT, "indelope_b200", "csrc"), src, "-o", so])
    lib = C.CDLL(so)
    lib.idl_test_inflate.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_uint32]
    lib.idl_test_inflate_at.argtypes = [C.c_char_p, C.c_size_t, C.c_size_t, C.c_char_p, C.c_uint32]
    lib.idl_test_crc32.argtypes = [C.c_char_p, C.c_uint32, C.c_int]
    lib.idl_test_crc32.restype = C.c_uint32
    return lib


def _payloads():
    rng = random.Random(1)
    r = np.random.default_rng(2)
    yield b""
    yield b"a"
    yield b"abc" * 1000
    yield bytes(rng.getrandbits(8) for _ in range(5000))                      # incompressible: stored blocks at any level
    yield bytes(rng.choice(b"ACGT") for _ in range(60000))
    yield b"\0" * 65280                                                       # distance-1 runs of the maximum match length
    yield bytes((i * 7) & 255 for i in range(65280))
    yield (r.integers(0, 4, 65000) + r.integers(0, 2, 65000) * 40).astype(np.uint8).tobytes()
    yield (r.integers(0, 256, 30000) % r.integers(1, 200, 30000)).astype(np.uint8).tobytes() + b"x" * 300 + bytes(range(256)) * 20
    # repeats at distances around and beyond the decoder's ring of recent output (2048 bytes): matches served from the ring and from global memory
    for period in (1700, 1789, 1790, 1791, 2047, 2048, 2049, 2300, 3000, 20000):
        blk = bytes(rng.getrandbits(8) for _ in range(period))
        yield (blk * (65000 // period + 1))[:65000]


def raw_deflate(d, level, strategy):
    co = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
    return co.compress(d) + co.flush()


def test_every_block_type_against_zlib(L):
    rng = random.Random(3)
    for d in _payloads():
        for level in (0, 1, 4, 6, 9):
            for strategy in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FILTERED):
                c = raw_deflate(d, level, strategy)
                out = C.create_string_buffer(len(d) + 1)
                assert L.idl_test_inflate(c, len(c), out, len(d)) == 0 and out.raw[:len(d)] == d, (len(d), level, strategy)
                off = rng.randrange(1, 23)                                    # the deflate data of a BGZF member is not word aligned
                assert L.idl_test_inflate_at(c, off, len(c), out, len(d)) == 0 and out.raw[:len(d)] == d, (len(d), level, strategy, off)


def test_multi_block_streams(L):
    """several deflate blocks in one member, stored + fixed + dynamic mixed (Z_FULL_FLUSH ends a block and byte-aligns with an empty stored block)"""
    rng = random.Random(5)
    parts = [bytes(rng.choice(b"ACGTN") for _ in range(rng.randrange(1, 9000))) for _ in range(7)]
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    c = b""
    for k, p in enumerate(parts):
        c += co.compress(p) + co.flush(zlib.Z_FULL_FLUSH if k % 2 else zlib.Z_SYNC_FLUSH)
    c += co.flush()
    d = b"".join(parts)
    out = C.create_string_buffer(len(d))
    assert L.idl_test_inflate(c, len(c), out, len(d)) == 0 and out.raw == d


def test_corrupt_streams_are_rejected(L):
    """flipped bits either fail the decoder or change the output (which the member's CRC then catches); wrong sizes always fail"""
    rng = random.Random(7)
    d = bytes(rng.choice(b"ACGT") for _ in range(20000))
    c = raw_deflate(d, 6, zlib.Z_DEFAULT_STRATEGY)
    out = C.create_string_buffer(len(d) + 64)
    assert L.idl_test_inflate(c, len(c), out, len(d) - 1) != 0        # ISIZE too small
    assert L.idl_test_inflate(c, len(c), out, len(d) + 1) != 0        # ISIZE too large
    assert L.idl_test_inflate(c, len(c) // 2, out, len(d)) != 0       # data ends early
    caught = 0
    for _ in range(300):
        bad = bytearray(c)
        bad[rng.randrange(len(bad))] ^= 1 << rng.randrange(8)
        rc = L.idl_test_inflate(bytes(bad), len(bad), out, len(d))
        assert rc != 0 or out.raw[:len(d)] != d or bytes(bad) == c
        caught += rc != 0
    assert caught > 30
    assert L.idl_test_inflate(b"\x07", 1, out, 0) != 0                # reserved block type 3


def test_crc32_by_slices(L):
    """the warp's CRC: n slices, each advanced by x^(8 * bytes behind it) mod P, XORed"""
    for d in _payloads():
        for nl in (1, 2, 7, 32):
            assert L.idl_test_crc32(d, len(d), nl) == zlib.crc32(d), (len(d), nl)
