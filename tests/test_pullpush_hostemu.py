"""The arithmetic the CUDA pull-push kernel executes (many-lights-gi_b200/csrc/pullpush_core.cuh, the per-thread phases of
pullpush_kernel) compiled for the HOST and compared with the oracle bit for bit -- the kernel's logic is checked without a GPU;
the -m gpu tests check the same phases as compiled for sm_100a. tests/hostemu is test infrastructure: it is never linked into
libmlgi.so."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_py as O

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostemu")


@pytest.fixture(scope="module")
def emu():
    subprocess.check_call(["make", "-s", "-C", HERE])
    lib = C.CDLL(os.path.join(HERE, "libpp_emu.so"))
    lib.pp_emu.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_void_p]

    def run(atlas, z_far=50.0):
        A = atlas.shape[0]
        out = np.empty((A, A), np.uint16)
        lib.pp_emu(np.ascontiguousarray(atlas).ctypes.data, A // 64, z_far, 64.0, out.ctypes.data)
        return out
    return run


def test_phases_match_oracle_on_splatted_atlases(emu, case_small, case_mid):
    for case in (case_small, case_mid):
        atlas = case.get("atlas")
        assert np.array_equal(emu(atlas), case.get("ism"))


@pytest.mark.parametrize("N", [64, 1024])
def test_phases_match_oracle_on_synthetic_atlases(emu, N):
    """empty / dense / single-texel / sparse atlases, zero depth (infinite pixel radius), zero depth AND zero radius (NaN
    radius), and uniformly random 32-bit texels"""
    A = 64 * O.ism_indices_1d(N)
    rng = np.random.default_rng(11)
    empty = np.full((A, A), 0xFFFFFFFF, np.uint32)
    dense = ((rng.integers(1 << 18, 1 << 24, (A, A), dtype=np.uint64) << 8) | rng.integers(0, 126, (A, A), dtype=np.uint64)).astype(np.uint32)
    single = empty.copy()
    single[32::64, 32::64] = (np.uint32(3_000_000) << 8) | np.uint32(40)
    sparse = empty.copy()
    m = rng.random((A, A)) < 0.05
    sparse[m] = dense[m]
    zero_depth = sparse.copy()
    zero_depth[5::64, 7::64] = np.uint32(9)
    zero_depth[9::64, 3::64] = np.uint32(0)
    anybits = ((rng.integers(0, 1 << 24, (A, A), dtype=np.uint64) << 8) | rng.integers(0, 256, (A, A), dtype=np.uint64)).astype(np.uint32)
    for name, atl in (("empty", empty), ("dense", dense), ("single", single), ("sparse", sparse), ("zero_depth", zero_depth), ("anybits", anybits)):
        assert np.array_equal(emu(atl), O.ism_pullpush(atl, N, 50.0)), name


def test_other_z_far(emu, case_small):
    atlas = case_small.get("atlas")
    assert np.array_equal(emu(atlas, 20.0), O.ism_pullpush(atlas, case_small.N, 20.0))


def test_fast_ufloat_decode_equals_reference_decode_for_every_code(emu):
    """common.cuh unpack_ufloat_fast (branch-free decode of the R11F_G11F_B10F channels used by the blur and de-interleave
    kernels) against unpack_ufloat_bits, all 2048 + 1024 codes including subnormals, inf and NaN"""
    lib = C.CDLL(os.path.join(HERE, "libpp_emu.so"))
    lib.ufloat_fast_mismatches.restype = C.c_int
    assert lib.ufloat_fast_mismatches() == 0
