"""Oracle vs the known-answer values of SURVEY.md Appendix E and structural invariants.
(The reference ships no tests or golden vectors; the oracle is pinned against its shader sources by
tests/test_reference_glsl.py, and additionally by these KATs and the committed golden hashes in tests/golden/.)"""
import ctypes as C

import numpy as np

import oracle_py as O


def test_shuffle_table_libstdcxx():
    # VPLProcessor.cpp:58-62 std::shuffle(iota(1024), mt19937(1979982))
    t = O.shuffle_table(1024)
    assert list(t[:16]) == [652, 300, 15, 926, 310, 942, 86, 391, 947, 848, 205, 796, 948, 193, 992, 787]
    assert int(sum((i + 1) * int(v) for i, v in enumerate(t))) == 269800309
    assert sorted(t.tolist()) == list(range(1024))


def test_jenkins_hash_and_random():
    lib = O.load()
    assert lib.orc_hash(1) == 0x124EA49D
    assert lib.orc_hash(0x3F800000) == 0xF2496DC0
    third = np.float32(1.0) / np.float32(3.0)
    r = lib.orc_random3(third, third, third)
    assert abs(r - 0.429122567) < 1e-8 and int(r * 1024) == 439
    t2 = np.float32(third + np.float32(5 / 4096.0))
    r = lib.orc_random3(t2, t2, t2)
    assert abs(r - 0.626130581) < 1e-8 and int(r * 1024) == 641


def test_pack3sn_asymmetric_bias():
    lib = O.load()
    assert lib.orc_pack3SN(0, 1, 0) == 0x200FFE00
    assert lib.orc_pack3SN(0.6, 0, -0.8) == 0x06780332
    o = np.empty(3, np.float32)
    lib.orc_unpack3SN(0x200FFE00, o.ctypes.data)
    np.testing.assert_allclose(o, [0.000977516174, 1.0, 0.0], rtol=0, atol=1e-9)
    lib.orc_unpack3SN(0x06780332, o.ctypes.data)
    np.testing.assert_allclose(o, [0.599218011, 0.000977516174, -0.801175296], rtol=0, atol=1e-8)


def test_ism_grid_constants():
    assert [O.ism_indices_1d(n) for n in (64, 1024, 4096, 16384)] == [8, 32, 64, 128]


def test_cluster_slices_table_matches_log2_formula():
    lib = O.load()
    # Appendix E values (zFar = 50); 50 -> 16 in the literal formula, clamped to 15 here (Q7)
    for d, s in ((0.05, 0), (0.5, 0), (1, 0), (5, 4), (20, 11), (49.9, 15)):
        assert lib.orc_slice_log2(-d, 50.0) == s
        assert lib.orc_slice_table(-d, 50.0) == s
    assert lib.orc_slice_log2(-50.0, 50.0) == 16 and lib.orc_slice_table(-50.0, 50.0) == 15
    # the table form equals the literal log2 form away from the boundaries (1e-5 relative guard band)
    B = O.slice_boundaries(50.0).astype(np.float64)
    rng = np.random.default_rng(1)
    ds = np.exp(rng.uniform(np.log(0.05), np.log(49.9), 20000))
    checked = 0
    for d in ds:
        if np.min(np.abs(d / B[1:16] - 1.0)) < 1e-5:
            continue
        checked += 1
        assert lib.orc_slice_log2(-float(d), 50.0) == lib.orc_slice_table(-float(d), 50.0)
    assert checked > 19000


def test_empty_atlas_texel_decode_q1():
    assert np.float32(0xFFFFFFFF >> 8) / np.float32(16777216.0) == np.float32(0.99999994)
    assert np.float32(0xFFFFFFFF >> 8) / np.float32(16777216.0) != np.float32(1.0)


def test_half_packing_matches_numpy_rne():
    lib = O.load()
    rng = np.random.default_rng(0)
    xs = np.concatenate([rng.standard_normal(4000).astype(np.float32) * s for s in (1e-8, 6e-8, 1e-5, 1e-3, 1, 100, 70000)])
    xs = np.concatenate([xs, np.array([0.0, -0.0, 65504.0, 65520.0, 5.9604645e-08, 2.9802322e-08, np.inf, -np.inf], np.float32)])
    for v in xs:
        got = lib.orc_packHalf2x16(float(v), 0.0) & 0xFFFF
        assert got == int(np.float16(v).view(np.uint16)), v
    o = np.empty(2, np.float32)
    for h in list(range(0, 0x7C01, 97)) + [0x8000, 0xFBFF, 0x0001, 0x03FF, 0x0400]:
        lib.orc_unpackHalf2x16(h, o.ctypes.data)
        assert o[0] == np.float32(np.uint16(h).view(np.float16))


def test_packunorm4x8():
    lib = O.load()
    assert lib.orc_packUnorm4x8(0.0, 1.0, 0.5, 2.0) == (0 | (255 << 8) | (128 << 16) | (255 << 24))
    assert lib.orc_packUnorm4x8(-1.0, 0.25, 0.75, 0.1) == (0 | (64 << 8) | (191 << 16) | (26 << 24))


def test_mat4_inverse_roundtrip():
    rng = np.random.default_rng(3)
    for _ in range(20):
        m = rng.standard_normal(16).astype(np.float32) + np.eye(4, dtype=np.float32).ravel() * 3
        inv = O.mat4_inverse(m)
        prod = O.mat4_mul(m, inv).reshape(4, 4)
        np.testing.assert_allclose(prod, np.eye(4), atol=2e-5)


def test_paraboloid_projection_stays_in_its_tile():
    lib = O.load()
    rng = np.random.default_rng(5)
    out = np.empty(3, np.float32)
    S = 32
    for _ in range(2000):
        n = rng.standard_normal(3)
        n = (n / np.linalg.norm(n)).astype(np.float32)
        rel = rng.standard_normal(3).astype(np.float32) * 5
        if float(np.dot(n, rel)) < 0.05 * np.linalg.norm(rel):
            continue
        idx = int(rng.integers(0, 1024))
        lib.orc_paraboloid_project(rel.ctypes.data, n.ctypes.data, 50.0, float(idx), S, 1, out.ctypes.data)
        tx, ty = idx % S, idx // S
        assert tx / S - 1e-6 <= out[0] <= (tx + 1) / S + 1e-6
        assert ty / S - 1e-6 <= out[1] <= (ty + 1) / S + 1e-6
        assert out[2] >= 0  # in front of the VPL -> positive depth with preserveSign


def test_r11g11b10_packing_kats():
    """GL_R11F_G11F_B10F (GIStage.cpp:470): 5-bit exponent bias 15, 6/6/5 mantissa bits, r in the low bits"""
    pk = lambda r, g, b: int(O.pack_r11g11b10(np.array([[r, g, b]], np.float32))[0])
    assert pk(1.0, 0.0, 0.0) == 0x3C0 and pk(0.0, 1.0, 0.0) == 0x3C0 << 11 and pk(0.0, 0.0, 1.0) == 0x1E0 << 22
    assert pk(0.5, 2.0, 0.25) == (0x380 | (0x400 << 11) | (0x1A0 << 22))
    assert pk(65024.0, 1e9, 64512.0) == (0x7BF | (0x7BF << 11) | (0x3DF << 22))      # largest finite / clamp
    assert pk(-1.0, -0.0, -np.inf) == 0
    assert pk(np.inf, np.nan, 0.0) == (0x7C0 | (0x7C1 << 11))
    assert pk(2.0 ** -20, 2.0 ** -21, 2.0 ** -19) == (1 | (0 << 11) | (1 << 22))      # smallest subnormals, tie to even -> 0
    assert pk(1.0 + 1.0 / 128, 1.0 + 3.0 / 128, 1.0 + 1.0 / 64) == (0x3C0 | (0x3C2 << 11) | (0x1E0 << 22))   # ties to even
    rng = np.random.default_rng(4)
    x = np.exp(rng.uniform(np.log(1e-6), np.log(6e4), (20000, 3))).astype(np.float32)
    back = O.unpack_r11g11b10(O.pack_r11g11b10(x))
    rel = np.abs(back - x) / x
    assert rel[:, :2][x[:, :2] > 7e-5].max() <= 2.0 ** -7 and rel[:, 2][x[:, 2] > 7e-5].max() <= 2.0 ** -6
    assert np.array_equal(O.pack_r11g11b10(back), O.pack_r11g11b10(x))                # idempotent


def test_composite_ambient_kats():
    """n4 ambient term: deferredshading.frag:69-80 + tonemap + toSRGB on hand-computed values"""
    gi = np.array([[[0.0, 1.0, 3.0], [0.5, 0.5, 0.5]]], np.float32)
    diffuse = np.array([[0x00FFFFFF, 0x000080FF]], np.uint32)          # (1,1,1) and (1, 128/255, 0)
    f, q = O.composite_ambient(gi, diffuse, None, 1.0)
    expect0 = [0.0, (1.0 / 2.0) ** (1 / 2.2), (3.0 / 4.0) ** (1 / 2.2)]
    np.testing.assert_allclose(f[0, 0], expect0, rtol=1e-6)
    d1 = (128 / 255.0) ** 2.2
    expect1 = [(0.5 / 1.5) ** (1 / 2.2), (0.5 * d1 / (1 + 0.5 * d1)) ** (1 / 2.2), 0.0]
    np.testing.assert_allclose(f[0, 1], expect1, rtol=1e-5)
    assert q[0, 0] == (0xFF000000 | (round(expect0[2] * 255) << 16) | (round(expect0[1] * 255) << 8))
    # occlusion scales the ambient term; exposure scales before the Reinhard curve
    f2, _ = O.composite_ambient(gi, diffuse, np.full((1, 2), 0.5, np.float32), 2.0)
    np.testing.assert_allclose(f2, f, rtol=1e-6)
