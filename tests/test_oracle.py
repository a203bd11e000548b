"""CPU tests pinning the oracle (oracle/) against the reference's golden vectors.

Golden values: tests/golden/single_step.json, produced by the reference's verbatim test_cpu
(tests/golden/make_golden.py); the first nine cases are the table of SURVEY.md section 8(c)."""
import hashlib
import json
import os

import numpy as np
import pytest

import oracle

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "single_step.json")))


def _interior(shape, dims, a):
    halo = oracle.HALO[oracle.dim_of(shape)]
    return a[tuple(slice(h, h + x) for h, x in zip(halo, dims))]


@pytest.mark.parametrize("g", GOLDEN, ids=lambda g: f"{g['shape']}-{'x'.join(map(str, g['dims']))}")
def test_single_step_matches_reference_golden(g):
    shape, dims = g["shape"], tuple(g["dims"])
    a = oracle.fill_rand(shape, dims)
    assert a.ravel()[:4].tolist() == g["in_first4"]  # unseeded glibc rand() fill
    out = oracle.step(shape, a, oracle.reference_params(shape))
    halo = oracle.HALO[oracle.dim_of(shape)]
    assert out[tuple(halo)] == g["first"]
    assert out[tuple(h + x // 2 for h, x in zip(halo, dims))] == g["centre"]
    assert out[tuple(h + x - 1 for h, x in zip(halo, dims))] == g["last"]
    inner = np.ascontiguousarray(_interior(shape, dims, out))
    assert float(inner.sum()) == g["interior_sum"]
    assert hashlib.sha256(inner.tobytes()).hexdigest() == g["interior_sha256"]


def test_survey_table_values():
    """The literal numbers of SURVEY.md section 8(c), independent of the json fixture."""
    a = oracle.fill_rand("box2d1r", (1024, 1024))
    o = oracle.step("box2d1r", a, oracle.reference_params("box2d1r"))
    assert (o[4, 4], o[516, 516], o[1027, 1027]) == (12293.0, 11314.0, 12409.0)
    assert int(o[4:1028, 4:1028].sum()) == 12042023154
    a = oracle.fill_rand("star3d1r", (16, 16, 64))
    o = oracle.step("star3d1r", a, oracle.reference_params("star3d1r"))
    assert (o[1, 2, 4], o[9, 10, 36], o[16, 17, 67]) == (361.0, 341.0, 538.0)


@pytest.mark.skipif(not oracle.ref_available("cpu", 2), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("shape,dims", [("1d2r", (4096,)), ("box2d3r", (96, 128)), ("star2d1r", (64, 64)),
                                        ("star2d3r", (37, 41)), ("box3d1r", (8, 16, 64)), ("star3d1r", (6, 7, 9))])
def test_oracle_equals_reference_test_cpu(shape, dims):
    """Bit-exact against the reference's own test_cpu, also with non-integer data and weights."""
    rng = np.random.default_rng(1)
    d = oracle.dim_of(shape)
    a = rng.standard_normal(oracle.padded_shape(shape, dims))
    for params in (oracle.reference_params(shape), rng.standard_normal({1: 9, 2: 49, 3: 27}[d])):
        assert np.array_equal(oracle.ref_cpu_step(d, a, params), oracle.step(d, a, params))


@pytest.mark.parametrize("shape,dims", [("1d1r", (256,)), ("box2d1r", (32, 64)), ("box3d1r", (8, 8, 64))])
def test_ping_pong_semantics(shape, dims):
    """S2/S3: halo of the result is the input halo for even `times`, zero for odd `times`; launch i
    reads what launch i-1 wrote plus that alternating halo; 1-D leaves out[cols-1] untouched."""
    a = oracle.fill_rand(shape, dims)
    p = oracle.reference_params(shape)
    d = oracle.dim_of(shape)
    halo = oracle.HALO[d]
    inner = tuple(slice(h, h + x) for h, x in zip(halo, dims))
    mask = np.ones(a.shape, dtype=bool)
    mask[inner] = False
    # manual ping-pong with explicit buffers
    buf = [a.copy(), np.zeros_like(a)]
    for t in range(1, 5):
        src, dst = buf[(t - 1) % 2], buf[t % 2]
        dst[inner] = oracle.step(shape, src, p)[inner]
        out = np.full_like(a, -7.0)
        oracle.run(shape, a, p, t, out=out)
        ref = buf[t % 2]
        if d == 1:
            assert out[-1] == -7.0  # src/1d/gpu_1r.cu:134 copies cols-1 doubles
            assert np.array_equal(out[:-1], ref[:-1])
        else:
            assert np.array_equal(out, ref)
        halo_vals = out[mask] if d > 1 else out[:-1][mask[:-1]]
        expect = (a[mask] if d > 1 else a[:-1][mask[:-1]]) if t % 2 == 0 else 0.0
        assert np.array_equal(halo_vals, expect if t % 2 == 0 else np.zeros_like(halo_vals))


def test_effective_params_of_reference_tables_are_the_tables():
    for s in oracle.ALL_SHAPES:
        assert np.array_equal(oracle.effective_params(s), oracle.reference_params(s))


def test_reference_peel_is_rank3_and_drops_nothing_on_shipped_table():
    u, v, c = oracle.reference_peel_box2d(oracle.reference_params("box2d1r"))
    assert c == 0.0
    assert u[0].tolist() == [1, 2, 3, 4, 3, 2, 1] and v[2].tolist() == [0, 0, 1, 3, 1, 0, 0]


def test_reference_quirks():
    rng = np.random.default_rng(3)
    w = rng.standard_normal(49)
    assert np.array_equal(oracle.effective_params("star2d1r", w), oracle.reference_params("star2d1r"))
    w3 = rng.standard_normal(27)
    assert np.array_equal(oracle.effective_params("star3d1r", w3), oracle.reference_params("star3d1r"))
    e = oracle.effective_params("box3d1r", w3).reshape(3, 3, 3)
    assert np.array_equal(e, np.broadcast_to(w3[:3], (3, 3, 3)))
