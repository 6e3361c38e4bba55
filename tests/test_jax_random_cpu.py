"""jax.random restatement (oracle) and the product's host key arithmetic: known answers.

Threefry-2x32: the three Random123 vectors JAX's own test-suite uses.  split / uniform / normal: values printed in the
JAX documentation for PRNGKey(0) and PRNGKey(42) (quoted from memory - the reference's generator cannot run here)."""
import numpy as np
import pytest

from latent_diffusion_planning_b200 import jax_random as JR
from oracle import ldp_oracle as O

KAT = [((0x0, 0x0), (0x0, 0x0), (0x6B200159, 0x99BA4EFE)),
       ((0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF, 0xFFFFFFFF), (0x1CB996FC, 0xBB002BE7)),
       ((0x13198A2E, 0x03707344), (0x243F6A88, 0x85A308D3), (0xC4923A9C, 0x483DF7A0))]


@pytest.mark.parametrize("key,ctr,expect", KAT)
def test_threefry2x32_known_answers(key, ctr, expect):
    with np.errstate(over="ignore"):
        y0, y1 = O.threefry2x32(key[0], key[1], np.array([ctr[0]], np.uint32), np.array([ctr[1]], np.uint32))
    assert (int(y0[0]), int(y1[0])) == expect
    assert JR.threefry2x32(key[0], key[1], ctr[0], ctr[1]) == expect


def test_split_and_draws_match_documented_values():
    k = O.jax_prng_key(0)
    assert k.tolist() == [0, 0] and O.jax_prng_key(42).tolist() == [0, 42]
    assert O.jax_split(k).tolist() == [[4146024105, 967050713], [2718843009, 1272950319]]       # jax.random.split(PRNGKey(0))
    assert float(O.jax_uniform(k, 1)[0]) == pytest.approx(0.41845703, abs=1e-7)                 # jax.random.uniform(PRNGKey(0))
    assert float(O.jax_normal(k, (1,))[0]) == pytest.approx(-0.20584226, abs=2e-7)              # jax.random.normal(PRNGKey(0))
    assert O.jax_normal(k, (3,)).tolist() == pytest.approx([1.8160863, -0.48262316, 0.33988908], abs=6e-7)   # erf_inv implementations differ by a few ulp
    assert float(O.jax_normal(O.jax_prng_key(42), (1,))[0]) == pytest.approx(-0.18471177, abs=2e-7)


def test_bits_layout_odd_and_even_lengths():
    """Counters are padded to even length with a ZERO (not the next index) and split into halves."""
    k = O.jax_prng_key(5)
    with np.errstate(over="ignore"):
        b5 = O.jax_threefry_bits(k, 5)
        y0, y1 = O.threefry2x32(k[0], k[1], np.array([0, 1, 2], np.uint32), np.array([3, 4, 0], np.uint32))
    assert b5.tolist() == np.concatenate([y0, y1])[:5].tolist()
    assert O.jax_threefry_bits(k, 4).tolist() != b5[:4].tolist()


def test_host_key_arithmetic_matches_oracle():
    for seed in (0, 3, 2 ** 40 + 7):
        k = JR.prng_key(seed)
        assert np.array_equal(k, O.jax_prng_key(seed))
        for num in (1, 2, 5):
            assert np.array_equal(JR.split(k, num), O.jax_split(k, num))
        a, b = JR.sampling_keys(k, 6), O.jax_sampling_keys(k, 6)
        assert all(np.array_equal(x, y) for x, y in zip(a, b)) and a[1].shape == (6, 2)
    assert JR.is_key(np.array([1, 2], np.uint32)) and JR.is_key((1, 2)) and not JR.is_key(7) and not JR.is_key(np.int64(7))


def test_randint_and_normal_statistics():
    k = O.jax_prng_key(11)
    r = O.jax_randint(k, 20000, 0, 100)
    assert r.min() == 0 and r.max() == 99 and abs(r.mean() - 49.5) < 1.0
    z = O.jax_normal(k, (100001,))
    assert abs(z.mean()) < 0.02 and abs(z.std() - 1.0) < 0.02 and np.isfinite(z).all()


def test_update_key_threading_matches_oracle():
    k = JR.prng_key(9)
    for up, ui in ((True, True), (True, False), (False, True)):
        a, b = JR.update_keys(k, up, ui), O.jax_update_keys(k, up, ui)
        assert set(a) == set(b) and all(np.array_equal(a[n][i], b[n][i]) for n in a for i in (0, 1))
    # the IDM's keys depend on whether the planner consumed a split before it
    assert not np.array_equal(JR.update_keys(k, True, True)["idm"][0], JR.update_keys(k, False, True)["idm"][0])


def test_two_threefry_implementations_agree_on_random_blocks():
    """numpy-vectorised oracle vs the product's pure-integer host implementation, 200 random (key, counter) blocks."""
    rs = np.random.default_rng(0)
    ks, cs = rs.integers(0, 2 ** 32, (200, 2), dtype=np.uint64), rs.integers(0, 2 ** 32, (200, 2), dtype=np.uint64)
    for (k0, k1), (c0, c1) in zip(ks, cs):
        with np.errstate(over="ignore"):
            y0, y1 = O.threefry2x32(int(k0), int(k1), np.array([c0], np.uint32), np.array([c1], np.uint32))
        assert JR.threefry2x32(int(k0), int(k1), int(c0), int(c1)) == (int(y0[0]), int(y1[0]))
