"""Oracle self-checks: generator and samplers. The Philox vectors are the Random123 known-answer
tests (kat_vectors: philox4x32 10 rounds); they are the one externally pinned piece of this path."""
import numpy as np

SEED = 0x5EED0001


def test_philox_known_answers(oracle):
    kat = [
        ([0, 0, 0, 0], [0, 0], [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
        ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
        ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0],
         [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]),
    ]
    for ctr, key, want in kat:
        assert [int(x) for x in oracle.philox4x32_10(ctr, key)] == want


def test_stream_layout(oracle):
    # draw i = word (i & 3) of Philox(counter = (pixel, sample, i >> 2, 0), key = seed)
    seed, pixel, sample = 0x0123456789ABCDEF, 77, 5
    draws = oracle.rng_draws(seed, pixel, sample, 10)
    key = [seed & 0xFFFFFFFF, seed >> 32]
    for i in range(10):
        assert int(draws[i]) == int(oracle.philox4x32_10([pixel, sample, i >> 2, 0], key)[i & 3])
    # different samples / pixels are different streams
    assert not np.array_equal(draws, oracle.rng_draws(seed, pixel, sample + 1, 10))
    assert not np.array_equal(draws, oracle.rng_draws(seed, pixel + 1, sample, 10))


def test_unit_sphere_distribution(oracle):
    v = oracle.unit_sphere(SEED, 1, 2, 200000).astype(np.float64)
    assert np.allclose(np.linalg.norm(v, axis=1), 1.0, atol=2e-6)
    assert np.all(np.abs(v.mean(0)) < 0.01)                    # E[v] = 0
    assert np.allclose((v ** 2).mean(0), 1.0 / 3.0, atol=0.01)  # E[v_i^2] = 1/3
    # z is uniform on [-1, 1]
    hist, _ = np.histogram(v[:, 2], bins=10, range=(-1, 1))
    assert np.all(np.abs(hist / len(v) - 0.1) < 0.01)


def test_unit_disc_distribution(oracle):
    v = oracle.unit_disc(SEED, 3, 4, 200000).astype(np.float64)
    r2 = (v ** 2).sum(1)
    assert np.all(r2 <= 1.0)
    assert abs(r2.mean() - 0.5) < 0.01  # E[r^2] = 1/2 for a uniform disc
    assert np.all(np.abs(v.mean(0)) < 0.01)
