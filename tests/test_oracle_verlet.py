"""CPU: the skin-list argument of data/verlet.py restated on the oracle -- filtering candidates built with
cutoff + skin reproduces the full neighbour search exactly while every atom stays within skin/2 of the coordinates the
candidates were built on, and stops doing so beyond that (the rebuild trigger is necessary, not a heuristic)."""
import numpy as np

from oracle import m3gnet_oracle as O


def _same(a, b):
    return all(np.array_equal(x, y) for x, y in zip(a, b))


def test_filtered_candidates_equal_full_search_within_half_skin():
    rng = np.random.default_rng(0)
    cases = [O.fcc_supercell(2, jitter=0.05, seed=1), O.mpf_like_structure(4),
             (np.array([[2.6, 0, 0], [0.3, 2.7, 0], [0.1, -0.2, 2.9]]), np.array([[0.1, 0.2, 0.3]]), np.array([29]))]
    for lat, cart, _ in cases:
        for r, skin in ((5.0, 0.5), (4.0, 1.0)):
            cs, cd, ci, _ = O.neighbor_list_bruteforce(lat, cart, r + skin)
            for _ in range(4):
                step = rng.normal(size=cart.shape)
                step *= (0.4999 * skin * rng.uniform(0.2, 1.0, size=(len(cart), 1))) / np.linalg.norm(step, axis=1,
                                                                                                    keepdims=True)
                new = cart + step                                     # every atom moves by < skin/2
                want = O.neighbor_list_bruteforce(lat, new, r)
                got = O.verlet_filter(lat, new, cs, cd, ci, r)
                assert _same(got[:3], want[:3]) and np.array_equal(got[3], want[3])


def test_filter_misses_bonds_beyond_the_skin():
    lat, cart, _ = O.fcc_supercell(2, jitter=0.0, seed=0)
    r, skin = 3.0, 0.2                                                # nearest neighbours at 2.556, next shell at 3.615
    cs, cd, ci, _ = O.neighbor_list_bruteforce(lat, cart, r + skin)
    new = cart.copy()
    new[0] += np.array([0.7, 0.0, 0.0])       # >> skin/2: the second-shell atom at (a, 0, 0) comes inside r (2.915 A)
    want = O.neighbor_list_bruteforce(lat, new, r)
    got = O.verlet_filter(lat, new, cs, cd, ci, r)
    assert len(want[0]) > len(got[0])
