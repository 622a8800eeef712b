"""The barrier protocol of the CTA-pair sweep kernel under random interleavings (tests/sweep_protocol_model.py; CPU only): progress and
resource safety for arbitrary item lists -- persistent pairs walking many items, items without column tiles, hard-negative-only items the
O-CTA sits out -- and proof that the model has teeth: three deliberately broken variants of the protocol are caught."""
import random

import pytest

from sweep_protocol_model import Deadlock, Hazard, Sim, random_items


def test_one_pair_per_item_and_persistent_lists_make_progress_and_keep_their_resources():
    rng = random.Random(1)
    lists = [[(5, False)], [(1, False)], [(2, True)], [(0, False), (3, False)], [(1, False), (1, False), (1, False), (2, False)],
             [(4, True), (4, False), (0, True), (1, True), (7, False)]]
    lists += [random_items(rng, rng.randint(1, 8)) for _ in range(150)]
    for items in lists:
        for seed in range(5):
            Sim(items, seed).run()


@pytest.mark.parametrize('variant,what', [
    ('no_o_empty', 'the next item\'s first GEMM-2 does not wait for the O write-out'),
    ('release_last_tile_by_commit', 'the last tile\'s P~ buffer is released by the MMA commit although the write-out stages O through it'),
    ('p_reload_without_barrier', 'a warpgroup stores the next probe tile without the end-of-item barrier'),
])
def test_broken_variants_are_caught(variant, what):
    rng = random.Random(2)
    caught = 0
    for _ in range(60):
        items = [(rng.randint(2, 6), False) for _ in range(rng.randint(2, 5))]
        for seed in range(3):
            try:
                Sim(items, seed, variant=variant).run()
            except (Hazard, Deadlock, AssertionError):
                caught += 1
    assert caught > 0, what
