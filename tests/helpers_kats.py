"""Known-answer vectors copied from the reference's Rust unit tests + a seeded label generator."""
import numpy as np

# ---- majority_voting: src/smooth/utils.rs:103-137 --------------------------------------------
MV_KATS = [
    ([1, 0, 0, 1, 1, 0, 1, 0, 0, 0, 1], 3, [1, 0, 0, 1, 1, 1, 0, 0, 0, 0, 0]),
    ([1, 0, 0, 1, 1, 0, 1, 1, 1, 0, 1], 3, [1, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1]),
    ([], 3, []),
    ([1, 0, 0, 1, 1, 0, 1, 0, 0, 0], 1, [1, 0, 0, 1, 1, 0, 1, 0, 0, 0]),
]
# ---- get_label_region: src/utils.rs:770-803 (+ derived start==0 quirk cases) ------
REGION_KATS = [
    ([], []),
    ([0, 0, 0, 0], []),
    ([0, 1, 0, 0, 0], [(1, 2)]),
    ([0, 1, 1, 0, 1, 1, 0], [(1, 3), (4, 6)]),
    ([0, 1, 1, 0, 1, 1], [(1, 3), (4, 6)]),
    ([1, 1, 1, 0], [(1, 3)]),
    ([1, 0, 0], []),
    ([1], []),
]



def _random_labels(rng, n):
    lab = (rng.random(n) < 0.03).astype(np.int8)
    for _ in range(rng.integers(0, 5)):
        if n < 4:
            break
        s = int(rng.integers(0, n))
        e = min(n, s + int(rng.integers(5, 120)))
        lab[s:e] = (rng.random(e - s) > 0.08)
    if rng.random() < 0.3:
        lab[: int(rng.integers(1, 40))] = 1
    if rng.random() < 0.5:
        lab[n - int(rng.integers(1, 80)):] = 1
    return lab


