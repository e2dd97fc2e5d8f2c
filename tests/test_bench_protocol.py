"""The bench loop rotates over NSETS world sets and a ring of RING pre-generated action tensors.  With the same index for both
(`sets[i % NSETS].step(actions[i % RING])`, RING == NSETS) every set is fed THE SAME tensor at every one of its steps: agents that
repeat one action for ever, a much cheaper rollout than the random one the bench claims (rounds 1 and early 2 measured exactly that,
DESIGN.md section 6).  The index used now must give every set a different tensor at each of its steps."""
import ast
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _index(i, nsets, ring):
    return (i // nsets + i % nsets) % ring  # set k = i % nsets, at its j-th visit (j = i // nsets): tensor (j + k) % ring


def test_every_world_set_sees_every_action_tensor():
    for nsets, ring in ((16, 16), (8, 8), (2, 8), (1, 16)):
        seen = {k: [] for k in range(nsets)}
        for i in range(nsets * ring * 2):
            seen[i % nsets].append(_index(i, nsets, ring))
        for k, idx in seen.items():
            assert len(set(idx[:ring])) == ring, (nsets, ring, k, idx[:ring])      # all tensors within one ring of visits
            assert all(a != b for a, b in zip(idx, idx[1:])) or ring == 1            # never the same tensor twice in a row


def test_bench_uses_that_index():
    src = open(os.path.join(ROOT, "bench.py")).read()
    ast.parse(src)
    assert "actions[(i // NSETS + i % NSETS) % RING]" in src
    assert "sets[i % NSETS].step(actions[i % RING])" not in src
