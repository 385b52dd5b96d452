"""bfq_range_lookup (batched TenantRangeLookupCache.lookup) against the oracle's literal restatement: the
reference's own vectors and random topics x random candidate chains (bounds drawn from random filters, incl. wildcards, empty
levels, '$' topics, bounds of other tenants, missing Facts)."""
import random

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu


def test_reference_vectors_through_the_kernel():
    from bifromq_b200 import dist as D
    from golden.range_lookup_vectors import T, VECTORS
    topics = [v[0] for v in VECTORS]
    # one tenant entry per vector (each vector has its own candidate chain), all named T
    got = D.range_lookup([T] * len(VECTORS), topics, np.arange(len(VECTORS), dtype=np.int32), [v[1] for v in VECTORS])
    assert got == [v[2] for v in VECTORS]


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_chains_vs_oracle(seed):
    from bifromq_b200 import dist as D
    rng = random.Random(seed)
    vocab = ["a", "b", "c", "dd", "", "!", "$s", "zz", "m", "+x", "#y", "~"]
    tenants = ["tA", "tB", "t"]

    def levels(n, wild):
        out = []
        for i in range(n):
            r = rng.random()
            if wild and r < 0.2:
                out.append("+")
            elif wild and r < 0.3 and i == n - 1:
                out.append("#")
            else:
                out.append(rng.choice(vocab))
        return out
    cands = []
    for t in tenants:
        bounds = sorted({tuple(levels(rng.randint(0, 4), True)) for _ in range(rng.randint(0, 9))})
        chain = []
        i = 0
        while i + 1 < len(bounds):
            r = rng.random()
            owner = t if rng.random() < 0.9 else rng.choice(tenants)      # a stray bound of another tenant
            if r < 0.12:
                chain.append(None)
            elif r < 0.18:
                chain.append(([owner] + list(bounds[i]), None))
            elif r < 0.24:
                chain.append((None, [owner] + list(bounds[i + 1])))
            else:
                chain.append(([owner] + list(bounds[i]), [owner] + list(bounds[i + 1])))
            i += rng.choice([1, 1, 2])
        cands.append(chain)
    topics, tt = [], []
    for _ in range(3000):
        topics.append("/".join(levels(rng.randint(1, 5), False)))
        tt.append(rng.randrange(len(tenants)))
    got = D.range_lookup(tenants, topics, np.asarray(tt, np.int32), cands)
    kept = 0
    for i, topic in enumerate(topics):
        want = O.range_lookup(tenants[tt[i]], topic, cands[tt[i]])
        assert got[i] == want, (topic, tenants[tt[i]], cands[tt[i]], got[i], want)
        kept += len(want)
    assert kept > 100
