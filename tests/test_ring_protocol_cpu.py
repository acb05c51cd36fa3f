"""Executable model of the operand-ring hand-off of the two-issuer (DUAL) kernels (csrc/conv3d.cu, csrc/gemm.cu).

One TMA producer fills the stages of a ring in k-block order; every stage has a `full` and an `empty` mbarrier that are waited on
by PHASE PARITY.  A parity wait can only tell the current phase from the previous one: `try_wait.parity(P)` returns at once when
the barrier's current (incomplete) phase has the other parity.  With two consumers that is safe only if each stage always has
the same consumer, who then waits for its phases strictly in order.  The first DUAL kernels split the k-blocks by their index
in the tile over an odd ring: a stage alternated between the issuers, and an issuer a whole ring ahead of the other passed its
wait on a barrier whose previous phase had not completed - MMAs on operands that had not landed (Findings).

The model replays both assignments under random timing (TMA completions out of order, issuers at different speeds) and counts
the k-blocks consumed before their data landed."""
import random

import pytest


class Barrier:
    def __init__(self):
        self.completed = 0                       # number of completed phases

    def parity_wait_passes(self, parity):       # mbarrier.try_wait.parity
        return (self.completed & 1) != parity


def simulate(stages, n_tiles, nkb, by_ring_position, seed, steps=200000):
    rng = random.Random(seed)
    full = [Barrier() for _ in range(stages)]
    empty = [Barrier() for _ in range(stages)]
    content = [None] * stages                    # k-block whose operands the stage holds (None while a load is in flight)
    total = n_tiles * nkb
    in_flight = []                               # (stage, g): issued loads, completing in random order
    draining = []                                # (stage,): MMAs issued, their commit arrives later
    prod_g = 0
    # per issuer: the list of k-blocks (ring positions g) it consumes, in order
    mine = [[], []]
    for tile in range(n_tiles):
        cnt = tile * nkb
        for i in range(nkb):
            j = ((cnt + i) & 1) if by_ring_position else (i & 1)
            mine[j].append(cnt + i)
    pos = [0, 0]
    speed = [rng.uniform(0.2, 1.0), rng.uniform(0.2, 1.0)]
    violations = 0
    for _ in range(steps):
        if pos[0] == len(mine[0]) and pos[1] == len(mine[1]):
            break
        agent = rng.choice(("producer", "tma", "commit", 0, 1))
        if agent == "producer" and prod_g < total:
            s = prod_g % stages
            if empty[s].parity_wait_passes(((prod_g // stages) & 1) ^ 1):
                content[s] = None
                in_flight.append((s, prod_g))
                prod_g += 1
        elif agent == "tma" and in_flight:
            s, g = in_flight.pop(rng.randrange(len(in_flight)))        # completion order is not issue order
            content[s] = g
            full[s].completed += 1
        elif agent == "commit" and draining:
            s = draining.pop(0)                                        # an issuer's commits complete in its issue order
            empty[s].completed += 1
        elif agent in (0, 1) and pos[agent] < len(mine[agent]) and rng.random() < speed[agent]:
            g = mine[agent][pos[agent]]
            s = g % stages
            if full[s].parity_wait_passes((g // stages) & 1):
                if content[s] != g:
                    violations += 1                                    # operands of another k-block / not landed yet
                draining.append(s)
                pos[agent] += 1
    done = pos[0] == len(mine[0]) and pos[1] == len(mine[1])
    return violations, done


def test_index_in_tile_split_over_an_odd_ring_can_consume_stale_operands():
    bad = sum(simulate(5, 6, 27, by_ring_position=False, seed=s)[0] > 0 for s in range(200))
    assert bad > 0, "the model should reproduce the aliasing of the first DUAL kernels"


@pytest.mark.parametrize("stages,nkb", [(4, 27), (6, 27), (4, 16), (6, 64), (4, 5)])
def test_ring_position_split_over_an_even_ring_never_does(stages, nkb):
    for s in range(300):
        violations, done = simulate(stages, 6, nkb, by_ring_position=True, seed=s)
        assert violations == 0 and done, (s, violations, done)
