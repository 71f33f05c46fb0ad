"""Model check (CPU, no GPU) of the exchange protocol that is fused into the sharded GEMV (csrc/gemv.cuh, PEER = true).

The kernel's rules, restated as a little asynchronous machine and run under random, deliberately unfair interleavings (some ranks 10-100x faster) for world sizes the
round could not put on real GPUs (4, 8) as well as 2:

  * every rank owns a buffer of slots  [set 0|1][source rank][element] -> (tag, value), all tags 0 at start;
  * launch number s (s = 1, 2, ...) uses set s & 1.  Each (rank, element) thread first stores (s, its partial) into
    slot [s & 1][me][element] of EVERY rank, then polls ITS OWN buffer: for source 0 .. W-1 in order it waits until the tag
    of slot [s & 1][source][element] equals s and adds the value (deterministic order);
  * the threads of one launch run independently (no barrier between push and poll), but launch s + 1 of a rank starts only
    after all threads of its launch s have finished (stream order / griddepcontrol.wait);
  * linears of different sizes follow each other, so an element index is not used by every launch.

Checked: no deadlock, every rank's result equals the sum over ranks in rank order for every launch, and a slot is never
overwritten before its reader has consumed it (the two-set alternation is enough: a rank can only be pushing launch s + 2
after it has received every rank's launch s + 1 words, which are pushed after that rank finished launch s).
A deliberately broken variant (ONE set) must be caught by the same checker.
"""
import random

import pytest


class Rank:
    def __init__(self, world, max_elems, n_sets):
        self.buf = [[[(0, None)] * max_elems for _ in range(world)] for _ in range(n_sets)]
        self.launch = 1          # launch currently executing (1-based)
        self.threads = None      # per element: [phase ('push'|'poll'), next peer index, accumulator list]
        self.results = {}        # launch -> list of per-element sums (as tuples of the addends, to check the order)


def partial(rank, launch, elem):
    return (rank, launch, elem)  # a unique token instead of a float: sums become exact, order-checkable tuples


def run(world, sizes, n_sets, seed):
    rng = random.Random(seed)
    speed = [rng.choice((1, 1, 10, 100)) for _ in range(world)]  # some ranks run far ahead of others
    max_elems = max(sizes)
    ranks = [Rank(world, max_elems, n_sets) for _ in range(world)]
    consumed = {}  # (dst, set, src, elem) -> True when the last value written there has been read by dst
    n_launches = len(sizes)

    def start(r):
        n = sizes[r.launch - 1]
        r.threads = [["push", 0, []] for _ in range(n)]

    for r in ranks:
        start(r)
    steps = 0
    while any(r.launch <= n_launches for r in ranks):
        steps += 1
        assert steps < 2_000_000, "no progress: deadlock"
        runnable = []
        for ri, r in enumerate(ranks):
            if r.launch > n_launches:
                continue
            s, st = r.launch, r.launch % n_sets
            for e, th in enumerate(r.threads):
                if th[0] == "push":
                    runnable.append((ri, e))
                elif th[0] == "poll":
                    tag, _ = r.buf[st][th[1]][e]
                    if tag == s:
                        runnable.append((ri, e))
        assert runnable, "deadlock: every live thread waits for a tag that nobody can write"
        ri, e = rng.choices(runnable, weights=[speed[a] for a, _ in runnable])[0]
        r = ranks[ri]
        s, st = r.launch, r.launch % n_sets
        th = r.threads[e]
        if th[0] == "push":
            dst = th[1]
            key = (dst, st, ri, e)
            old_tag, _ = ranks[dst].buf[st][ri][e]
            # overwriting a value that its reader has not consumed yet would lose data
            assert old_tag == 0 or consumed.get(key, False), f"rank {ri} launch {s} overwrites an unread slot of rank {dst}"
            ranks[dst].buf[st][ri][e] = (s, partial(ri, s, e))
            consumed[key] = False
            th[1] += 1
            if th[1] == world:
                th[0], th[1] = "poll", 0
        else:
            src = th[1]
            tag, val = r.buf[st][src][e]
            assert tag == s and val == partial(src, s, e), "read a value of another launch"
            consumed[(ri, st, src, e)] = True
            th[2].append(val)
            th[1] += 1
            if th[1] == world:
                th[0] = "done"
        if all(t[0] == "done" for t in r.threads):
            r.results[s] = [tuple(t[2]) for t in r.threads]
            r.launch += 1
            if r.launch <= n_launches:
                start(r)
    for ri, r in enumerate(ranks):
        for s in range(1, n_launches + 1):
            want = [tuple(partial(src, s, e) for src in range(world)) for e in range(sizes[s - 1])]
            assert r.results[s] == want, (ri, s)
    return steps


@pytest.mark.parametrize("world", [2, 4, 8])
def test_two_set_tagged_exchange_is_safe_under_random_interleavings(world):
    sizes = [3, 5, 2, 5, 5, 1, 4, 3]  # launches of different widths, like q/k/v, o, gate/up, down of consecutive layers
    for seed in range(12):
        run(world, sizes, n_sets=2, seed=1000 * world + seed)


def test_checker_catches_a_single_set_protocol():
    """With ONE set a fast rank's launch s + 1 push can overwrite a slot whose launch-s value a slow rank has not read."""
    caught = 0
    for seed in range(40):
        try:
            run(4, [3, 3, 3, 3, 3, 3], n_sets=1, seed=seed)
        except AssertionError:
            caught += 1
    assert caught > 0
