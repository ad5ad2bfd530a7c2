"""Model check of RegW::insert_many (csrc/scan_reg.cuh): inserting an expansion's admitted candidates in ONE ranked merge
equals inserting them one after the other -- the sequential HnswSearchLayer loop with this library's rule for W (trim to
ef entries plus the run of entries tying with entry ef-1; a candidate is admitted while W is short or when it is strictly
nearer than entry ef-1) -- PROVIDED no candidate's distance equals another key's.  The kernel falls back to one-at-a-time
insertion whenever that cannot be guaranteed, so the claim below is exactly what its fast path relies on.  The GPU tests
compare the kernel itself with the oracle; this file pins the argument on the CPU, ties included."""
import random


def trim(w, ef):
    if len(w) > ef:
        f = w[ef - 1][0]
        keep = 0
        for e in w[ef:]:
            if e[0] != f:
                break
            keep += 1
        w = w[:ef + keep]
    return w


def sequential(w, cands, ef):
    w = list(w)
    for c in cands:
        if len(w) >= ef and not c[0] < w[ef - 1][0]:
            continue
        w.append(c)
        w.sort()
        w = trim(w, ef)
    return w


def insert_many(w, cands, ef, cap):
    """the kernel's fast path; None = it declines (the caller inserts one at a time)"""
    f = w[ef - 1][0] if len(w) >= ef else None
    adm = [c for c in cands if f is None or c[0] < f]           # the pre-filter (amask)
    if len(adm) < 2 or len(w) + len(adm) > cap:
        return None
    dists = [e[0] for e in w]
    for i, c in enumerate(adm):
        if c[0] in dists or any(j != i and o[0] == c[0] for j, o in enumerate(adm)):
            return None
    return trim(sorted(w + adm), ef)


def test_ranked_merge_equals_sequential_insertion():
    rng = random.Random(7)
    merged = declined = 0
    for _ in range(60000):
        ef = rng.choice([1, 2, 3, 5, 8, 13])
        spread = rng.choice([6, 12, 1000])                      # small spreads force ties
        w = sorted((float(rng.randint(0, spread)), i) for i in range(rng.randint(0, ef + 3)))
        w = trim(w, ef)
        cands = [(float(rng.randint(0, spread)) + rng.choice([0, 0, 0.5]), 100 + i) for i in range(rng.randint(0, 7))]
        got = insert_many(w, cands, ef, cap=ef + 8)
        if got is None:
            declined += 1
            continue
        merged += 1
        assert got == sequential(w, cands, ef), (w, cands, ef)
    assert merged > 10000 and declined > 10000


def test_look_ahead_guess_is_certain_when_nothing_sorts_before_it():
    """csrc/scan_cta.cuh: the look-ahead warp does the NEXT expansion's front half (visited filter, row staging) only
    when that expansion is certain: the second nearest unexpanded entry of W stays the nearest unexpanded one after
    this expansion's insertions unless an admitted candidate sorts before it by (distance, id).  Model: W entries are
    (distance, id, expanded); the pick takes the first unexpanded entry in (distance, id) order."""
    rng = random.Random(11)
    sure = 0
    for _ in range(40000):
        ef = rng.choice([1, 2, 4, 8, 13])
        spread = rng.choice([5, 10, 1000])
        w = trim(sorted((float(rng.randint(0, spread)), i) for i in range(rng.randint(2, ef + 3))), ef)
        expanded = {e[1] for e in w if rng.random() < 0.5}
        open_ = [e for e in w if e[1] not in expanded]
        if len(open_) < 2:
            continue
        cur, nxt = open_[0], open_[1]
        expanded.add(cur[1])
        cands = [(float(rng.randint(0, spread)) + rng.choice([0, 0.5]), 100 + i) for i in range(rng.randint(0, 7))]
        if any(c < nxt for c in cands):
            continue                                            # the kernel does nothing ahead in this case
        sure += 1
        after = sequential(w, cands, ef)
        first_open = next(e for e in after if e[1] not in expanded)
        assert first_open == nxt, (w, cands, ef, nxt, first_open)
    assert sure > 5000
