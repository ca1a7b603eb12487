"""CPU model of the one place where csrc/group.cu departs from the reference's control flow: the detections of
a joint step are placed in one parallel move instead of one by one (grouping.py:131-143), unless a dict-key
clash makes their order observable.  The model replays both on random steps, including persons beyond the
first M and planted key clashes, and requires identical person tables."""
import copy

import numpy as np


def place_one_by_one(state, dets, star, ok, k, M):
    """grouping.py:131-143 with the dict keyed by float32 tag[0]; only the first M persons are stored."""
    for a, (tag, joint) in enumerate(dets):
        if ok[a]:
            p = star[a]
            state["joints"][p][k] = joint
            state["tags"][p].append(tag)
            continue
        hit = [p for p in range(len(state["keys"])) if state["keys"][p] == tag[0]]
        if hit:
            p = hit[0]
        else:
            state["created"] += 1
            if len(state["keys"]) < M:
                state["keys"].append(tag[0])
                state["joints"].append({})
                state["tags"].append([])
                p = len(state["keys"]) - 1
            else:
                p = None
        if p is not None:
            state["joints"][p][k] = joint
            state["tags"][p] = [tag]


def place_at_once(state, dets, star, ok, k, M):
    """group_kernel: clash test, then every detection placed from the step's initial state."""
    new = [a for a in range(len(dets)) if not ok[a]]
    keys0 = list(state["keys"])
    clash = any(dets[a][0][0] in keys0 for a in new) or len({dets[a][0][0] for a in new}) != len(new)
    if clash:
        return place_one_by_one(state, dets, star, ok, k, M)
    P = len(keys0)
    ntag0 = [len(t) for t in state["tags"]]
    for a, (tag, joint) in enumerate(dets):
        if ok[a]:
            p = star[a]
            state["joints"][p][k] = joint
            assert len(state["tags"][p]) == ntag0[p]          # a person is matched by at most one detection
            state["tags"][p].append(tag)
    for rank, a in enumerate(new):
        tag, joint = dets[a]
        if P + rank < M:
            state["keys"].append(tag[0])
            state["joints"].append({k: joint})
            state["tags"].append([tag])
    state["created"] += len(new)


def test_parallel_placement_equals_the_reference_loop():
    rng = np.random.default_rng(11)
    clashes = 0
    for trial in range(3000):
        M = int(rng.choice([3, 8, 30]))
        levels = int(rng.choice([4, 50, 10 ** 6]))             # coarse tag grids make key clashes common
        P = int(rng.integers(0, M + 1))
        pop = max(levels * 4, 2 * M)                           # distinct keys for the existing persons
        keys = rng.choice(pop, P, replace=False).astype(np.float32) / np.float32(levels) if P else np.zeros(0, np.float32)
        state = {"keys": [np.float32(x) for x in keys], "joints": [{} for _ in range(P)],
                 "tags": [[(np.float32(x), np.float32(0))] for x in keys], "created": P}
        nr = int(rng.integers(1, M + 1))
        dets = [((np.float32(rng.integers(0, pop) / levels), np.float32(rng.random())), ("joint", trial, a))
                for a in range(nr)]
        # an assignment as the solver returns it: distinct columns; matched if the column is a real person
        cols = rng.permutation(max(P, nr))[:nr]
        star = [int(c) for c in cols]
        ok = [bool(star[a] < P and rng.random() < 0.7) for a in range(nr)]
        s1, s2 = copy.deepcopy(state), copy.deepcopy(state)
        place_one_by_one(s1, dets, star, ok, k=5, M=M)
        new = [a for a in range(nr) if not ok[a]]
        clashes += any(dets[a][0][0] in state["keys"] for a in new) or len({dets[a][0][0] for a in new}) != len(new)
        place_at_once(s2, dets, star, ok, k=5, M=M)
        assert s1 == s2
    assert 100 < clashes < 2900                                # both branches are exercised
