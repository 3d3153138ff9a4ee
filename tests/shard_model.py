"""TEST INFRASTRUCTURE: numpy model of the position-sharded ordered insert (abyss_b200/csrc/abb_shard.cuh and
sharded_ordered_insert in abb_api.cu), with the collective passed in as a function, so that the PROTOCOL -- partial
minima + veto flags, one all-reduce(min) per step, file-order carry list, oldest-prefix drain steps -- can be checked
against the sequential oracle on CPU over gloo at any world size.  Same structure as the CUDA path: every rank walks
the same global windows, evaluates every lane, and touches only the counters it owns."""
import numpy as np


def shard_range(m, rank, world):
    chunk = ((m + world - 1) // world + 15) & ~15
    lo = min(m, rank * chunk)
    return lo, min(m, lo + chunk)


def sharded_insert_model(pos, m, rank, world, allreduce_min, window=64, map_entries=1 << 10, carry_lanes=32):
    """pos: (n_slots, H) filter positions in file order (identical on every rank).
    Returns this rank's counters (only [lo, hi) is meaningful) and the number of steps (= collectives)."""
    pos = np.asarray(pos, dtype=np.int64)
    n_slots, H = pos.shape
    lo, hi = shard_range(m, rank, world)
    counters = np.zeros(m, dtype=np.uint8)
    own = (pos >= lo) & (pos < hi)
    pending = np.zeros(0, dtype=np.int64)  # carried slots, ascending = file order
    steps = 0

    def step(new):
        nonlocal pending, steps
        lanes_c = pending[:carry_lanes]
        rest = pending[carry_lanes:]
        lanes = np.concatenate([lanes_c, new])
        if lanes.size == 0:
            return
        P, O = pos[lanes], own[lanes]
        pm = np.where(O, counters[P], 255).min(axis=1).astype(np.uint8)
        ok = np.ones(lanes.size, dtype=np.uint8)
        nc = lanes_c.size
        # conflict map (two-bit entries): marks of the new slots' own positions, "touched again" on a second mark or a carried mark
        ent = P & (map_entries - 1)
        cnt = np.bincount(ent[nc:][O[nc:]], minlength=map_entries)
        cnt += 2 * np.bincount(ent[:nc][O[:nc]], minlength=map_entries).clip(0, 1)
        again = cnt >= 2
        ok[nc:] = ~(again[ent[nc:]] & O[nc:]).any(axis=1)
        # tag table: a carried slot must be the oldest carried slot on each own position
        first = {}
        for j in range(nc):
            for p in P[j][O[j]]:
                first.setdefault(int(p), j)
        for j in range(nc):
            ok[j] = all(first[int(p)] == j for p in P[j][O[j]])
        buf = np.concatenate([pm, ok])
        buf = allreduce_min(buf)
        steps += 1
        pm, ok = buf[:lanes.size], buf[lanes.size:]
        good = ok.astype(bool)
        for j in np.nonzero(good)[0]:
            mn = pm[j]
            if mn == 255:
                continue
            for p in set(int(x) for x in P[j][O[j]]):
                if counters[p] == mn:
                    counters[p] = mn + 1
        pending = np.sort(np.concatenate([lanes[~good], rest]))

    for w0 in range(0, n_slots, window):
        while pending.size > carry_lanes:
            step(np.zeros(0, dtype=np.int64))
        step(np.arange(w0, min(n_slots, w0 + window), dtype=np.int64))
    while pending.size:
        step(np.zeros(0, dtype=np.int64))
    return counters, steps
