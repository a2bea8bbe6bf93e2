"""Multi-GPU sharding of read groups: contiguous query-name ranges, cut at group boundaries, one
range per rank, results concatenated in rank order.  No collective on the data path.

Reference precedent: secphase_index records an offset every `step_size` read names
(programs/src/secphase_index.c:76-109) and get_offset_array splits them evenly across workers
(programs/src/secphase.c:372-382).  Read groups are independent jobs (secphase.c:303), so the
only cross-shard state is the order of the output records (compare as a set, or
concatenate in shard order as done here) and the un-seeded rand() stream of the tie-breaks
(ptAlignment.c:156-171), which each shard replays from seed 1 in its own order.
"""
import numpy as np


def shard_range(n_groups, rank, world):
    """[g0, g1) of `rank`: sizes differ by at most one, earlier ranks take the larger shards."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(n_groups, world)
    g0 = rank * base + min(rank, extra)
    return g0, g0 + base + (1 if rank < extra else 0)


def merge_results(parts):
    """Concatenate per-shard result dicts (as returned by Secphase.wait / run) in rank order."""
    out = {}
    for key in ("groups", "scores", "extents", "markers_final"):
        out[key] = np.concatenate([p[key] for p in parts])
    off = [np.zeros(1, np.int64)]
    base = 0
    for p in parts:
        o = np.asarray(p["markers_final_off"], np.int64)
        off.append(o[1:] + base)
        base += int(o[-1])
    out["markers_final_off"] = np.concatenate(off)
    for key in ("hmm_instances", "hmm_cells", "n_groups", "n_alns"):
        if all(key in p for p in parts):
            out[key] = sum(int(p[key]) for p in parts)
    return out
