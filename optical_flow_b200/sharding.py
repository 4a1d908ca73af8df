"""Host-side partitioning of a shot's frame pairs over GPUs.

Pairs are independent (visualize_optical_flow.py:38 and optical_flow.py:98 each depend only on their two
frames; the min-max of the picture is per frame, visualize_optical_flow.py:54), so there is NO collective
on the data path: every rank takes a contiguous range of pairs plus one overlap frame and writes its own
slice of the output.
"""


def shard_pairs(n_pairs, world_size, rank):
    """Contiguous, balanced range [start, stop) of pair indices for `rank`; pair t uses frames t and t+1,
    so the rank needs frames [start, stop] (one frame of overlap with its right neighbour)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world_size")
    base, rem = divmod(int(n_pairs), world_size)
    start = rank * base + min(rank, rem)
    stop = start + base + (1 if rank < rem else 0)
    return start, stop


def shard_shots(shot_lengths, world_size):
    """Greedy longest-first assignment of whole shots (lengths in pairs) to ranks; shots longer than
    total/world_size are split into contiguous pieces first.  Returns per rank a list of
    (shot_index, first_pair, n_pairs)."""
    total = sum(int(x) for x in shot_lengths)
    cap = max(1, -(-total // world_size))
    pieces = []
    for i, n in enumerate(shot_lengths):
        n = int(n)
        off = 0
        while n - off > cap:
            pieces.append((i, off, cap))
            off += cap
        if n - off > 0:
            pieces.append((i, off, n - off))
    pieces.sort(key=lambda p: (-p[2], p[0], p[1]))
    loads = [0] * world_size
    out = [[] for _ in range(world_size)]
    for p in pieces:
        r = min(range(world_size), key=lambda j: (loads[j], j))
        out[r].append(p)
        loads[r] += p[2]
    for r in range(world_size):
        out[r].sort()
    return out
