"""Frame-group sharding across GPUs: all-intra pictures are independent, every rank takes
one contiguous group of pictures, there is no collective on the data path; the host splices results in POC order
(reference: TAppEncTop::xWriteOutput, source/App/TAppEncoder/TAppEncTop.cpp:629-693)."""


def frame_group(n_frames, rank, world_size):
    """Contiguous [first, last) picture range of `rank`; groups differ by at most one picture."""
    base, rem = divmod(n_frames, world_size)
    first = rank * base + min(rank, rem)
    return first, first + base + (1 if rank < rem else 0)


def splice_in_poc_order(parts):
    """parts: iterable of (first_poc, [per-picture results]) from the ranks, in any order -> flat list in POC order."""
    out = []
    expect = 0
    for first, items in sorted(parts, key=lambda p: p[0]):
        if first != expect:
            raise ValueError("frame groups are not contiguous: expected POC %d, got %d" % (expect, first))
        out.extend(items)
        expect += len(items)
    return out
