"""Multi-GPU partitioning of the frame-analysis path (host logic only; one process per GPU).

Frame shards (the reference's frame-parallel mode, -F): each rank analyses its own frame against its own nearest
reference frames: no data-path collective inside a frame.  The single exchange step is the broadcast of the newest
reconstructed reference plane from the rank that owns (reconstructed) it -- the producer side of m_reconRowFlag
(/root/reference/source/encoder/framefilter.cpp:664) -- once per step."""


def frame_of(step, rank, world, first=0):
    """Display index of the frame rank `rank` analyses in step `step` (frames are dealt round-robin)."""
    return first + step * world + rank


def ref_owner(step, world):
    """Rank that holds the newest reconstructed reference plane for step `step`."""
    return step % world


def exchange_ref(dist, plane_tensor, step, world, recv=None):
    """Broadcast the newest reference plane (whole padded allocation, margins included) from its owner.
    `dist` is torch.distributed (NCCL on GPUs, gloo in the CPU tests).  With `recv`, the non-owners receive into that
    tensor (an incoming-reference plane the caller swaps in when it wants to) instead of overwriting `plane_tensor`."""
    if world <= 1:
        return plane_tensor
    owner = ref_owner(step, world)
    t = plane_tensor if (recv is None or dist.get_rank() == owner) else recv
    dist.broadcast(t, src=owner)
    return t


# ---- CTU-row shards inside one frame (BASELINE configs[4]: WPP CTU rows sharded per GPU) -----------------------
# A frame's CTU rows are dealt to the ranks; each rank analyses its rows with x265cu_analyser_run_rows() and owns
# their reconstructed pixels.  The one exchange is the broadcast of every owner's reconstructed rows (the
# producer side of m_reconRowFlag, /root/reference/source/encoder/framefilter.cpp:664; the consumer waits on
# refpic->m_reconRowFlag[row + m_refLagRows], frameencoder.cpp:850-868).  Sharding changes placement only:
# every row's arithmetic is the same slice of the same launch lists, so results do not depend on N.

def row_blocks(nrows, rank, world, mode="block"):
    """CTU-row ranges [(r0, r1), ...] owned by `rank`.
    "block": one contiguous chunk per rank (sizes differ by at most one row) -- one launch set per frame;
    "cyclic": row r belongs to rank r % world -- the order a WPP wavefront releases rows in."""
    if mode == "block":
        base, extra = divmod(nrows, world)
        r0 = rank * base + min(rank, extra)
        r1 = r0 + base + (1 if rank < extra else 0)
        return [(r0, r1)] if r1 > r0 else []
    if mode == "cyclic":
        return [(r, r + 1) for r in range(rank, nrows, world)]
    raise ValueError("unknown row shard mode %r" % mode)


def row_owner(row, nrows, world, mode="block"):
    for g in range(world):
        for r0, r1 in row_blocks(nrows, g, world, mode):
            if r0 <= row < r1:
                return g
    raise ValueError("row %d outside the frame" % row)


def band_slice(r0, r1, height, stride, margin_y, es=1, ctu=64):
    """Byte range, inside a margin-extended plane allocation, of the picture rows of CTU rows [r0, r1)
    (whole stride rows: one contiguous band)."""
    y0, y1 = r0 * ctu, min(height, r1 * ctu)
    return (margin_y + y0) * stride * es, (margin_y + y1) * stride * es


def exchange_rows(dist, plane_tensor, nrows, world, height, stride, margin_y, es=1, mode="block"):
    """Every rank broadcasts the bands of `plane_tensor` (flat uint8 view of a whole plane allocation) it owns;
    afterwards every rank holds the complete plane.  Broadcast only, as in the single-plane exchange."""
    if world <= 1:
        return plane_tensor
    works = []
    for g in range(world):
        for r0, r1 in row_blocks(nrows, g, world, mode):
            b0, b1 = band_slice(r0, r1, height, stride, margin_y, es)
            works.append(dist.broadcast(plane_tensor[b0:b1], src=g, async_op=True))
    for w in works:
        w.wait()
    return plane_tensor
