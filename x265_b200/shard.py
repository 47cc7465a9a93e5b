"""Multi-GPU partitioning of the frame-analysis path (host logic only; one process per GPU).

The frames of a mini-GOP share one reference set, so each rank analyses its own frame: no data-path
collective inside a frame.  The single exchange step is the broadcast of the newest reconstructed
reference plane from the rank that owns (reconstructed) it -- the producer side of m_reconRowFlag
(/root/reference/source/encoder/framefilter.cpp:664) -- once per step."""


def frame_of(step, rank, world, first=0):
    """Display index of the frame rank `rank` analyses in step `step` (frames are dealt round-robin)."""
    return first + step * world + rank


def ref_owner(step, world):
    """Rank that holds the newest reconstructed reference plane for step `step`."""
    return step % world


def exchange_ref(dist, plane_tensor, step, world):
    """Broadcast the newest reference plane (whole padded allocation, margins included) from its owner.
    `dist` is torch.distributed (NCCL on GPUs, gloo in the CPU tests)."""
    if world > 1:
        dist.broadcast(plane_tensor, src=ref_owner(step, world))
    return plane_tensor
