"""Multi-GPU partitioning of the ray-cast path (host logic, no GPU needed).

Every pixel of every frame is independent and the scene is read-only, so
the path shards with no data-path collective: the scene is replicated on every rank, a single frame
is split into interleaved 32x32 image tiles (tile t belongs to rank t % world), an animation sweep is
split into blocks of consecutive frames, and the only exchange is the delivery of finished tiles to the
rank that owns the frame (frame f -> rank f % world in the fused peer push; rank 0 in the NCCL gather).  The functions here define who renders what and how rank 0 reassembles it; they
are shared by bench.py (NCCL) and tests/test_multi_cpu.py (gloo).
"""
import numpy as np

TILE = 32  # must equal rtb::kTile (csrc/rtb_kernels.cuh)


def tiles_xy(W, H):
    return (W + TILE - 1) // TILE, (H + TILE - 1) // TILE


def tile_owner_map(W, H, world):
    """(H, W) int array: rank that renders each pixel when tiles are dealt round-robin."""
    tx, ty = tiles_xy(W, H)
    t = (np.arange(H)[:, None] // TILE) * tx + (np.arange(W)[None, :] // TILE)
    return (t % world).astype(np.int32)


def tile_mask(W, H, rank, world):
    """Flat boolean mask (row 0 = bottom, like the frame) of the pixels rank `rank` renders."""
    return (tile_owner_map(W, H, world) == rank).reshape(-1)


def frame_block(step, rank, world, frames_per_step):
    """Global frame indices of the block rank `rank` renders in step `step` of a sweep."""
    first = (step * world + rank) * frames_per_step
    return first, first + frames_per_step


def frame_owner(frame, world):
    """Striped ownership of the fused tile exchange: frame `frame` of a step is assembled on rank frame % world, as that
    rank's frame frame // world (rtb_render_frames_push_striped_async, RenderParams::push_owners)."""
    return frame % world, frame // world


def owned_frames(rank, world, frames_per_step):
    """Frames of a step of `frames_per_step` frames that rank `rank` owns, in the order of its local frame index."""
    return list(range(rank, frames_per_step, world))


def compose_tiles(parts, W, H):
    """Reassemble a frame from `world` full-size buffers, each valid only on its owner's tiles."""
    world = len(parts)
    owner = tile_owner_map(W, H, world).reshape(-1)
    out = np.empty_like(np.asarray(parts[0]))
    for r, p in enumerate(parts):
        m = owner == r
        out[m] = np.asarray(p)[m]
    return out


def compose_tiles_torch(parts, W, H):
    """Same as compose_tiles for torch tensors on one device (rank 0 after the NCCL gather)."""
    import torch
    world = len(parts)
    owner = torch.from_numpy(tile_owner_map(W, H, world).reshape(-1)).to(parts[0].device)
    out = parts[0].clone()
    for r in range(1, world):
        out = torch.where(owner == r, parts[r], out)
    return out
