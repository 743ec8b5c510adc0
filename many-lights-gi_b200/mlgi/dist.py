"""Host-side logic of the multi-GPU decomposition.

One process per GPU. The frame shards by interleave class: rank r of G owns classes
[r*16/G, (r+1)*16/G) = pixels with (x&3) + 4*(y&3) in that range = VPL slots [lo*N/16, hi*N/16) =
a contiguous band of ISM rows. Each rank writes its classes' planes into the same
[16][ceil(H/4)][ceil(W/4)][4] fp32 exchange buffer; ONE in-place all-gather assembles the image
(equal-sized contributions), and the consumer de-interleaves it (mlgi_planar_to_raster on the GPU).
No arithmetic depends on G, so the G-GPU image is bit-identical to the 1-GPU image.
Backend-agnostic torch.distributed calls: NCCL over NVLink on the GPU box, gloo in the CPU tests.
"""
import numpy as np

NUM_CLASSES = 16


def class_range(rank, world):
    if world < 1 or NUM_CLASSES % world:
        raise ValueError("world size must divide 16 interleave classes, got %d" % world)
    per = NUM_CLASSES // world
    return rank * per, (rank + 1) * per


TILE = 64


def row_mask(rank, world, height, tile=TILE):
    """image rows rank `rank` of `world` is responsible for in the sharded frames (mlgi_sharded_rows): the 64-pixel tile
    rows t with t mod world == rank -- the tiles it bounds, the G-buffer rows it uploads and the image rows it returns."""
    y = np.arange(height)
    return (y // tile) % world == rank


def host_frame_phases(tile_rows, world, phases):
    """phases of mlgi_frame_sharded_host as (first tile row, end tile row) blocks: whole tile rows, every block but the
    last a multiple of `world` rows high, so that each rank owns the same share of every phase (same arithmetic as the library)"""
    k = max(1, min(phases, (tile_rows + world - 1) // world))
    rows_per_phase = world * ((tile_rows + k * world - 1) // (k * world))
    return [(a, min(tile_rows, a + rows_per_phase)) for a in range(0, tile_rows, rows_per_phase)]


def planar_shape(width, height):
    return (NUM_CLASSES, (height + 3) // 4, (width + 3) // 4, 4)


def allgather_planar(planar, rank, world, group=None):
    """in-place all-gather of the class planes: `planar` is the full [16,PH,PW,4] tensor on every rank,
    rank r has filled planar[lo:hi]."""
    import torch.distributed as dist
    lo, hi = class_range(rank, world)
    dist.all_gather_into_tensor(planar.view(-1), planar[lo:hi].reshape(-1), group=group)
    return planar


def raster_to_planar(raster):
    """numpy reference of the layout: raster [H,W,C] -> planar [16,PH,PW,C] (zero padded)"""
    H, W, C = raster.shape
    PH, PW = (H + 3) // 4, (W + 3) // 4
    out = np.zeros((NUM_CLASSES, PH, PW, C), raster.dtype)
    for cls in range(NUM_CLASSES):
        cx, cy = cls & 3, cls >> 2
        sub = raster[cy::4, cx::4]
        out[cls, :sub.shape[0], :sub.shape[1]] = sub
    return out


def planar_to_raster(planar, width, height):
    """numpy reference of mlgi_planar_to_raster"""
    out = np.zeros((height, width, planar.shape[-1]), planar.dtype)
    for cls in range(NUM_CLASSES):
        cx, cy = cls & 3, cls >> 2
        sub = out[cy::4, cx::4]
        sub[...] = planar[cls, :sub.shape[0], :sub.shape[1]]
    return out
