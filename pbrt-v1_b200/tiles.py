"""Multi-GPU partition of an image into independent pixel tiles + the one film exchange.

The reference's only distribution story is: render disjoint `cropwindow`s in separate
processes and paste them by data-window offset (film/image.cpp:79-85,148-156,
tools/exrassemble.cpp:53-66).  A crop render samples the filter apron around its pixels, so
every pixel it owns is complete.  With the per-sample keyed RNG a tile rendered alone is
bit-identical to the same pixels of a full-frame render, so the exchange is a pure gather:
no arithmetic, one collective.

One process per GPU (torch.distributed): the raster-ordered tile list is dealt round-robin with
the start shifted by one per tile row (diagonals: balances cheap ceiling tiles against expensive
glass tiles in both directions), packs their resolved RGBA pixels into one buffer and takes part in a single
`all_gather_into_tensor`; rank 0 pastes tiles into the frame.
"""
import numpy as np


def make_tiles(x0, y0, width, height, tile=128):
    """Raster-ordered list of (x0, y0, w, h) pixel windows covering the crop window."""
    tiles = []
    for ty in range(y0, y0 + height, tile):
        for tx in range(x0, x0 + width, tile):
            tiles.append((tx, ty, min(tile, x0 + width - tx), min(tile, y0 + height - ty)))
    return tiles


def tiles_of_rank(tiles, rank, world):
    """Tile i (raster order) goes to rank (i + row(i)) % world: round-robin with the start shifted by one per
    tile row, so that a rank's tiles lie on diagonals instead of in one column of the image (with 8 tiles
    per row and 8 ranks plain round-robin hands one rank the whole glass-sphere column)."""
    if world <= 1 or not tiles:
        return list(tiles)
    ncols = sum(1 for t in tiles if t[1] == tiles[0][1])
    return [t for i, t in enumerate(tiles) if (i + i // ncols) % world == rank]


def slot_floats(tiles, world):
    """Floats each rank contributes to the gather (padded to the largest rank)."""
    per_rank = [sum(t[2] * t[3] * 4 for t in tiles_of_rank(tiles, r, world)) for r in range(world)]
    return max(per_rank) if per_rank else 0


def pack_offsets(my_tiles):
    offs, o = [], 0
    for t in my_tiles:
        offs.append(o)
        o += t[2] * t[3] * 4
    return offs, o


def assemble(gathered, tiles, world, x0, y0, width, height):
    """gathered: float32 [world, slot_floats] (numpy).  Returns the [height, width, 4] frame."""
    frame = np.zeros((height, width, 4), np.float32)
    for r in range(world):
        mine = tiles_of_rank(tiles, r, world)
        offs, _ = pack_offsets(mine)
        for t, o in zip(mine, offs):
            tx, ty, tw, th = t
            frame[ty - y0:ty - y0 + th, tx - x0:tx - x0 + tw] = gathered[r, o:o + tw * th * 4].reshape(th, tw, 4)
    return frame


def render_distributed(render_tiles, window, rank, world, dist=None, device=None, tile=128):
    """render_tiles(list of (x0,y0,w,h)) -> list of arrays/tensors [h,w,4] (numpy or torch).

    Runs this rank's share, then ONE all_gather of the packed tiles.  Returns the assembled
    frame on rank 0 (numpy) and None elsewhere.  `dist` is torch.distributed (any backend) or
    None for world == 1.
    """
    x0, y0, width, height = window
    tiles = make_tiles(x0, y0, width, height, tile)
    mine = tiles_of_rank(tiles, rank, world)
    images = render_tiles(mine) if mine else []
    n = slot_floats(tiles, world)
    if world == 1 or dist is None:
        buf = np.zeros(n, np.float32)
        offs, _ = pack_offsets(mine)
        for img, o in zip(images, offs):
            a = img.detach().cpu().numpy() if hasattr(img, "detach") else np.asarray(img)
            buf[o:o + a.size] = a.reshape(-1)
        return assemble(buf[None], tiles, 1, x0, y0, width, height)
    import torch
    dev = device if device is not None else torch.device("cpu")
    send = torch.zeros(n, dtype=torch.float32, device=dev)
    offs, _ = pack_offsets(mine)
    for img, o in zip(images, offs):
        t = img if hasattr(img, "detach") else torch.from_numpy(np.ascontiguousarray(img))
        send[o:o + t.numel()] = t.reshape(-1).to(dev)
    recv = torch.empty(world * n, dtype=torch.float32, device=dev)
    dist.all_gather_into_tensor(recv, send)      # the single film exchange
    if rank != 0:
        return None
    return assemble(recv.reshape(world, n).cpu().numpy(), tiles, world, x0, y0, width, height)
