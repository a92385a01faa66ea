"""Plain numpy restatement of ``mpg_compact_map`` (csrc/extra.cu, include/mpgan_b200.h): test infrastructure only.
Jet j gets max(n_j, 15) consecutive positions of the compacted row space (its unmasked particles first, in order);
tile t = positions [128 t, 128 t + 127]."""
import numpy as np

MINW, TILE = 15, 128


def tiles_max(B, N):
    return (B * max(N, MINW) + TILE - 1) // TILE


def compact_map_ref(mask):
    mask = np.asarray(mask).reshape(mask.shape[0], -1)
    B, N = mask.shape
    live = mask != 0
    cnt = live.sum(1)
    width = np.maximum(cnt, MINW)
    start = np.concatenate(([0], np.cumsum(width)[:-1]))
    total = int(width.sum())
    nt, tmax = (total + TILE - 1) // TILE, tiles_max(B, N)
    rowmap = -np.ones(tmax * TILE, dtype=np.int64)
    for j in range(B):
        idx = np.nonzero(live[j])[0]
        rowmap[start[j]:start[j] + cnt[j]] = j * N + idx
    tile_j0 = np.zeros(tmax, dtype=np.int64)
    tile_nj = np.zeros(tmax, dtype=np.int64)
    for t in range(nt):
        a, b = t * TILE, t * TILE + TILE - 1
        jets = [j for j in range(B) if cnt[j] > 0 and start[j] <= b and start[j] + cnt[j] > a]
        if jets:
            tile_j0[t], tile_nj[t] = jets[0], jets[-1] - jets[0] + 1
    return {"tiles": nt, "total": total, "tile_j0": tile_j0, "tile_nj": tile_nj, "rowmap": rowmap, "tmax": tmax}


def as_cmap(ref):
    """The int32 layout of the device map: [tiles, total, tile_j0[tmax], tile_nj[tmax], rowmap[tmax * 128]]."""
    return np.concatenate(([ref["tiles"], ref["total"]], ref["tile_j0"], ref["tile_nj"], ref["rowmap"])).astype(np.int32)
