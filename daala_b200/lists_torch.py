"""Work lists of a keyframe batch built with tensor ops ON THE DEVICE that holds the block-size maps.

Same arrays, same order as `pvq.native_keyframe_lists` / the numpy builders (tests/test_host_logic.py
compares them on CPU tensors; the ops are device-agnostic).  This is index bookkeeping -- sorts, prefix
sums, gathers over <= 10^6 blocks -- i.e. the kind of plumbing torch is here for; the PVQ / transform
kernels that consume the lists are the library's own.  Purpose: a live encoder changes block sizes
every frame; with the maps already on the GPU the lists never touch the host (the native host builder
costs 4x the GPU step for 16 4K frames).

Reference semantics: leaf blocks of the block-size quadtree (src/block_size.h), bs = max(obs, xdec) on
chroma (src/encode.c:1467), same-size top / left neighbour of od_hv_intra_pred (src/intra.c:46-47),
co-located-luma-is-4x4 flag of od_resample_luma_coeffs (src/intra.c:78), band layout OD_BAND_OFFSETS
(src/partition.c:85-91).
"""
import torch

NBANDS = (1, 4, 7, 9, 9)


def _records(coef_off, x0, y0, bs, pli, xdec, frame):
    """[n, 12] uint8: daala_b200_pvq_block records (little endian)."""
    w0 = coef_off.to(torch.int64)
    w1 = x0.to(torch.int64) | (y0.to(torch.int64) << 16)
    w2 = bs.to(torch.int64) | (pli.to(torch.int64) << 8) | (xdec.to(torch.int64) << 16) | (frame.to(torch.int64) << 24)
    w = torch.stack([w0, w1, w2], dim=1)
    w = torch.where(w >= 2 ** 31, w - 2 ** 32, w).to(torch.int32)
    return w.contiguous().view(torch.uint8).reshape(-1, 12)


def _offsets(bs):
    length = torch.clamp(16 << (2 * bs.to(torch.int64)), max=512)
    off = torch.cumsum(length, 0) - length
    return off, int(length.sum().item()) if len(length) else 0


def _depth(*parents):
    """Longest chain ending at each element; parents: index tensors (-1 = none).  Iterated relaxation,
    one pass per wavefront (parents always precede their children)."""
    n = len(parents[0])
    depth = torch.ones(n, dtype=torch.int32, device=parents[0].device)
    if n == 0:
        return depth
    has = [p >= 0 for p in parents]
    safe = [torch.clamp(p, min=0).to(torch.int64) for p in parents]
    while True:
        new = torch.ones_like(depth)
        for h, s in zip(has, safe):
            new = torch.maximum(new, torch.where(h, depth[s] + 1, torch.ones_like(depth)))
        if torch.equal(new, depth):
            return depth
        depth = new


def keyframe_lists(maps, nhsb, nvsb):
    """maps: uint8 tensor [F, nvsb*8, nhsb*8] (any device).  Returns a dict of tensors on that device:
    luma / chroma ([n, 12] uint8 block records), dep_top / dep_left / depth (int32), luma_total /
    chroma_total (python ints), chain / chain_wave / bulk / chroma_lists (per class 16 / 32 / 128) and
    chain_slices (python lists of (first, count))."""
    dev = maps.device
    F, bh, bw = maps.shape
    assert (bh, bw) == (nvsb * 8, nhsb * 8) and F <= 255
    m = maps.to(torch.int64)
    h4, w4 = bh * 2, bw * 2
    # ---- luma blocks in (frame, y, x) order ----
    fs, ys, xs, bss = [], [], [], []
    f0, u0, v0 = torch.nonzero(m == 0, as_tuple=True)             # 8x8 units coded as four 4x4 blocks
    for dy in (0, 1):
        for dx in (0, 1):
            fs.append(f0); ys.append(u0 * 2 + dy); xs.append(v0 * 2 + dx); bss.append(torch.zeros_like(f0))
    uy = torch.arange(bh, device=dev).view(1, bh, 1)
    ux = torch.arange(bw, device=dev).view(1, 1, bw)
    for lvl in range(1, 5):
        span = 1 << (lvl - 1)
        sel = (m == lvl) & (uy % span == 0) & (ux % span == 0)
        f, u, v = torch.nonzero(sel, as_tuple=True)
        fs.append(f); ys.append(u * 2); xs.append(v * 2); bss.append(torch.full_like(f, lvl))
    fr, y4, x4, bs = (torch.cat(t) for t in (fs, ys, xs, bss))
    order = torch.argsort((fr * h4 + y4) * w4 + x4)
    fr, y4, x4, bs = fr[order], y4[order], x4[order], bs[order]
    n = len(fr)
    index = torch.full((F, h4, w4), -1, dtype=torch.int64, device=dev)
    index[fr, y4, x4] = torch.arange(n, device=dev)
    n4 = 1 << bs
    same_top = (y4 > 0) & (m[fr, torch.clamp(y4 - 1, min=0) >> 1, x4 >> 1] == bs)
    same_left = (x4 > 0) & (m[fr, y4 >> 1, torch.clamp(x4 - 1, min=0) >> 1] == bs)
    top = torch.where(same_top, index[fr, torch.clamp(y4 - n4, min=0), x4], torch.full_like(fr, -1))
    left = torch.where(same_left, index[fr, y4, torch.clamp(x4 - n4, min=0)], torch.full_like(fr, -1))
    none = torch.full_like(top, -1)
    d0, dh, dv = _depth(top, left), _depth(top, none), _depth(none, left)
    # ---- stable sort by depth, remap the neighbour indices ----
    order = torch.sort(d0, stable=True).indices
    inv = torch.empty_like(order)
    inv[order] = torch.arange(n, device=dev)
    fr, y4, x4, bs, d0, dh, dv = (t[order] for t in (fr, y4, x4, bs, d0, dh, dv))
    top, left = top[order], left[order]
    top = torch.where(top >= 0, inv[torch.clamp(top, min=0)], top).to(torch.int32)
    left = torch.where(left >= 0, inv[torch.clamp(left, min=0)], left).to(torch.int32)
    off, luma_total = _offsets(bs)
    zero = torch.zeros_like(bs)
    out = dict(luma=_records(off, x4 * 4, y4 * 4, bs, zero, zero, fr), dep_top=top, dep_left=left, depth=d0,
               luma_total=luma_total, chain={}, chain_wave={}, chain_slices={}, bulk={}, chroma_lists={})
    # ---- band-granular waves: key (wave, band, block) ----
    nbands = torch.tensor(NBANDS, device=dev)[bs]
    blk = torch.arange(n, device=dev)
    per_band = (d0, dh, dv)
    for c, key in enumerate((16, 32, 128)):
        ents, waves, bulk = [], [], []
        for band in range(3 * c, 3 * c + 3):
            own = nbands > band
            e = (blk[own] << 4) | band
            if band in (3, 6):
                bulk.append(e)
            else:
                ents.append(e)
                waves.append(per_band[band % 3][own].to(torch.int64))
        e = torch.cat(ents) if ents else torch.zeros(0, dtype=torch.int64, device=dev)
        w = torch.cat(waves) if waves else torch.zeros(0, dtype=torch.int64, device=dev)
        o = torch.argsort((w * 16 + (e & 15)) * max(n, 1) + (e >> 4))
        e, w = e[o], w[o]
        out["chain"][key] = e.to(torch.int32)                       # bit pattern of the uint32 entries
        out["chain_wave"][key] = (w - 1).to(torch.int16)
        top_d = int(w.max().item()) if len(w) else 0
        counts = torch.bincount(w, minlength=top_d + 1)[1:] if len(w) else torch.zeros(0, dtype=torch.int64)
        first = torch.cumsum(counts, 0) - counts
        out["chain_slices"][key] = [(int(a), int(b)) for a, b in zip(first.tolist(), counts.tolist())]
        out["bulk"][key] = (torch.cat(bulk) if bulk else torch.zeros(0, dtype=torch.int64, device=dev)).to(torch.int32)
    # ---- chroma: order (size, frame, plane, y, x) ----
    eff = torch.clamp(m, min=1)
    keys, recs = [], []
    for cbs in range(4):
        lvl, span = cbs + 1, 1 << cbs
        sel = (eff == lvl) & (uy % span == 0) & (ux % span == 0)
        f, u, v = torch.nonzero(sel, as_tuple=True)
        flag = ((m[f, u, v] == 0) & (cbs == 0)).to(torch.int64) * 0x80
        for pli in (1, 2):
            keys.append(((cbs * F + f) * 2 + (pli - 1)) * (bh * bw) + u * bw + v)
            recs.append((f, u * 4, v * 4, torch.full_like(f, cbs), torch.full_like(f, pli), 1 | flag))
    key = torch.cat(keys)
    o = torch.argsort(key)
    cf, cy, cx, cbs_t, cpli, cxdec = (torch.cat([r[i] for r in recs])[o] for i in range(6))
    coff, chroma_total = _offsets(cbs_t)
    out["chroma"] = _records(coff, cx, cy, cbs_t, cpli, cxdec, cf)
    out["chroma_total"] = chroma_total
    nb_c = torch.tensor(NBANDS, device=dev)[cbs_t]
    cblk = torch.arange(len(cbs_t), device=dev)
    for c, key in enumerate((16, 32, 128)):
        parts = [(cblk[nb_c > band] << 4) | band for band in range(3 * c, 3 * c + 3)]
        out["chroma_lists"][key] = torch.cat(parts).to(torch.int32)
    return out
