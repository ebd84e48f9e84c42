"""Superblock-row sharding across ranks and the one exchange the path needs.

Each rank owns a contiguous range of superblock rows of every frame
(SURVEY.md 8(e)).  Forward transform and PVQ need no communication.  The
superblock-edge postfilter (reference src/filter.c:1561) reaches two samples
across every superblock edge, so before it runs each rank needs the two lapped
rows just outside its range: ONE all-gather per step of every rank's top-2 and
bottom-2 int32 rows of all planes and frames (NCCL on GPUs, gloo in the CPU
tests), then local copies into the halo rows.
"""
import torch
import torch.distributed as dist


def plane_rows(geom, pli, sb_row0, sb_rows, halo=0):
    """Pixel rows [a, b) of plane `pli` covered by the superblock rows, widened by `halo`."""
    sb = 64 >> geom.xdec[pli]
    ph = geom.plane_shape(pli)[0]
    return max(0, sb_row0 * sb - halo), min(ph, (sb_row0 + sb_rows) * sb + halo)


class BorderExchange:
    """All-gather of the lapped border rows of `planes` (list of [F, h, w] int32
    tensors, full-size on every rank; each rank has computed only its rows)."""

    def __init__(self, geom, planes, rank, world, group=None):
        self.geom, self.planes, self.rank, self.world, self.group = geom, planes, rank, world, group
        self.r0, self.nrows = geom.shard_rows(rank, world)
        F = planes[0].shape[0]
        self.layout = []
        off = 0
        for pli, t in enumerate(planes):
            a, b = plane_rows(geom, pli, self.r0, self.nrows)
            w = t.shape[2]
            n = F * 4 * w
            self.layout.append((off, n, w, a, b))
            off += n
        dev = planes[0].device
        self.send = torch.zeros(off, dtype=torch.int32, device=dev)
        self.gathered = torch.zeros((world, off), dtype=torch.int32, device=dev)
        self.bytes_per_rank = off * 4

    def __call__(self):
        if self.world == 1:
            return
        F = self.planes[0].shape[0]
        for t, (off, n, w, a, b) in zip(self.planes, self.layout):
            v = self.send[off:off + n].view(F, 4, w)
            v[:, 0:2].copy_(t[:, a:a + 2])
            v[:, 2:4].copy_(t[:, b - 2:b])
        dist.all_gather_into_tensor(self.gathered.view(-1), self.send, group=self.group)
        for t, (off, n, w, a, b) in zip(self.planes, self.layout):
            if self.rank > 0:
                t[:, a - 2:a].copy_(self.gathered[self.rank - 1, off:off + n].view(F, 4, w)[:, 2:4])
            if self.rank < self.world - 1:
                t[:, b:b + 2].copy_(self.gathered[self.rank + 1, off:off + n].view(F, 4, w)[:, 0:2])
