"""YUV4MPEG2 (Y4M) reader / writer for 8-bit 4:2:0 material -- the container the reference's tools feed the
encoder with (examples/encoder_example.c reads Y4M, examples/dump_video.c writes it).  Host-side plumbing
for tools/encode_y4m.py; no codec logic here."""
import numpy as np


class Y4MError(ValueError):
    pass


def _parse_header(line):
    if not line.startswith(b"YUV4MPEG2"):
        raise Y4MError("not a YUV4MPEG2 stream")
    info = {"W": None, "H": None, "F": "30:1", "I": "p", "A": "1:1", "C": "420jpeg"}
    for tok in line.split()[1:]:
        key, val = chr(tok[0]), tok[1:].decode("ascii")
        if key in info:
            info[key] = val
    if info["W"] is None or info["H"] is None:
        raise Y4MError("header without W/H")
    if info["C"] not in ("420", "420jpeg", "420mpeg2", "420paldv"):
        raise Y4MError("only 8-bit 4:2:0 is supported, got C%s" % info["C"])
    return dict(width=int(info["W"]), height=int(info["H"]), fps=info["F"], interlace=info["I"],
                aspect=info["A"], chroma=info["C"])


def read_frames(path, max_frames=None):
    """Returns (header dict, list of [y, u, v] uint8 planes); chroma planes are ((w+1)//2, (h+1)//2)."""
    frames = []
    with open(path, "rb") as f:
        hdr = _parse_header(f.readline().rstrip(b"\n"))
        w, h = hdr["width"], hdr["height"]
        cw, ch = (w + 1) // 2, (h + 1) // 2
        while max_frames is None or len(frames) < max_frames:
            line = f.readline()
            if not line:
                break
            if not line.startswith(b"FRAME"):
                raise Y4MError("expected FRAME, got %r" % line[:16])
            buf = f.read(w * h + 2 * cw * ch)
            if len(buf) != w * h + 2 * cw * ch:
                raise Y4MError("truncated frame %d" % len(frames))
            a = np.frombuffer(buf, np.uint8)
            frames.append([a[:w * h].reshape(h, w).copy(), a[w * h:w * h + cw * ch].reshape(ch, cw).copy(),
                           a[w * h + cw * ch:].reshape(ch, cw).copy()])
    return hdr, frames


def write_frames(path, frames, fps="30:1", aspect="1:1", chroma="420jpeg"):
    """frames: iterable of [y, u, v] uint8 planes of one geometry."""
    frames = list(frames)
    if not frames:
        raise Y4MError("no frames")
    h, w = frames[0][0].shape
    with open(path, "wb") as f:
        f.write(("YUV4MPEG2 W%d H%d F%s Ip A%s C%s\n" % (w, h, fps, aspect, chroma)).encode("ascii"))
        for planes in frames:
            if planes[0].shape != (h, w) or planes[1].shape != ((h + 1) // 2, (w + 1) // 2) or planes[2].shape != planes[1].shape:
                raise Y4MError("frame geometry changed")
            f.write(b"FRAME\n")
            for p in planes:
                f.write(np.ascontiguousarray(p, np.uint8).tobytes())
