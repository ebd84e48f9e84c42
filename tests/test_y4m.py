"""Y4M container plumbing (daala_b200/y4m.py): round trip, odd picture sizes (chroma rounds up like the reference's
y4m reader, examples/encoder_example.c), header parsing and rejection of formats the path does not handle."""
import numpy as np
import pytest

from daala_b200 import y4m


@pytest.mark.parametrize("w,h", [(64, 48), (33, 17)])
def test_round_trip(tmp_path, w, h):
    rng = np.random.default_rng(w)
    frames = [[rng.integers(0, 256, (h, w), dtype=np.uint8)] +
              [rng.integers(0, 256, ((h + 1) // 2, (w + 1) // 2), dtype=np.uint8) for _ in range(2)] for _ in range(3)]
    path = tmp_path / "a.y4m"
    y4m.write_frames(path, frames, fps="25:1")
    hdr, back = y4m.read_frames(path)
    assert (hdr["width"], hdr["height"], hdr["fps"], hdr["chroma"]) == (w, h, "25:1", "420jpeg")
    assert len(back) == 3
    for a, b in zip(frames, back):
        for p in range(3):
            assert np.array_equal(a[p], b[p])
    assert len(y4m.read_frames(path, max_frames=2)[1]) == 2


def test_rejects_what_it_does_not_handle(tmp_path):
    p = tmp_path / "b.y4m"
    p.write_bytes(b"YUV4MPEG2 W16 H16 F30:1 Ip A1:1 C444\nFRAME\n" + bytes(16 * 16 * 3))
    with pytest.raises(y4m.Y4MError):
        y4m.read_frames(p)
    p.write_bytes(b"YUV4MPEG2 W16 H16 F30:1 Ip A1:1 C420p10\nFRAME\n" + bytes(16 * 16 * 3))
    with pytest.raises(y4m.Y4MError):
        y4m.read_frames(p)
    p.write_bytes(b"YUV4MPEG2 W16 H16 C420jpeg\nFRAME\n" + bytes(100))
    with pytest.raises(y4m.Y4MError):
        y4m.read_frames(p)
    p.write_bytes(b"RIFF....")
    with pytest.raises(y4m.Y4MError):
        y4m.read_frames(p)
