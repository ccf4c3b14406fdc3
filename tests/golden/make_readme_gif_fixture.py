"""Generate tests/golden/readme_gif_classes.npz from the reference's README GIF.

Run HERE (the container that has /root/reference); the GPU box only sees the committed .npz.
The GIF is the literal output of BASELINE config 1 (README.fut:15: `:video (step, init 30 30 10 123, 400)`).
Frame k = render of the grid after k ticks (animation.fut:22-24 renders before stepping).

Per frame and per 10x10 block we store the majority colour class:
  0 '.' black (no road)   1 'W' white (empty road)   2 'R' red   3 'B' blue
  4 'O' orange (renders yellow)   5 'M' magenta/violet   6 'N' brown
The ring that moved on each transition (tape) is re-derived and stored too.
"""
import hashlib
import sys
from pathlib import Path

import numpy as np
from PIL import Image

GIF = Path("/root/reference/README-img/16217933fbe9784886c68978d477c849-video.gif")
OUT = Path(__file__).resolve().parent / "readme_gif_classes.npz"
SHA = "ef2692f2bd5a08b5352e614095fa1bf92fcf1b14c16e28de4e1e0af9ffd6228b"


def classify(rgb):
    r, g, b = (rgb[..., i].astype(np.int32) for i in range(3))
    cls = np.full(r.shape, 255, np.uint8)
    cls[(r < 60) & (g < 60) & (b < 60)] = 0
    cls[(r > 200) & (g > 200) & (b > 200)] = 1
    cls[(r > 200) & (g < 80) & (b < 100)] = 2
    cls[(r < 60) & (g < 60) & (b > 150)] = 3
    cls[(r > 200) & (g > 200) & (b < 100)] = 4
    cls[(r > 200) & (g < 80) & (b > 150)] = 5
    cls[(r > 80) & (r < 180) & (g < 110) & (b < 120) & ~((r < 60) & (g < 60) & (b < 60))] = 6
    return cls


def main():
    data = GIF.read_bytes()
    assert hashlib.sha256(data).hexdigest() == SHA
    im = Image.open(GIF)
    assert im.n_frames == 400 and im.size == (300, 300)
    frames = np.zeros((400, 30, 30), np.uint8)
    for k in range(400):
        im.seek(k)
        cls = classify(np.asarray(im.convert("RGB")))
        blocks = cls.reshape(30, 10, 30, 10).transpose(0, 2, 1, 3).reshape(30, 30, 100)
        for y in range(30):
            for x in range(30):
                v = blocks[y, x]
                v = v[v != 255]          # grey arrow pixels / dither noise are unclassified
                cnt = np.bincount(v, minlength=7)
                frames[k, y, x] = int(np.argmax(cnt))
    np.savez_compressed(OUT, frames=frames, gif_sha256=SHA)
    print("wrote", OUT, "non-empty per frame:", sorted(set(int((f >= 2).sum()) for f in frames)))


if __name__ == "__main__":
    sys.exit(main())
