"""A small corpus of BC1 payloads with realistic endpoint statistics for estimator studies: procedural images (fractal
noise, gradients, flat patches, UI-like rectangles, detail maps, white noise) pushed through a simple range-fit BC1
encoder, plus the reference's real-texture fixture and the bench generator."""
import sys
import zlib
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
def fbm(h, w, rng, octaves=5, base=4, persistence=0.5):
    img = np.zeros((h, w))
    amp = 1.0
    for o in range(octaves):
        n = base * 2**o
        g = rng.standard_normal((n + 2, n + 2))
        ys = np.linspace(0, n, h, endpoint=False); xs = np.linspace(0, n, w, endpoint=False)
        y0 = ys.astype(int); x0 = xs.astype(int); fy = (ys - y0)[:, None]; fx = (xs - x0)[None, :]
        fy = fy*fy*(3-2*fy); fx = fx*fx*(3-2*fx)
        a = g[y0][:, x0]; b = g[y0][:, x0+1]; c = g[y0+1][:, x0]; d = g[y0+1][:, x0+1]
        img += amp * ((a*(1-fx)+b*fx)*(1-fy) + (c*(1-fx)+d*fx)*fy)
        amp *= persistence
    return img
def make_image(kind, size, rng):
    h = w = size
    if kind == "fbm":
        ch = [fbm(h, w, rng, persistence=rng.uniform(0.35, 0.7)) for _ in range(3)]
        img = np.stack(ch, -1); img = (img - img.min()) / (np.ptp(img) + 1e-9) * 255
    elif kind == "fbm_gray":
        g = fbm(h, w, rng, persistence=rng.uniform(0.4, 0.7)); g = (g - g.min()) / np.ptp(g)
        tint = rng.uniform(0.5, 1.0, 3)
        img = g[..., None] * tint * 255
    elif kind == "gradient":
        y, x = np.mgrid[0:h, 0:w] / size
        img = np.stack([x * 255, y * 255, (x + y) * 127], -1)
        img += rng.normal(0, rng.uniform(0, 3), img.shape)
    elif kind == "flat_patches":
        n = 8
        pal = rng.integers(0, 256, (n, n, 3))
        img = np.kron(pal, np.ones((h // n, w // n, 1)))
        img += rng.normal(0, 1.0, img.shape)
    elif kind == "ui":
        img = np.full((h, w, 3), 30.0)
        for _ in range(40):
            x0, y0 = rng.integers(0, size - 8, 2); ww, hh = rng.integers(4, size // 3, 2)
            img[y0:y0 + hh, x0:x0 + ww] = rng.integers(0, 256, 3)
    elif kind == "noise":
        img = rng.integers(0, 256, (h, w, 3)).astype(float)
    elif kind == "detail":
        base = fbm(h, w, rng, octaves=3); base = (base - base.min()) / np.ptp(base)
        det = fbm(h, w, rng, octaves=7, base=8, persistence=0.8); det = (det - det.min()) / np.ptp(det)
        col = rng.uniform(0.3, 1.0, 3)
        img = (0.6 * base[..., None] * col + 0.4 * det[..., None]) * 255
    return np.clip(img, 0, 255)
def encode_bc1(img):
    h, w, _ = img.shape
    b = img.reshape(h // 4, 4, w // 4, 4, 3).transpose(0, 2, 1, 3, 4).reshape(-1, 16, 3)
    mx = b.max(1); mn = b.min(1)
    inset = (mx - mn) / 16
    mx = np.clip(mx - inset, 0, 255); mn = np.clip(mn + inset, 0, 255)
    def q(c):
        r = np.round(c[:, 0] * 31 / 255).astype(np.uint32); g = np.round(c[:, 1] * 63 / 255).astype(np.uint32); bb = np.round(c[:, 2] * 31 / 255).astype(np.uint32)
        return (r << 11) | (g << 5) | bb
    c0 = q(mx); c1 = q(mn)
    sw = c0 < c1
    c0, c1 = np.where(sw, c1, c0), np.where(sw, c0, c1)
    def ex(c):
        r = (c >> 11) & 31; g = (c >> 5) & 63; bb = c & 31
        return np.stack([(r << 3) | (r >> 2), (g << 2) | (g >> 4), (bb << 3) | (bb >> 2)], -1).astype(float)
    e0 = ex(c0); e1 = ex(c1)
    pal = np.stack([e0, e1, (2 * e0 + e1) / 3, (e0 + 2 * e1) / 3], 1)  # n,4,3
    d = ((b[:, :, None, :] - pal[:, None, :, :]) ** 2).sum(-1)  # n,16,4
    sel = d.argmin(-1).astype(np.uint32)
    eq = c0 == c1
    sel[eq] = 0
    idx = np.zeros(len(b), np.uint32)
    for p in range(16): idx |= sel[:, p] << (2 * p)
    out = np.zeros((len(b), 8), np.uint8)
    out[:, 0:4] = (c0 | (c1 << 16)).astype(np.uint32).view(np.uint8).reshape(-1, 4)
    out[:, 4:8] = idx.view(np.uint8).reshape(-1, 4)
    return out.reshape(-1)
def corpus(seed=0, sizes=(256, 512, 1024)):
    rng = np.random.default_rng(seed)
    items = []
    for kind in ["fbm", "fbm_gray", "gradient", "flat_patches", "ui", "detail", "noise"]:
        for size in sizes:
            for rep in range(2 if kind != "noise" else 1):
                items.append((f"{kind}-{size}-{rep}", encode_bc1(make_image(kind, size, rng))))
    p = zlib.decompress((ROOT / "tests" / "golden" / "r2-256-bc1.payload.zlib").read_bytes())
    items.append(("r2-256-bc1", np.frombuffer(p, np.uint8).copy()))
    from dxt_lossless_transform_b200 import synth
    for s in range(3):
        items.append((f"synthT-{s}", synth.texture_blocks(1, 1 << 16, seed=100 + s)))
    return items
if __name__ == "__main__":
    c = corpus()
    print(len(c), [(n, len(d)) for n, d in c][:5])
