"""Seeded synthetic inputs for the BASELINE.json configs.

Everything is generated on the host from numpy PCG64 seeds; the same arrays feed the
CUDA path and (in tests / the CPU baseline) the oracle.
"""
import numpy as np
from scipy import ndimage


def _noise_image(w, h, seed, sigma=3.0):
    rng = np.random.Generator(np.random.PCG64(seed))
    img = rng.integers(0, 256, size=(h, w, 3)).astype(np.float32)
    for c in range(3):
        img[..., c] = ndimage.gaussian_filter(img[..., c], sigma, mode="reflect")
    lo, hi = img.min(), img.max()
    img = 16.0 + (img - lo) * (224.0 / max(hi - lo, 1e-6))
    return img


def smooth_warp(w, h, seed, amp):
    """Sum of 3 low-frequency sinusoids per component, |field| <= amp px.  Returns (h,w,2) float32."""
    rng = np.random.Generator(np.random.PCG64(seed))
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    out = np.zeros((h, w, 2), np.float32)
    for c in range(2):
        acc = np.zeros((h, w), np.float32)
        for _ in range(3):
            fx, fy = rng.uniform(0.5, 2.0, 2)
            ph = rng.uniform(0, 2 * np.pi)
            acc += np.sin(2 * np.pi * (fx * xx / w + fy * yy / h) + ph).astype(np.float32)
        out[..., c] = acc * (amp / 3.0)
    return out


def warp_image(img, field):
    """img1(p) = img0(p - field(p)) (so that the halfway vector is ~ field/2)."""
    h, w, _ = img.shape
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    out = np.empty_like(img)
    for c in range(3):
        out[..., c] = ndimage.map_coordinates(img[..., c], [yy - field[..., 1], xx - field[..., 0]], order=1, mode="nearest")
    return out


def image_pair(w, h, seed_img, seed_warp, amp, noise=0.02):
    """Returns rgb0, rgb1 as (1,h,w,3) uint8 and the true warp (h,w,2)."""
    img0 = _noise_image(w, h, seed_img)
    field = smooth_warp(w, h, seed_warp, amp)
    img1 = warp_image(img0, field)
    rng = np.random.Generator(np.random.PCG64(seed_warp + 7))
    img1 = img1 + rng.normal(0, noise * 255.0, img1.shape).astype(np.float32)
    to8 = lambda a: np.clip(np.rint(a), 0, 255).astype(np.uint8)[None]
    return to8(img0), to8(img1), field


def point_pairs(n, w, h, seed, field, margin=32):
    """n UI point pairs: lp uniform in [margin, w-margin) x [margin, h-margin); rp = lp + warp(lp) rounded.
    Returns (lp, lw, rp, rw) in the resolved-connection layout: (n,4) int32 [x,y,frame,keyflag], (n,) float32."""
    rng = np.random.Generator(np.random.PCG64(seed))
    lx = rng.integers(margin, w - margin, n)
    ly = rng.integers(margin, h - margin, n)
    rx = np.clip(np.rint(lx + field[ly, lx, 0]), 0, w - 1).astype(np.int32)
    ry = np.clip(np.rint(ly + field[ly, lx, 1]), 0, h - 1).astype(np.int32)
    lp = np.stack([lx, ly, np.zeros(n, np.int64), np.ones(n, np.int64)], 1).astype(np.int32)
    rp = np.stack([rx, ry, np.zeros(n, np.int64), np.ones(n, np.int64)], 1).astype(np.int32)
    return lp, np.ones(n, np.float32), rp, np.ones(n, np.float32)


def extended_rgba(rgb, ex):
    """Pyramid::_extends (pyramid.cu:186-200): white opaque border of ex px, alpha 0 inside. rgb: (h,w,3) u8."""
    h, w, _ = rgb.shape
    out = np.full((h + 2 * ex, w + 2 * ex, 4), 255, np.uint8)
    out[ex:ex + h, ex:ex + w, :3] = rgb
    out[ex:ex + h, ex:ex + w, 3] = 0
    return out


def smoothstep(t):
    """RenderWidget::SmoothStep (UI/RenderWidget.cpp:268-273) with a=0,b=1."""
    t = np.float32(t)
    if t < 0:
        return np.float32(0)
    if t > 1:
        return np.float32(1)
    return np.float32(t * t * (np.float32(3) - np.float32(2) * t))


def video_offset_fn(seed_warp, vel=1.5, wobble=1.0):
    """Camera offset (x, y) of frame t - d/2 of video_pair's synthetic motion."""
    rng = np.random.Generator(np.random.PCG64(seed_warp + 11))
    ph = rng.uniform(0, 2 * np.pi)

    def offset(t):
        return np.float32(vel * t + wobble * np.sin(0.2 * t + ph)), np.float32(wobble * np.cos(0.15 * t + ph))
    return offset


def video_tracks(w, h, d, seed, seed_warp, field, ntracks=4, margin=96, vel=1.5, wobble=1.0):
    """`ntracks` UI point tracks propagated by the analytic motion of video_pair, all frames connected.
    Track k follows one scene point: left point in video 0 at frame t, right point = left + warp (rounded), weight 1.
    Returns the resolved connections (lp, lw, rp, rw): (ntracks*d, 4) int32 [x, y, frame, keyflag] and weights."""
    rng = np.random.Generator(np.random.PCG64(seed))
    offset = video_offset_fn(seed_warp, vel, wobble)
    bx = rng.integers(margin, w - margin, ntracks)
    by = rng.integers(margin, h - margin, ntracks)
    lp, rp = [], []
    for k in range(ntracks):
        fx, fy = field[by[k], bx[k]]
        for t in range(d):
            ox, oy = offset(t - d / 2)
            ox0, oy0 = offset(0 - d / 2 + d // 2)                      # the point is picked in the middle frame
            x = int(np.clip(np.rint(bx[k] + (ox - ox0)), 0, w - 1)); y = int(np.clip(np.rint(by[k] + (oy - oy0)), 0, h - 1))
            lp.append((x, y, t, 1 if t == d // 2 else 0))
            rp.append((int(np.clip(np.rint(x + fx), 0, w - 1)), int(np.clip(np.rint(y + fy), 0, h - 1)), t, 1 if t == d // 2 else 0))
    n = len(lp)
    return np.asarray(lp, np.int32), np.ones(n, np.float32), np.asarray(rp, np.int32), np.ones(n, np.float32)


def video_pair(w, h, d, seed_img, seed_warp, amp, vel=1.5, wobble=1.0):
    """Two videos (d,h,w,3) u8 plus analytic forward/backward flows (d,h,w,2) float32 for each.

    Video 0 frame t = base image translated by vel*t px in x plus a small smooth wobble; video 1 = the same motion
    applied to the warped base.  Forward flow of frame t maps t -> t+1 (zero for the last frame), backward flow maps
    t -> t-1 (zero for frame 0), as UI/MdiEditor.cpp:1637-1641,1668-1672 leaves them; clamped to +-50.
    """
    base0 = _noise_image(w + 2 * 64, h + 2 * 64, seed_img)
    field = smooth_warp(w + 128, h + 128, seed_warp, amp)
    base1 = warp_image(base0, field)
    offset = video_offset_fn(seed_warp, vel, wobble)

    v0 = np.zeros((d, h, w, 3), np.uint8)
    v1 = np.zeros((d, h, w, 3), np.uint8)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    for t in range(d):
        ox, oy = offset(t - d / 2)
        for src, dst in ((base0, v0), (base1, v1)):
            for c in range(3):
                dst[t, ..., c] = np.clip(np.rint(ndimage.map_coordinates(
                    src[..., c], [yy + 64 - oy, xx + 64 - ox], order=1, mode="nearest")), 0, 255)
    f = np.zeros((d, h, w, 2), np.float32)
    b = np.zeros((d, h, w, 2), np.float32)
    for t in range(d):
        ox, oy = offset(t - d / 2)
        if t + 1 < d:
            nx, ny = offset(t + 1 - d / 2)
            f[t, ..., 0] = nx - ox
            f[t, ..., 1] = ny - oy
        if t > 0:
            px, py = offset(t - 1 - d / 2)
            b[t, ..., 0] = px - ox
            b[t, ..., 1] = py - oy
    f = np.clip(f, -50, 50)
    b = np.clip(b, -50, 50)
    return v0, v1, (f, f.copy(), b, b.copy()), field[64:64 + h, 64:64 + w]


def video_pair_shift(w, h, d, seed_img, seed_warp, amp, vel=2):
    """A cheap large video pair (cfg5 probes): frame t = the base image shifted by vel * (t - d // 2) WHOLE pixels in x (plain
    slicing, no interpolation), video 1 = the same motion of the warped base; flows are the exact constant shift (zero
    for the last / first frame like UI/MdiEditor.cpp:1637-1641,1668-1672)."""
    pad = vel * (d // 2 + 1)
    base0 = _noise_image(w + 2 * pad, h, seed_img)
    field = smooth_warp(w + 2 * pad, h, seed_warp, amp)
    base1 = warp_image(base0, field)
    to8 = lambda a: np.clip(np.rint(a), 0, 255).astype(np.uint8)
    b0, b1 = to8(base0), to8(base1)
    v0 = np.empty((d, h, w, 3), np.uint8); v1 = np.empty((d, h, w, 3), np.uint8)
    for t in range(d):
        o = pad - vel * (t - d // 2)                     # content moves by +vel px per frame
        v0[t] = b0[:, o:o + w]; v1[t] = b1[:, o:o + w]
    f = np.zeros((d, h, w, 2), np.float32); b = np.zeros((d, h, w, 2), np.float32)
    f[:-1, ..., 0] = vel
    b[1:, ..., 0] = -vel
    return v0, v1, (f, f, b, b), field[:, pad:pad + w]


CONFIGS = {
    # name: (w, h, d, seed_img, seed_warp, amp)
    "cfg1": (256, 256, 1, 1001, 1002, 6.0),
    "cfg2": (512, 512, 1, 2001, 2002, 12.0),
    "cfg3": (1920, 1080, 1, 3001, 3002, 24.0),
    "cfg4": (1280, 720, 120, 4001, 4002, 8.0),
    "cfg5": (3840, 2160, 240, 5001, 5002, 16.0),
}
