"""Seeded synthetic planar YUV 4:2:0 content.

Per frame n:  Y = clip(128 + 60 sin((x+3n)/17) + 40 cos((y-2n)/11) + N(0, sigma)),
a darker rectangle moving 4 px/frame, U = 128 + 30 sin(x/23) + N(0,2),
V = 128 + 30 cos(y/19) + N(0,2).  8-bit for Main; for Main10 the same signal
scaled by 4 with extra noise, stored as 16-bit little endian.
"""
import numpy as np


def synth_frames(width, height, n_frames, bit_depth=8, seed=1234, sigma=6.0, first_frame=0):
    """Return (Y, U, V) arrays of shape (n, H, W), (n, H/2, W/2) x2; dtype uint8 or uint16.

    Frame k depends only on (seed, first_frame + k), so frame groups generated
    on different ranks are identical to the corresponding slice of one long run.
    """
    assert width % 2 == 0 and height % 2 == 0
    cw, ch = width // 2, height // 2
    scale = 1 << (bit_depth - 8)
    maxv = (1 << bit_depth) - 1
    dt = np.uint8 if bit_depth == 8 else np.uint16
    Y = np.empty((n_frames, height, width), dt)
    U = np.empty((n_frames, ch, cw), dt)
    V = np.empty((n_frames, ch, cw), dt)
    x = np.arange(width, dtype=np.float64)[None, :]
    y = np.arange(height, dtype=np.float64)[:, None]
    xc = np.arange(cw, dtype=np.float64)[None, :]
    yc = np.arange(ch, dtype=np.float64)[:, None]
    rw, rh = max(8, width * 400 // 1920), max(8, height * 600 // 1080)
    for k in range(n_frames):
        n = first_frame + k
        rng = np.random.default_rng([seed, n])
        lum = 128 + 60 * np.sin((x + 3 * n) / 17.0) + 40 * np.cos((y - 2 * n) / 11.0)
        lum = lum + rng.normal(0.0, sigma, (height, width))
        rx = (width // 8 + 4 * n) % max(1, width - rw)
        ry = (height // 8 + 2 * n) % max(1, height - rh)
        lum[ry:ry + rh, rx:rx + rw] = 60 + rng.normal(0.0, 1.0, (rh, rw))
        u = 128 + 30 * np.sin(xc / 23.0) + rng.normal(0.0, 2.0, (ch, cw)) + 0 * yc
        v = 128 + 30 * np.cos(yc / 19.0) + rng.normal(0.0, 2.0, (ch, cw)) + 0 * xc
        if scale > 1:
            lum = lum * scale + rng.normal(0.0, 1.5, lum.shape)
            u = u * scale + rng.normal(0.0, 1.0, u.shape)
            v = v * scale + rng.normal(0.0, 1.0, v.shape)
        Y[k] = np.clip(np.rint(lum), 0, maxv).astype(dt)
        U[k] = np.clip(np.rint(u), 0, maxv).astype(dt)
        V[k] = np.clip(np.rint(v), 0, maxv).astype(dt)
    return Y, U, V


def write_yuv(path, Y, U, V):
    """Planar 4:2:0 file, frame by frame (the layout TVideoIOYuv::read expects;
    reference: source/Lib/TLibVideoIO/TVideoIOYuv.cpp:675)."""
    with open(path, "wb") as f:
        for k in range(Y.shape[0]):
            for p in (Y[k], U[k], V[k]):
                f.write(np.ascontiguousarray(p).astype(p.dtype.newbyteorder("<"), copy=False).tobytes())
