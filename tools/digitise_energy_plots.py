"""Digitise the published energy plots of the reference (energy_plots/*/*.png, 800x600 CairoMakie figures)
into tests/golden/published_traces.json: KE(t) and ME(t) sampled at whole and half time units.

    python tools/digitise_energy_plots.py [/root/reference]

Method: the panel frame is the 2-px line of colour (127,127,127); the light grid lines sit at the
labelled ticks (sub-pixel centre = darkness-weighted mean over the two anti-aliased rows/columns); the
curve centre in a pixel column is the colour-weighted mean row of the red (KE) / blue (ME) line.
The tick VALUES cannot be read without OCR: they are listed below per figure, read off the images.
Resolution: a grid spacing is ~43 px, the curve centre is good to ~0.2 px, i.e. ~5e-3 of a tick spacing.
The reference tree is not available on the GPU box: the JSON is the committed fixture."""
import json
import sys
from pathlib import Path

import numpy as np
from PIL import Image

ROOT = Path(__file__).resolve().parent.parent
REF = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")

# "formulation/figure" -> panel -> (value of the LOWEST grid line, tick spacing) for the y axis, then (t of the first
# vertical grid line, t tick spacing).  "pe" only where the panel shows the perturbation energy 1/2 g (h - h_i)^2
# (offset: the Jacobian 64^2 figures plot 1/2 g h^2 = 490.5 + PE).
J, D = "jacobian_formulation", "divergence_formulation"
TICKS = {
    f"{J}/64x64_low_B_low_U": {"ke": (0.20, 0.05), "me": (0.15, 0.05), "pe": (490.500 - 490.5, 0.005), "t": (0.0, 5.0)},
    f"{D}/64x64_low_B_low_U": {"ke": (0.20, 0.05), "me": (0.15, 0.05), "pe": (0.0, 0.005), "t": (0.0, 5.0)},
    f"{J}/128x128_low_B_low_U": {"ke": (0.20, 0.10), "me": (0.15, 0.05), "pe": (0.0, 0.01), "t": (0.0, 5.0)},
    f"{D}/128x128_low_B_low_U": {"ke": (0.20, 0.10), "me": (0.15, 0.05), "pe": (0.0, 0.005), "t": (0.0, 5.0)},
    f"{J}/64x64_two_Gaussians_high_B": {"ke": (0.00, 0.02), "me": (0.46, 0.02), "t": (0.0, 10.0)},
    f"{D}/64x64_two_Gaussians_high_B": {"ke": (0.00, 0.02), "me": (0.475, 0.025), "pe": (0.0, 0.002), "t": (0.0, 5.0)},
    f"{J}/128x128_two_Gaussians_high_B": {"ke": (0.00, 0.02), "me": (0.46, 0.02), "pe": (0.0, 0.002), "t": (0.0, 10.0)},
    f"{D}/128x128_two_Gaussians_high_B": {"ke": (0.00, 0.05), "me": (0.50, 0.05), "pe": (0.0, 0.002), "t": (0.0, 10.0)},
    f"{J}/64x64_two_Gaussians_low_B": {"ke": (0.000, 0.001), "me": (0.019, 0.001), "t": (0.0, 25.0)},
    f"{D}/64x64_two_Gaussians_low_B": {"ke": (0.000, 0.001), "me": (0.019, 0.001), "pe": (0.0, 0.00005), "t": (0.0, 10.0)},
    f"{J}/128x128_two_Gaussians_low_B": {"ke": (0.000, 0.001), "me": (0.019, 0.001), "pe": (0.0, 0.00005), "t": (0.0, 10.0)},
    f"{D}/128x128_two_Gaussians_low_B": {"ke": (0.000, 0.001), "me": (0.019, 0.001), "pe": (0.0, 0.00005), "t": (0.0, 10.0)},
}
PANELS = {"ke": (0, 0, "r"), "me": (0, 1, "b"), "pe": (1, 0, "g")}   # (row, column) of the panel, curve colour


def frames(im):
    """(x0, x1, y0, y1) of the panel interiors, found from the gray frame lines."""
    gray = np.all(np.abs(im - 127) <= 2, axis=2)
    rows = np.where(gray.sum(axis=1) > 250)[0]
    cols = np.where(gray.sum(axis=0) > 150)[0]
    def runs(v):
        out, s = [], v[0]
        for a, b in zip(v, v[1:]):
            if b != a + 1:
                out.append((s, a)); s = b
        out.append((s, v[-1]))
        return out
    r, c = runs(list(rows)), runs(list(cols))
    assert len(r) == 4 and len(c) == 4, (r, c)
    return {(i, j): (c[2 * j][1] + 1, c[2 * j + 1][0] - 1, r[2 * i][1] + 1, r[2 * i + 1][0] - 1) for i in (0, 1) for j in (0, 1)}


def grid_lines(sub, axis):
    """Sub-pixel positions of the light grid lines of a panel interior along `axis` (0: rows, 1: columns)."""
    lum = sub.sum(axis=2) / 3.0
    is_gray = (np.ptp(sub, axis=2) <= 3) & (lum < 252) & (lum > 200)
    frac = is_gray.mean(axis=1 - axis)
    dark = np.where(is_gray, 255.0 - lum, 0.0).sum(axis=1 - axis)
    idx = np.where(frac > 0.5)[0]
    out, k = [], 0
    while k < len(idx):
        j = k
        while j + 1 < len(idx) and idx[j + 1] == idx[j] + 1:
            j += 1
        sel = idx[k:j + 1]
        out.append(float((sel * dark[sel]).sum() / dark[sel].sum()))
        k = j + 1
    return out


def curve(sub, colour):
    """Centre row of the coloured line in every pixel column (nan where absent)."""
    r, g, b = sub[..., 0].astype(float), sub[..., 1].astype(float), sub[..., 2].astype(float)
    w = np.clip({"r": r - np.maximum(g, b), "b": b - np.maximum(r, g), "g": g - np.maximum(r, b)}[colour], 0, None)
    w[w < 40] = 0.0
    rows = np.arange(sub.shape[0])[:, None]
    s = w.sum(axis=0)
    with np.errstate(invalid="ignore", divide="ignore"):
        c = (w * rows).sum(axis=0) / s
    c[s < 200] = np.nan
    return c


def digitise(png, ticks):
    im = np.array(Image.open(png).convert("RGB")).astype(int)
    fr = frames(im)
    out = {}
    for name, (row, col, colour) in PANELS.items():
        if name not in ticks:
            continue
        x0, x1, y0, y1 = fr[(row, col)]
        sub = im[y0:y1 + 1, x0:x1 + 1]
        gy, gx = grid_lines(sub, 0), grid_lines(sub, 1)
        (ylow, dy), (tlow, dt) = ticks[name], ticks["t"]
        py = (gy[-1] - gy[0]) / (len(gy) - 1)               # pixels per y tick (rows grow downwards)
        px = (gx[-1] - gx[0]) / (len(gx) - 1)
        c = curve(sub, colour)
        cols = np.arange(sub.shape[1])
        t = tlow + (cols - gx[0]) / px * dt
        val = ylow + (gy[-1] - c) / py * dy
        ok = ~np.isnan(val)
        ts = np.arange(0.5, t[ok].max() - 0.2, 0.5)
        out[name] = {"t": [float(x) for x in ts], "v": [float(np.interp(x, t[ok], val[ok])) for x in ts],
                     "resolution": float(0.3 * dy / py)}
    return out


if __name__ == "__main__":
    res = {key: digitise(REF / "energy_plots" / f"{key}.png", ticks) for key, ticks in TICKS.items()}
    dst = ROOT / "tests" / "golden" / "published_traces.json"
    dst.write_text(json.dumps(res, indent=1))
    for k, v in res.items():
        print(k, {n: ([round(x, 6) for x in p["v"][:2]], "...", round(p["v"][-1], 6), "T", p["t"][-1], "res", round(p["resolution"], 7)) for n, p in v.items()})
