"""TEST INFRASTRUCTURE ONLY -- digitises the one output of the REAL reference stack that ships with the reference:
`docs/resources/controller_comparison_plot.png`, the figure `comparison/controller_comparison_plot.py:60-186` drew from
runs of `comparison/controller_comparison.py` (Genesis physics, rsl_rl PPO policy, cvxpy/OSQP data-driven MPC, the SE(3)
tracking controller; `configs/comparison/controller_comparison_config.yaml`: hover at [0, 0, 1.5], setpoints
[1, -1, 2.5] and [0, 0, 1.5], 175 steps of 0.04 s each).  Neither Genesis nor cvxpy can be installed here, so this
figure is the only place where numbers produced by them can be read: the three "System Outputs" axes give x(t), y(t),
z(t) of the tracking-controller drone (red, dash-dot), the PPO drone (green) and the mean of the data-driven MPC runs
(goldenrod, with its +/- one-standard-deviation band).

    python oracle/digitize_reference_plot.py      ->  tests/golden/ref_comparison_plot_curves.npz

Method: the axes frames and tick marks are located in the bitmap (frame = t in [0, 14] s exactly; y ticks at known
values), every pixel whose colour is the pure line colour of a controller (anti-aliased rim pixels are excluded) and
that lies outside the legend box becomes a point (t, value).  Resolution: 1 px = 0.047 s (1.17 control steps) and
6.4 / 6.8 / 5.8 mm in x / y / z; the lines are 2 px wide, so a curve that reproduces the reference passes within about
1.5 px of every point.
"""

import os

import numpy as np
from PIL import Image

REF_PNG = "/root/reference/docs/resources/controller_comparison_plot.png"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ref_comparison_plot_curves.npz")

# frame (left, right, top, bottom) in pixels; two y ticks (pixel, value); legend box (x0, x1, y0, y1) - located with
# the scans documented in DESIGN.md section 2 (dark frame lines, tick marks left of / below the frames, legend edges)
AXES = {
    "x": dict(frame=(80, 378, 476, 739), ticks=((719, 0.0), (483, 1.5)), legend=(212, 371, 567, 648)),
    "y": dict(frame=(493, 791, 476, 739), ticks=((513, 0.0), (735, -1.5)), legend=(625, 784, 567, 648)),
    "z": dict(frame=(881, 1179, 476, 739), ticks=((706, 1.5), (490, 2.75)), legend=(888, 1050, 651, 732)),
    # "Control Inputs" row: total thrust [N], roll and pitch rate commands [rad/s] (units per pixel, not metres)
    "thrust": dict(frame=(70, 374, 89, 359), ticks=((346, 0.0), (115, 0.5)), legend=(208, 367, 276, 352)),
    "roll": dict(frame=(473, 777, 89, 359), ticks=((346, -15.0), (101, 15.0)), legend=(611, 770, 96, 171)),
    "pitch": dict(frame=(875, 1179, 89, 359), ticks=((346, -15.0), (101, 15.0)), legend=(882, 1041, 276, 352)),
}
T_END = 14.0  # the frame spans t = 0 .. 14 s (x ticks at every second from the left to the right frame line)


def classify(im):
    r, g, b = im[..., 0], im[..., 1], im[..., 2]
    return {
        "tracking": (r > 215) & (g < 90) & (b < 90),                                  # red (255, 0, 0)
        "rl": (r < 60) & (g > 100) & (g < 160) & (b < 60),                            # green (0, 128, 0)
        "dd_mpc": (abs(r - 218) < 14) & (abs(g - 165) < 14) & (b < 70),               # goldenrod (218, 165, 32)
        "band": (abs(r - 243) < 7) & (abs(g - 227) < 7) & (abs(b - 187) < 9),         # goldenrod at alpha 0.3 on white
    }


def main():
    im = np.array(Image.open(REF_PNG).convert("RGB")).astype(int)
    cls = classify(im)
    out = {}
    for name, ax in AXES.items():
        l, r, t, b = ax["frame"]
        (p0, v0), (p1, v1) = ax["ticks"]
        per_px = (v1 - v0) / (p1 - p0)
        lx0, lx1, ly0, ly1 = ax["legend"]
        out[f"{name}_px_per_s"] = (r - l) / T_END
        out[f"{name}_px_per_m"] = abs(1.0 / per_px)
        for c, mask in cls.items():
            ys, xs = np.nonzero(mask[t + 1 : b, l + 1 : r])
            ys, xs = ys + t + 1, xs + l + 1
            keep = ~((xs >= lx0 - 1) & (xs <= lx1 + 1) & (ys >= ly0 - 1) & (ys <= ly1 + 1))
            xs, ys = xs[keep], ys[keep]
            tt = (xs - l) / (r - l) * T_END
            vv = v0 + (ys - p0) * per_px
            out[f"{name}_{c}"] = np.stack([tt, vv], 1)
    np.savez_compressed(OUT, **out)
    for k, v in out.items():
        print(k, v.shape if hasattr(v, "shape") and v.shape else v)


if __name__ == "__main__":
    main()
