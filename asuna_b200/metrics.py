"""Image-comparison metrics for the parity thresholds of BASELINE.json: mean relative error and
an in-repo implementation of LDR-FLIP (Andersson et al. 2020, "FLIP: A Difference Evaluator for
Alternating Images") -- `flip_evaluator` is not installed and there is no network.

FLIP here follows the published algorithm: colour pipeline (YCxCz opponent space -> contrast
sensitivity filtering -> Hunt adjustment -> HyAB distance -> redistribution) and feature pipeline
(edge / point detection on the achromatic channel), combined as dE_c ** (1 - dE_f).
"""
import numpy as np
from scipy.ndimage import convolve

_PPD_DEFAULT = 67.0  # 0.7 m from a 0.7 m wide 3840-pixel display: the paper's default


def tonemap_for_flip(hdr):
    """Radiance -> [0,1] sRGB-ish LDR.  Per-sample radiance is already clamped to [0,10] by the
    integrator (rgen:147); a Reinhard curve keeps highlights comparable."""
    x = np.clip(np.asarray(hdr, np.float64)[..., :3], 0.0, None)
    x = x / (1.0 + x)
    return np.clip(x, 0.0, 1.0) ** (1.0 / 2.2)


def mean_relative_error(test, ref):
    t = np.asarray(test, np.float64)[..., :3]
    r = np.asarray(ref, np.float64)[..., :3]
    return float(np.abs(t - r).mean() / max(np.abs(r).mean(), 1e-12))


def _srgb_to_linear(c):
    return np.where(c <= 0.04045, c / 12.92, ((c + 0.055) / 1.055) ** 2.4)


_RGB2XYZ = np.array([[10135552 / 24577794, 8788810 / 24577794, 4435075 / 24577794],
                     [2613072 / 12288897, 8788810 / 12288897, 887015 / 12288897],
                     [1425312 / 73733382, 8788810 / 73733382, 70074185 / 73733382]])
_WHITE = _RGB2XYZ @ np.ones(3)


def _linrgb_to_ycxcz(rgb):
    xyz = rgb @ _RGB2XYZ.T
    xyz = xyz / _WHITE
    y = 116.0 * xyz[..., 1] - 16.0
    cx = 500.0 * (xyz[..., 0] - xyz[..., 1])
    cz = 200.0 * (xyz[..., 1] - xyz[..., 2])
    return np.stack([y, cx, cz], -1)


def _ycxcz_to_linrgb(ycc):
    y = (ycc[..., 0] + 16.0) / 116.0
    x = ycc[..., 1] / 500.0 + y
    z = y - ycc[..., 2] / 200.0
    xyz = np.stack([x, y, z], -1) * _WHITE
    return xyz @ np.linalg.inv(_RGB2XYZ).T


def _xyz_to_lab(xyz):
    d = 6.0 / 29.0
    t = xyz / _WHITE
    f = np.where(t > d ** 3, np.cbrt(t), t / (3 * d * d) + 4.0 / 29.0)
    return np.stack([116.0 * f[..., 1] - 16.0, 500.0 * (f[..., 0] - f[..., 1]), 200.0 * (f[..., 1] - f[..., 2])], -1)


def _linrgb_to_lab(rgb):
    return _xyz_to_lab(rgb @ _RGB2XYZ.T)


def _csf_kernels(ppd):
    a1 = {"A": (1.0, 0.0047, 0.0, 1e-5), "RG": (1.0, 0.0053, 0.0, 1e-5), "BY": (34.1, 0.04, 13.5, 0.025)}
    b_max = max(a1["A"][1], a1["A"][3], a1["RG"][1], a1["RG"][3], a1["BY"][1], a1["BY"][3])
    r = int(np.ceil(3.0 * np.sqrt(b_max / (2.0 * np.pi ** 2)) * ppd))
    xs = np.arange(-r, r + 1) / ppd
    X, Y = np.meshgrid(xs, xs)
    d2 = X * X + Y * Y
    ks = []
    for key in ("A", "RG", "BY"):
        aa1, bb1, aa2, bb2 = a1[key]
        k = aa1 * np.sqrt(np.pi / bb1) * np.exp(-np.pi ** 2 * d2 / bb1) + aa2 * np.sqrt(np.pi / bb2) * np.exp(-np.pi ** 2 * d2 / bb2)
        ks.append(k / k.sum())
    return ks


def _spatial_filter(ycc, ppd):
    ks = _csf_kernels(ppd)
    out = np.stack([convolve(ycc[..., c], ks[c], mode="nearest") for c in range(3)], -1)
    return np.clip(_ycxcz_to_linrgb(out), 0.0, 1.0)


def _hunt(lab):
    out = lab.copy()
    out[..., 1] = 0.01 * lab[..., 0] * lab[..., 1]
    out[..., 2] = 0.01 * lab[..., 0] * lab[..., 2]
    return out


def _hyab(a, b):
    return np.abs(a[..., 0] - b[..., 0]) + np.sqrt((a[..., 1] - b[..., 1]) ** 2 + (a[..., 2] - b[..., 2]) ** 2)


def _redistribute(d, cmax, pc=0.4, pt=0.95):
    lim = pc * cmax
    return np.where(d < lim, pt / lim * d, pt + (d - lim) / (cmax - lim) * (1.0 - pt))


def _feature_kernels(ppd):
    w = 0.082
    sd = 0.5 * w * ppd
    r = int(np.ceil(3 * sd))
    xs = np.arange(-r, r + 1)
    X, Y = np.meshgrid(xs, xs)
    g = np.exp(-(X * X + Y * Y) / (2 * sd * sd))
    edge = -X * g
    edge_pos = edge * (edge > 0)
    edge_neg = -edge * (edge < 0)
    edge = edge_pos / edge_pos.sum() - edge_neg / edge_neg.sum()
    point = (X * X / (sd * sd) - 1) * g
    point_pos = point * (point > 0)
    point_neg = -point * (point < 0)
    point = point_pos / point_pos.sum() - point_neg / point_neg.sum()
    return edge, point


def _features(y, ppd):
    edge, point = _feature_kernels(ppd)
    ex, ey = convolve(y, edge, mode="nearest"), convolve(y, edge.T, mode="nearest")
    px, py = convolve(y, point, mode="nearest"), convolve(y, point.T, mode="nearest")
    return np.sqrt(ex * ex + ey * ey), np.sqrt(px * px + py * py)


def flip_map(test_ldr, ref_ldr, ppd=_PPD_DEFAULT):
    """Per-pixel LDR-FLIP error in [0,1] between two sRGB images with values in [0,1]."""
    t = _linrgb_to_ycxcz(_srgb_to_linear(np.clip(np.asarray(test_ldr, np.float64)[..., :3], 0, 1)))
    r = _linrgb_to_ycxcz(_srgb_to_linear(np.clip(np.asarray(ref_ldr, np.float64)[..., :3], 0, 1)))
    # colour pipeline
    tf, rf = _hunt(_linrgb_to_lab(_spatial_filter(t, ppd))), _hunt(_linrgb_to_lab(_spatial_filter(r, ppd)))
    green = _hunt(_linrgb_to_lab(np.array([[[0.0, 1.0, 0.0]]])))
    blue = _hunt(_linrgb_to_lab(np.array([[[0.0, 0.0, 1.0]]])))
    qc = 0.7
    cmax = _hyab(green, blue)[0, 0] ** qc
    de_c = _redistribute(_hyab(tf, rf) ** qc, cmax)
    # feature pipeline on normalised achromatic channel
    yt, yr = (t[..., 0] + 16.0) / 116.0, (r[..., 0] + 16.0) / 116.0
    et, pt_ = _features(yt, ppd)
    er, pr = _features(yr, ppd)
    qf = 0.5
    de_f = (np.maximum(np.abs(et - er), np.abs(pt_ - pr)) / np.sqrt(2.0)) ** qf
    return np.clip(de_c ** (1.0 - de_f), 0.0, 1.0)


def flip(test_hdr, ref_hdr, ppd=_PPD_DEFAULT):
    """Mean LDR-FLIP between two radiance images after the shared tone map."""
    return float(flip_map(tonemap_for_flip(test_hdr), tonemap_for_flip(ref_hdr), ppd).mean())
