"""Host-side spectral curves: restatement of the `math::Curve` surface the PT path touches.

The reference evaluates `Curve` / `CurveWithCDF` objects from the un-vendored crate
`math` (github.com/gillett-hernandez/rust_cg_math, git dependency with no rev pinned:
reference Cargo.toml:52-55). That source is NOT under /root/reference, so everything here is
a restatement of the crate's published behaviour — **parity unpinned** (SURVEY.md Appendix B).
The device never evaluates a `Curve`: every curve is baked into a uniform LUT here
(include/rpt.h `curve_lut`), and a Rust host would bake the same LUT by calling the real
`evaluate_power`.

Constructors mirror reference src/parsing/curves.rs:298-372 (CurveData -> Curve) and
src/curves.rs:7-78.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

# math::spectral::{BOUNDED_VISIBLE_RANGE, EXTENDED_VISIBLE_RANGE} (values recalled; the
# reference's data/config.toml:23 comment "wavelength_bounds = [380.0, 750.0]" agrees).
BOUNDED_VISIBLE_RANGE = (380.0, 750.0)
EXTENDED_VISIBLE_RANGE = (370.0, 790.0)

F32 = np.float32


def _f32(x) -> np.ndarray:
    return np.asarray(x, dtype=F32)


def _interp_weights(t: np.ndarray, mode: str) -> Tuple[np.ndarray, np.ndarray]:
    """InterpolationMode::{Linear,Nearest,Cubic} weights (left, right) for parameter t."""
    if mode == "Linear":
        return (F32(1.0) - t, t)
    if mode == "Nearest":
        right = (t >= F32(0.5)).astype(F32)
        return (F32(1.0) - right, right)
    if mode == "Cubic":
        t2 = F32(2.0) * t
        one_sub_t = F32(1.0) - t
        h00 = (F32(1.0) + t2) * one_sub_t * one_sub_t
        h01 = t * t * (F32(3.0) - t2)
        return (h00, h01)
    raise ValueError(f"unknown interpolation mode {mode}")


class Curve:
    """Base class; `evaluate(lam)` takes/returns float32 numpy arrays (vectorised)."""

    def evaluate(self, lam: np.ndarray) -> np.ndarray:  # pragma: no cover - abstract
        raise NotImplementedError

    # SpectralPowerDistributionFunction<f32>
    def evaluate_power(self, lam) -> np.ndarray:
        return self.evaluate(_f32(lam))

    def evaluate_clamped(self, lam) -> np.ndarray:
        return np.clip(self.evaluate(_f32(lam)), F32(0.0), F32(1.0))

    def evaluate_integral(self, bounds: Tuple[float, float], samples: int, clamped: bool = False) -> float:
        """Riemann sum with `samples` left-edge samples (math::Curve::evaluate_integral)."""
        lo, hi = bounds
        step = (hi - lo) / samples
        lam = _f32(lo + step * np.arange(samples, dtype=np.float64))
        v = self.evaluate_clamped(lam) if clamped else self.evaluate(lam)
        return float(np.sum(v.astype(np.float64)) * step)

    def to_cdf(self, bounds: Tuple[float, float], resolution: int) -> "CurveWithCDF":
        return CurveWithCDF.from_curve(self, bounds, resolution)


@dataclass
class Const(Curve):
    value: float

    def evaluate(self, lam):
        return np.full_like(_f32(lam), max(self.value, 0.0))


@dataclass
class Linear(Curve):
    """Uniformly spaced samples over `bounds` (Curve::Linear)."""

    signal: np.ndarray
    bounds: Tuple[float, float]
    mode: str = "Linear"

    def evaluate(self, lam):
        lam = _f32(lam)
        sig = _f32(self.signal)
        n = len(sig)
        lo, hi = F32(self.bounds[0]), F32(self.bounds[1])
        step = (hi - lo) / F32(n)
        inside = (lam >= lo) & (lam <= hi)
        x = np.where(inside, lam, lo)
        idx = np.floor((x - lo) / step).astype(np.int64)
        idx = np.clip(idx, 0, n - 1)
        has_right = idx + 1 < n
        left = sig[idx]
        right = sig[np.minimum(idx + 1, n - 1)]
        t = (x - (lo + idx.astype(F32) * step)) / step
        wl, wr = _interp_weights(t.astype(F32), self.mode)
        val = np.where(has_right, wl * left + wr * right, left)
        return np.where(inside, val, F32(0.0)).astype(F32)


@dataclass
class Tabulated(Curve):
    """(x, y) pairs sorted by x (Curve::Tabulated); clamps to the end values outside."""

    xs: np.ndarray
    ys: np.ndarray
    mode: str = "Linear"

    def evaluate(self, lam):
        lam = _f32(lam)
        xs, ys = _f32(self.xs), _f32(self.ys)
        n = len(xs)
        # binary_search: Ok(i) -> i (exact match), Err(i) -> insertion point
        idx = np.searchsorted(xs, lam, side="left")
        at_end = idx >= n
        at_start = idx == 0
        i = np.clip(idx, 1, n - 1)
        lx, rx = xs[i - 1], xs[i]
        ly, ry = ys[i - 1], ys[i]
        denom = np.where(rx == lx, F32(1.0), rx - lx)
        t = ((lam - lx) / denom).astype(F32)
        wl, wr = _interp_weights(t, self.mode)
        val = wl * ly + wr * ry
        val = np.where(at_end, ys[n - 1], val)
        val = np.where(at_start, ys[0], val)
        return val.astype(F32)


@dataclass
class Cauchy(Curve):
    a: float
    b: float

    def evaluate(self, lam):
        lam = _f32(lam)
        return (F32(self.a) + F32(self.b) / (lam * lam)).astype(F32)


def _gaussian(x, alpha, mu, sigma1, sigma2):
    s = (x - F32(mu)) / np.where(x < F32(mu), F32(sigma1), F32(sigma2))
    return F32(alpha) * np.exp(-(s * s) / F32(2.0))


@dataclass
class Exponential(Curve):
    """Sum of asymmetric Gaussians (offset, sigma1, sigma2, multiplier)."""

    signal: List[Tuple[float, float, float, float]]

    def evaluate(self, lam):
        lam = _f32(lam)
        val = np.zeros_like(lam)
        for (mu, s1, s2, mult) in self.signal:
            val = val + _gaussian(lam, mult, mu, s1, s2)
        return val.astype(F32)


@dataclass
class InverseExponential(Curve):
    signal: List[Tuple[float, float, float, float]]

    def evaluate(self, lam):
        lam = _f32(lam)
        val = np.ones_like(lam)
        for (mu, s1, s2, mult) in self.signal:
            val = val - _gaussian(lam, mult, mu, s1, s2)
        return np.maximum(val, F32(0.0)).astype(F32)


_HCC2 = 1.1910429723971884140794892e-29
_HKC = 1.438777085924334052222404423195819240925e-2


def _blackbody(temperature: float, lam_nm: np.ndarray) -> np.ndarray:
    lam = lam_nm.astype(np.float64) * 1e-9
    return (lam ** -5.0) * _HCC2 / (np.exp(_HKC / (lam * temperature)) - 1.0)


@dataclass
class Blackbody(Curve):
    temperature: float
    boost: float

    def evaluate(self, lam):
        lam = _f32(lam)
        bb = _blackbody(self.temperature, lam)
        if self.boost == 0.0:
            return bb.astype(F32)
        peak_lambda = np.array([2.8977721e-3 / (self.temperature * 1e-9)])
        return (self.boost * bb / _blackbody(self.temperature, peak_lambda)[0]).astype(F32)


@dataclass
class Machine(Curve):
    """seed (op curve)* evaluated left to right, result clamped at 0 (Curve::Machine)."""

    seed: float
    ops: List[Tuple[str, Curve]] = field(default_factory=list)

    def evaluate(self, lam):
        lam = _f32(lam)
        val = np.full_like(lam, F32(self.seed))
        for op, c in self.ops:
            e = c.evaluate(lam)
            val = val + e if op == "Add" else val * e
        return np.maximum(val, F32(0.0)).astype(F32)


# ---- CurveWithCDF -------------------------------------------------------------------------


@dataclass
class CurveWithCDF:
    """pdf curve + tabulated cdf (Curve::Linear) + integral, as built by Curve::to_cdf."""

    pdf: Curve
    cdf_signal: np.ndarray
    cdf_bounds: Tuple[float, float]
    cdf_mode: str
    pdf_integral: float

    @staticmethod
    def from_curve(curve: Curve, bounds: Tuple[float, float], resolution: int) -> "CurveWithCDF":
        if isinstance(curve, Linear):
            sig = _f32(curve.signal).astype(np.float64)
            b = curve.bounds
            step = (b[1] - b[0]) / len(sig)
            cdf = np.cumsum(sig * step)
            total = float(cdf[-1]) if len(cdf) else 0.0
            cdf = cdf / total if total != 0.0 else cdf
            return CurveWithCDF(curve, cdf.astype(F32), b, curve.mode, total)
        lo, hi = bounds
        step = (hi - lo) / resolution
        lam = _f32(lo + step * np.arange(resolution, dtype=np.float64))
        v = curve.evaluate(lam).astype(np.float64)
        cdf = np.cumsum(v * step)
        total = float(cdf[-1])
        cdf = cdf / total if total != 0.0 else cdf
        return CurveWithCDF(curve, cdf.astype(F32), bounds, "Linear", total)

    def evaluate_power(self, lam):
        return self.pdf.evaluate_power(lam)


# ---- constructors mirroring parsing/curves.rs ---------------------------------------------


def _domain_funcs(domain_mapping: Optional[dict]):
    dm = domain_mapping or {}
    xo, xs = dm.get("x_offset", 0.0) or 0.0, dm.get("x_scale", 1.0)
    yo, ys = dm.get("y_offset", 0.0) or 0.0, dm.get("y_scale", 1.0)
    xs = 1.0 if xs is None else xs
    ys = 1.0 if ys is None else ys
    return (lambda x: (F32(x) - F32(xo)) * F32(xs)), (lambda y: (F32(y) - F32(yo)) * F32(ys))


def parse_tabulated_csv(text: str, column: int, mode: str, fx, fy) -> Tabulated:
    """reference src/parsing/curves.rs:136-173: x = first field, y = `column`-th field after it."""
    xs, ys = [], []
    for line in text.split("\n"):
        if line == "":
            continue
        parts = line.split(",")
        if len(parts) <= column:
            continue
        try:
            x = float(parts[0].strip())
            y = float(parts[column].strip())
        except ValueError:
            continue
        xs.append(fx(x))
        ys.append(fy(y))
    return Tabulated(_f32(xs), _f32(ys), mode)


def parse_linear(text: str, mode: str, fx, fy) -> Linear:
    """reference src/parsing/curves.rs:175-213: first line `start_x, step`, then one value per line."""
    lines = [l for l in text.split("\n")]
    if lines and lines[-1] == "":
        lines = lines[:-1]
    first = lines[0].split(",")
    start_x, step = float(first[0].strip()), float(first[1].strip())
    values = [fy(float(l.strip())) for l in lines[1:]]
    end_x = start_x + step * len(values)
    return Linear(_f32(values), (float(fx(start_x)), float(fx(end_x))), mode)


def curve_from_data(data: dict, read_text) -> Curve:
    """CurveData -> Curve (reference src/parsing/curves.rs:298-372). `read_text(path)` resolves files."""
    t = data["type"]
    if t == "Blackbody":
        return Blackbody(float(data["temperature"]), float(data["strength"]))
    if t == "Linear":
        fx, fy = _domain_funcs(data.get("domain_mapping"))
        return parse_linear(read_text(data["filename"]), data["interpolation_mode"], fx, fy)
    if t == "TabulatedCSV":
        fx, fy = _domain_funcs(data.get("domain_mapping"))
        return parse_tabulated_csv(read_text(data["filename"]), int(data["column"]), data["interpolation_mode"], fx, fy)
    if t == "Flat":
        return Linear(_f32([data["strength"]]), EXTENDED_VISIBLE_RANGE, "Linear")
    if t == "Cauchy":
        return Cauchy(float(data["a"]), float(data["b"]))
    if t == "SimpleSpike":
        return Exponential([(float(data["lambda"]), float(data["left_taper"]), float(data["right_taper"]), float(data["strength"]))])
    raise ValueError(f"unknown curve type {t}")


def cie_e(power: float) -> Curve:
    return Linear(_f32([power]), EXTENDED_VISIBLE_RANGE, "Linear")


def void() -> Curve:
    return cie_e(0.0)


def mauve(power: float) -> Curve:
    """reference src/curves.rs:44-51."""
    return Exponential([(650.0, 300.0, 300.0, power), (460.0, 200.0, 400.0, 0.75 * power)])


def y_bar_curve() -> Curve:
    """math::Curve::y_bar() used as the default importance-map luminance curve
    (reference src/parsing/environment.rs:128-131). Recalled, unpinned."""
    return Exponential([(568.0, 46.9, 40.5, 0.821), (530.9, 16.3, 31.1, 0.286)])


# ---- CIE 1931 colour matching (math::spectral::{x_bar,y_bar,z_bar}, angstrom input) --------


def _g(x, mu, s1, s2):
    s = (x - mu) / np.where(x < mu, s1, s2)
    return np.exp(-0.5 * s * s)


def cie_xyz_bar(lam_nm: np.ndarray) -> np.ndarray:
    """Wyman/Sloan/Shirley 2013 multi-lobe fit; the reference passes lambda*10 (angstrom) to
    x_bar/y_bar/z_bar (src/world/importance_map.rs:364,461). Returns (3, n) float32."""
    a = lam_nm.astype(np.float64) * 10.0
    x = 1.056 * _g(a, 5998.0, 379.0, 310.0) + 0.362 * _g(a, 4420.0, 160.0, 267.0) - 0.065 * _g(a, 5011.0, 204.0, 262.0)
    y = 0.821 * _g(a, 5688.0, 469.0, 405.0) + 0.286 * _g(a, 5309.0, 163.0, 311.0)
    z = 1.217 * _g(a, 4370.0, 118.0, 360.0) + 0.681 * _g(a, 4590.0, 260.0, 138.0)
    return np.stack([x, y, z]).astype(F32)


def lut_grid(lo: float, hi: float, n: int) -> np.ndarray:
    """The uniform LUT grid shared by host bake, oracle and device: n points, endpoints included."""
    return (lo + (hi - lo) * (np.arange(n, dtype=np.float64) / (n - 1))).astype(F32)
