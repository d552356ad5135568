#!/usr/bin/env python3
"""XYZ film (.npy, HxWx4) -> sRGB PNG for eyeballing (not part of the product path)."""
import sys
import numpy as np
from PIL import Image

M = np.array([[3.2404542, -1.5371385, -0.4985314], [-0.9692660, 1.8760108, 0.0415560], [0.0556434, -0.2040259, 1.0572252]])

def to_png(film, path, exposure=1.0):
    rgb = np.clip(film[..., :3] @ M.T * exposure, 0, None)
    rgb = rgb / (1.0 + rgb)
    rgb = np.where(rgb <= 0.0031308, 12.92 * rgb, 1.055 * np.power(rgb, 1 / 2.4) - 0.055)
    Image.fromarray((np.clip(rgb, 0, 1) * 255).astype(np.uint8)).save(path)

if __name__ == "__main__":
    to_png(np.load(sys.argv[1]), sys.argv[2], float(sys.argv[3]) if len(sys.argv) > 3 else 1.0)
