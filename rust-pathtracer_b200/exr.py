"""Minimal OpenEXR writer / reader for the linear-RGB payload of output_film (reference src/tonemap/mod.rs:225-247 writes
`<filename>.exr` through the `exr` crate's write_rgb_file: three f32 channels). File encoding is host work on both sides of
the boundary ("EXR/PNG encoding stays on host"); the device produces the payload (rpt_output_film).

Written here: single-part scanline file, version 2, NO_COMPRESSION, channels B, G, R (FLOAT), one chunk per scanline."""
from __future__ import annotations

import struct

import numpy as np

MAGIC = 20000630


def _attr(name: str, typ: str, value: bytes) -> bytes:
    return name.encode() + b"\0" + typ.encode() + b"\0" + struct.pack("<i", len(value)) + value


def write_exr_rgb(path: str, rgb: np.ndarray) -> None:
    """rgb: (H, W, 3) float32 linear RGB."""
    rgb = np.ascontiguousarray(rgb, dtype=np.float32)
    h, w, c = rgb.shape
    assert c == 3
    chlist = b"".join(n.encode() + b"\0" + struct.pack("<iB3xii", 2, 0, 1, 1) for n in ("B", "G", "R")) + b"\0"
    box = struct.pack("<4i", 0, 0, w - 1, h - 1)
    header = struct.pack("<ii", MAGIC, 2)
    header += _attr("channels", "chlist", chlist)
    header += _attr("compression", "compression", b"\0")
    header += _attr("dataWindow", "box2i", box)
    header += _attr("displayWindow", "box2i", box)
    header += _attr("lineOrder", "lineOrder", b"\0")
    header += _attr("pixelAspectRatio", "float", struct.pack("<f", 1.0))
    header += _attr("screenWindowCenter", "v2f", struct.pack("<2f", 0.0, 0.0))
    header += _attr("screenWindowWidth", "float", struct.pack("<f", 1.0))
    header += b"\0"
    line_bytes = 3 * w * 4
    chunk = 8 + line_bytes
    table_at = len(header)
    first = table_at + 8 * h
    offsets = struct.pack(f"<{h}Q", *[first + y * chunk for y in range(h)])
    bgr = rgb[:, :, ::-1].transpose(0, 2, 1)  # (H, channel B/G/R, W): channels are stored one after the other per scanline
    with open(path, "wb") as f:
        f.write(header)
        f.write(offsets)
        for y in range(h):
            f.write(struct.pack("<ii", y, line_bytes))
            f.write(np.ascontiguousarray(bgr[y]).tobytes())


def read_exr_rgb(path: str) -> np.ndarray:
    """Reads back what write_exr_rgb writes (uncompressed scanline, FLOAT channels); -> (H, W, 3) float32 RGB."""
    data = open(path, "rb").read()
    magic, version = struct.unpack_from("<ii", data, 0)
    if magic != MAGIC or (version & 0xFF) != 2 or (version & ~0xFF):
        raise ValueError("not a single-part scanline OpenEXR 2 file")
    pos, attrs = 8, {}
    while data[pos] != 0:
        e = data.index(b"\0", pos)
        name = data[pos:e].decode()
        e2 = data.index(b"\0", e + 1)
        typ = data[e + 1:e2].decode()
        (size,) = struct.unpack_from("<i", data, e2 + 1)
        attrs[name] = (typ, data[e2 + 5:e2 + 5 + size])
        pos = e2 + 5 + size
    pos += 1
    if attrs["compression"][1] != b"\0":
        raise ValueError("compressed EXR files are not supported by this reader")
    names, p, ch = [], 0, attrs["channels"][1]
    while ch[p] != 0:
        e = ch.index(b"\0", p)
        names.append(ch[p:e].decode())
        (ptype,) = struct.unpack_from("<i", ch, e + 1)
        if ptype != 2:
            raise ValueError("only FLOAT channels are supported")
        p = e + 1 + 16
    x0, y0, x1, y1 = struct.unpack("<4i", attrs["dataWindow"][1])
    w, h = x1 - x0 + 1, y1 - y0 + 1
    offsets = struct.unpack_from(f"<{h}Q", data, pos)
    out = np.zeros((h, w, len(names)), dtype=np.float32)
    for off in offsets:
        y, size = struct.unpack_from("<ii", data, off)
        line = np.frombuffer(data, dtype="<f4", count=len(names) * w, offset=off + 8).reshape(len(names), w)
        out[y - y0] = line.T
    order = [names.index(n) for n in ("R", "G", "B")]
    return out[:, :, order]
