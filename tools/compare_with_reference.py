#!/usr/bin/env python
"""Reference-pin hook: compares an image of this renderer with one rendered by the reference (gatling's Vulkan path).

  python tools/compare_with_reference.py REF [--config c1] [--spp N] [--test-image IMG] [--aov color] [--threshold K]

REF is what the reference produced for the same scene and settings:
  * a float image (.pfm, .npy; an .exr converted to .pfm) of the colour AOV -> per-channel mean ratio, RMSE, relative RMSE
    and the two-seed noise floor of OUR render at the same sample count (an RMSE at the noise floor means "same estimator");
    the north-star bar is RMSE < 1e-3 at 1024 spp;
  * an 8-bit PNG written by the reference's graphical test harness (src/hdGatling/testenv/*/ref*.png once the Git-LFS
    content is present) -> our float image goes through the harness's own conversion (main.cpp:458-487) and byte compare
    (main.cpp:332-376): the number of differing bytes against the test's errorPixelThreshold.
The test image is rendered on cuda:0 through the product (config c1 = /root/reference/cornell.usda at 512x512, spp 64,
max-bounces 4 -- BASELINE.json configs[0]) unless --test-image names a float image to use instead.

This module is also the restatement of the reference test harness used by tests/ (harness_bytes, diff_against_ref,
stb_write_png): the one reference golden image that can be recovered without its LFS content -- Render.Empty1x1/ref.png,
70 bytes, pinned by the sha256 in its LFS pointer -- is reproduced bit for bit from a render of the empty stage.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import struct
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# ---- the reference harness: float AOV -> 8-bit RGBA (src/hdGatling/main.cpp:136-143, 458-487) -----------------------------
def _powf(x, e):
    """std::pow(float, float) of the C library the harness links (numpy's own float32 power differs from glibc's powf by one ulp on a sixth
    of all inputs: found by comparing with main.cpp compiled here, tests/test_reference_glsl_cpu.py::test_test_harness_conversion)."""
    try:
        import ctypes
        libm = ctypes.CDLL("libm.so.6")
        libm.powf.restype, libm.powf.argtypes = ctypes.c_float, [ctypes.c_float, ctypes.c_float]
        e = float(np.float32(e))
        flat = np.asarray(x, np.float32).ravel()
        return np.array([libm.powf(float(v), e) for v in flat], np.float32).reshape(np.shape(x))
    except OSError:                                             # no glibc: correctly rounded double power, equal to powf on all but ~0.06 % of inputs
        return np.power(np.asarray(x, np.float64), np.float64(np.float32(e))).astype(np.float32)


def accurate_linear_to_srgb(x):
    """_AccurateLinearToSrgb (main.cpp:136-143), float32 arithmetic as in the harness."""
    x = np.asarray(x, np.float32)
    lo = x * np.float32(12.92)
    hi = _powf(np.abs(x), np.float32(1.0) / np.float32(2.4)) * np.float32(1.055) - np.float32(0.055)
    return np.where(x <= np.float32(0.0031308), lo, hi).astype(np.float32)


def harness_bytes(img, gamma_encode=True):
    """main.cpp:458-487: (h, w, c) float32 (c = 1, 3 or 4) or (h, w) int32 -> (h, w, 4) uint8.  `uint8_t(r * 255.0)` is a
    double product truncated towards zero; values outside [0, 256) are undefined behaviour in C++ and wrap here the way
    x86-64 compilers make them wrap (convert to a 32-bit integer, keep the low byte)."""
    a = np.asarray(img)
    if a.dtype == np.int32 or a.dtype == np.uint32:                      # id AOVs: the low three bytes
        r = a.reshape(a.shape[0], a.shape[1]).astype(np.int64)
        out = np.zeros(r.shape + (4,), np.uint8)
        for k in range(3):
            out[..., k] = (r >> (8 * k)) & 0xFF
        out[..., 3] = 255
        return out
    a = a.astype(np.float32)
    if a.ndim == 2:
        a = a[..., None]
    c = a.shape[2]
    out = np.zeros(a.shape[:2] + (4,), np.uint8)
    for j in range(4):
        r = a[..., min(j, c - 1)]
        if gamma_encode:
            r = accurate_linear_to_srgb(r)
        v = r.astype(np.float64) * 255.0
        v = np.where(np.isfinite(v), v, 0.0)
        out[..., j] = (np.trunc(np.clip(v, -2147483648.0, 2147483647.0)).astype(np.int64) & 0xFF).astype(np.uint8)
    out[..., 3] = 255                                                    # "always opaque"
    return out


def diff_against_ref(test_bytes, ref_bytes, threshold=0):
    """main.cpp:332-376: counts differing BYTES (the harness calls them error pixels), builds the diff image 255 - |diff|.
    Returns (error_count, passed, diff_image)."""
    t = np.asarray(test_bytes, np.uint8).astype(np.int32)
    r = np.asarray(ref_bytes, np.uint8).astype(np.int32)
    if t.shape != r.shape:
        raise ValueError("image sizes differ: %s vs %s" % (t.shape, r.shape))
    d = np.abs(r - t)
    errors = int((d > 0).sum())
    return errors, errors <= threshold, (255 - d).astype(np.uint8)


# ---- stb_image_write's PNG encoder (Hio's stb plugin writes the harness images), restated ------------------------------------
def _stb_zlib_compress(data: bytes, quality=8) -> bytes:
    """stbi_zlib_compress: one fixed-Huffman block, hash-chained LZ77 with one step of lazy matching, stored blocks when that
    is smaller.  Verified against the reference only on the literal-only path (the 70-byte Render.Empty1x1/ref.png)."""
    lengthc = [3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258, 259]
    lengtheb = [0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0]
    distc = [1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577, 32768]
    disteb = [0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13]
    quality = max(quality, 5)
    out = bytearray([0x78, 0x5E])
    st = [0, 0]

    def add(code, bits):
        st[0] |= code << st[1]; st[1] += bits
        while st[1] >= 8:
            out.append(st[0] & 0xFF); st[0] >>= 8; st[1] -= 8

    def rev(code, bits):
        r = 0
        for _ in range(bits):
            r = (r << 1) | (code & 1); code >>= 1
        return r

    def huff(n):
        if n <= 143: add(rev(0x30 + n, 8), 8)
        elif n <= 255: add(rev(0x190 + n - 144, 9), 9)
        elif n <= 279: add(rev(n - 256, 7), 7)
        else: add(rev(0xC0 + n - 280, 8), 8)

    def zhash(i):
        h = (data[i] + (data[i + 1] << 8) + (data[i + 2] << 16)) & 0xFFFFFFFF
        h ^= (h << 3) & 0xFFFFFFFF; h = (h + (h >> 5)) & 0xFFFFFFFF
        h ^= (h << 4) & 0xFFFFFFFF; h = (h + (h >> 17)) & 0xFFFFFFFF
        h ^= (h << 25) & 0xFFFFFFFF; h = (h + (h >> 6)) & 0xFFFFFFFF
        return h & 16383

    def countm(a, b, limit):
        n = 0
        while n < limit and n < 258 and data[a + n] == data[b + n]:
            n += 1
        return n

    add(1, 1); add(1, 2)
    table = {}
    n = len(data)
    i = 0
    while i < n - 3:
        h = zhash(i); best = 3; bestloc = None
        hl = table.setdefault(h, [])
        for p in hl:
            if p > i - 32768:
                d = countm(p, i, n - i)
                if d >= best:
                    best = d; bestloc = p
        if len(hl) == 2 * quality:
            del hl[:quality]
        hl.append(i)
        if bestloc is not None:
            for p in table.get(zhash(i + 1), []):
                if p > i - 32767 and countm(p, i + 1, n - i - 1) > best:
                    bestloc = None
                    break
        if bestloc is not None:
            d = i - bestloc
            j = 0
            while best > lengthc[j + 1] - 1: j += 1
            huff(j + 257)
            if lengtheb[j]: add(best - lengthc[j], lengtheb[j])
            j = 0
            while d > distc[j + 1] - 1: j += 1
            add(rev(j, 5), 5)
            if disteb[j]: add(d - distc[j], disteb[j])
            i += best
        else:
            huff(data[i]); i += 1
    while i < n:
        huff(data[i]); i += 1
    huff(256)
    while st[1]:
        add(0, 1)
    if len(out) > n + 2 + ((n + 32766) // 32767) * 5:
        del out[2:]
        j = 0
        while j < n:
            bl = min(n - j, 32767)
            out.append(1 if n - j == bl else 0)
            out += struct.pack("<HH", bl, (~bl) & 0xFFFF)
            out += data[j:j + bl]
            j += bl
    out += struct.pack(">I", zlib.adler32(data) & 0xFFFFFFFF)
    return bytes(out)


def stb_write_png(rgba8) -> bytes:
    """stbi_write_png of an (h, w, 4) uint8 image in MEMORY order (the harness passes `flipped = true`, i.e. its first
    row is the bottom of the picture; Hio flips before it encodes)."""
    a = np.ascontiguousarray(rgba8, np.uint8)
    h, w, n = a.shape
    a = a[::-1]                                                          # StorageSpec.flipped
    lines = bytearray()
    firstmap = [0, 1, 0, 5, 6]
    for y in range(h):
        row = a[y].reshape(-1).astype(np.int32)
        prev = a[y - 1].reshape(-1).astype(np.int32) if y else np.zeros_like(row)
        best = None
        for ft in range(5):
            t = ft if y else firstmap[ft]
            left = np.concatenate([np.zeros(n, np.int32), row[:-n]])
            ul = np.concatenate([np.zeros(n, np.int32), prev[:-n]])
            if t == 0: line = row.copy()
            elif t == 1: line = row - left
            elif t == 2: line = row - prev
            elif t == 3: line = row - ((left + prev) >> 1)
            elif t == 4:
                p = left + prev - ul
                pa, pb, pc = np.abs(p - left), np.abs(p - prev), np.abs(p - ul)
                pred = np.where((pa <= pb) & (pa <= pc), left, np.where(pb <= pc, prev, ul))
                line = row - pred
            elif t == 5: line = row - (left >> 1)
            else: line = row - left                                      # 6: paeth on the first row = sub
            line = (line & 0xFF).astype(np.uint8)
            est = int(np.abs(line.view(np.int8).astype(np.int32)).sum())
            if best is None or est < best[0]:
                best = (est, ft, line)
        lines.append(best[1]); lines += best[2].tobytes()

    def chunk(tag, payload):
        return struct.pack(">I", len(payload)) + tag + payload + struct.pack(">I", zlib.crc32(tag + payload) & 0xFFFFFFFF)
    ctype = {1: 0, 2: 4, 3: 2, 4: 6}[n]
    return (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, ctype, 0, 0, 0)) +
            chunk(b"IDAT", _stb_zlib_compress(bytes(lines))) + chunk(b"IEND", b""))


def lfs_pointer(path):
    """(sha256 hex, size) of a Git-LFS pointer file, or None when `path` is real content."""
    with open(path, "rb") as f:
        head = f.read(200)
    if not head.startswith(b"version https://git-lfs"):
        return None
    oid = size = None
    for ln in head.decode("ascii", "replace").splitlines():
        if ln.startswith("oid sha256:"): oid = ln.split(":", 1)[1].strip()
        if ln.startswith("size "): size = int(ln.split()[1])
    return oid, size


# ---- float images ----------------------------------------------------------------------------------------------------
def read_pfm(path):
    with open(path, "rb") as f:
        magic = f.readline().strip()
        dims = f.readline().split()
        while len(dims) < 2:
            dims += f.readline().split()
        scale = float(f.readline().strip())
        w, h = int(dims[0]), int(dims[1])
        ch = 3 if magic == b"PF" else 1
        a = np.frombuffer(f.read(w * h * ch * 4), "<f4" if scale < 0 else ">f4").reshape(h, w, ch)
    return a[::-1].astype(np.float32)                                    # PFM rows run bottom to top


def write_pfm(path, img):
    a = np.asarray(img, np.float32)[..., :3][::-1]
    with open(path, "wb") as f:
        f.write(b"PF\n%d %d\n-1.0\n" % (a.shape[1], a.shape[0]))
        f.write(np.ascontiguousarray(a, "<f4").tobytes())


def read_float_image(path):
    if path.endswith(".npy"):
        return np.load(path).astype(np.float32)
    if path.endswith(".pfm"):
        return read_pfm(path)
    raise SystemExit("float reference images: .pfm or .npy (convert an .exr with `oiiotool ref.exr -o ref.pfm`)")


def float_report(test, ref, noise=None):
    t = np.asarray(test, np.float64)[..., :3]; r = np.asarray(ref, np.float64)[..., :3]
    if t.shape != r.shape:
        raise SystemExit("image sizes differ: %s vs %s" % (t.shape, r.shape))
    rmse = float(np.sqrt(np.mean((t - r) ** 2)))
    rep = {"rmse": rmse, "relative_rmse": float(rmse / max(np.sqrt(np.mean(r ** 2)), 1e-30)),
           "mean_test": t.mean(axis=(0, 1)).tolist(), "mean_ref": r.mean(axis=(0, 1)).tolist(),
           "mean_ratio_per_channel": (t.mean(axis=(0, 1)) / np.maximum(r.mean(axis=(0, 1)), 1e-30)).tolist(),
           "max_abs": float(np.abs(t - r).max()), "north_star_bar_rmse": 1e-3, "below_bar": bool(rmse < 1e-3)}
    if noise is not None:
        rep["noise_floor_rmse"] = noise
        # two independent estimates of the same image differ by sqrt(2) x the noise of one
        rep["rmse_over_expected_noise"] = float(rmse / max(noise * np.sqrt(2.0), 1e-30))
    return rep


def render_config(config, spp):
    """Renders through the product on cuda:0 (no CPU fallback: fails loudly without the CUDA library or a GPU).  Returns two
    images of `spp` samples each over disjoint sample ranges: the second giRender call accumulates progressively, as the
    reference's does (rgen:506-515), so samples [spp, 2 spp) alone are 2 x (second result) - (first result)."""
    from gatling_b200 import gi, scenes
    if config != "c1":
        raise SystemExit("configs other than c1 need the production asset; pass --test-image")
    st = gi.giInitialize(0)
    if st != gi.GiStatus.Ok:
        raise SystemExit("giInitialize failed: " + gi.giLastError())
    d = scenes.cornell(512, 512, spp=spp, max_bounces=4)
    b = gi.BuiltScene(d)
    try:
        first = b.render().copy()
        both = b.render().copy()
    finally:
        b.close()
        gi.giTerminate()
    second = 2.0 * both - first
    second[..., 3] = first[..., 3]
    return [first, second.astype(np.float32)]


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("ref")
    ap.add_argument("--config", default="c1")
    ap.add_argument("--spp", type=int, default=64)
    ap.add_argument("--test-image", default=None, help="float image (.npy/.pfm) to compare instead of rendering")
    ap.add_argument("--second-seed-image", default=None, help="a second render of ours with other samples (noise floor) when --test-image is used")
    ap.add_argument("--aov", default="color", help="colour AOVs are gamma-encoded by the harness, the others are not")
    ap.add_argument("--threshold", type=int, default=0, help="the test's gtl:errorPixelThreshold")
    a = ap.parse_args()

    if a.test_image:
        imgs = [read_float_image(a.test_image)] + ([read_float_image(a.second_seed_image)] if a.second_seed_image else [])
    else:
        imgs = render_config(a.config, a.spp)
    test = imgs[0]
    noise = None
    if len(imgs) > 1:
        noise = float(np.sqrt(np.mean((imgs[0][..., :3].astype(np.float64) - imgs[1][..., :3]) ** 2)) / np.sqrt(2.0))

    if a.ref.endswith(".png"):
        ptr = lfs_pointer(a.ref)
        ours = harness_bytes(test, gamma_encode=(a.aov == "color"))
        if ptr is not None:
            png = stb_write_png(ours)
            out = {"mode": "lfs-pointer", "ref_sha256": ptr[0], "ref_size": ptr[1], "our_png_size": len(png),
                   "our_png_sha256": hashlib.sha256(png).hexdigest(), "identical": hashlib.sha256(png).hexdigest() == ptr[0],
                   "note": "the reference image is a Git-LFS pointer: only a byte-identical 8-bit image re-encoded by the same PNG writer can be confirmed"}
        else:
            from PIL import Image
            ref = np.asarray(Image.open(a.ref).convert("RGBA"))[::-1]     # PNG rows top-down -> the harness's bottom-up memory order
            errors, ok, _ = diff_against_ref(ours, ref, a.threshold)
            out = {"mode": "harness-8bit", "error_bytes": errors, "threshold": a.threshold, "passed": ok, "bytes": int(ref.size)}
    else:
        out = {"mode": "float", **float_report(test, read_float_image(a.ref), noise)}
    print(json.dumps(out, indent=1))
    return 0


if __name__ == "__main__":
    sys.exit(main())
