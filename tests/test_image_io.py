"""The library's own image decoders (csrc/image_io.cpp, `vr_image_load_rgb32f`) against PIL / OpenCV on the
repository's texture files and on files written here in every supported encoding. They stand in for
`image::open(path).to_rgb32f()` (core/texture.rs:37, environments.rs:43). Lossless formats must match bit for
bit; JPEG is compared within the usual inter-decoder IDCT / colour rounding difference. Host-only: no GPU."""
import os

os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")

import numpy as np
import pytest

from voidray_b200 import _lib, assets

ASSETS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "assets")
F32 = np.float32


def _rng_image(h, w, c, dtype, seed=0):
    rng = np.random.default_rng(seed)
    hi = np.iinfo(dtype).max
    # smooth + noise so every PNG filter / LZW path gets exercised
    y, x = np.mgrid[0:h, 0:w]
    base = ((np.sin(x / 7.0)[..., None] + np.cos(y / 5.0)[..., None] + 2.0) / 4.0 * hi)
    img = base + rng.integers(0, hi // 8 + 1, size=(h, w, c))
    return np.clip(img, 0, hi).astype(dtype)


@pytest.mark.parametrize("name", ["test.png", "uv_test.png", "wood_albedo.tif", "wood_normal.tif"])
def test_lossless_assets_match_pil(name):
    path = os.path.join(ASSETS, name)
    assert np.array_equal(assets.load_image_native(path), assets.load_image_rgb32f(path))


@pytest.mark.parametrize("name", ["mossy_ground_albedo.jpg", "mushroom_albedo.jpg"])
def test_jpeg_assets_match_libjpeg_within_rounding(name):
    path = os.path.join(ASSETS, name)
    a = assets.load_image_native(path)
    b = assets.load_image_rgb32f(path)
    assert a.shape == b.shape == (2048, 2048, 3)
    d = np.abs(np.rint(a * 255.0) - np.rint(b * 255.0))
    assert d.max() <= 3 and d.mean() < 0.01 and (d > 0).mean() < 0.01


def test_png_variants(tmp_path):
    from PIL import Image

    rgb = _rng_image(37, 53, 3, np.uint8)
    cases = {}
    Image.fromarray(rgb).save(tmp_path / "rgb.png")
    cases["rgb.png"] = rgb
    rgba = np.dstack([rgb, _rng_image(37, 53, 1, np.uint8, 1)])
    Image.fromarray(rgba).save(tmp_path / "rgba.png")
    cases["rgba.png"] = rgb
    grey = rgb[:, :, 0]
    Image.fromarray(grey).save(tmp_path / "grey.png")
    cases["grey.png"] = np.repeat(grey[:, :, None], 3, 2)
    Image.fromarray(np.dstack([grey, rgba[:, :, 3]]), "LA").save(tmp_path / "la.png")
    cases["la.png"] = np.repeat(grey[:, :, None], 3, 2)
    pal = Image.fromarray(rgb).quantize(17)
    pal.save(tmp_path / "pal.png", bits=8)
    cases["pal.png"] = np.asarray(pal.convert("RGB"))
    pal4 = Image.fromarray(rgb).quantize(13)
    pal4.save(tmp_path / "pal4.png", bits=4)
    cases["pal4.png"] = np.asarray(pal4.convert("RGB"))
    one = Image.fromarray((grey > 128).astype(np.uint8) * 255).convert("1")
    one.save(tmp_path / "bit.png")
    cases["bit.png"] = np.repeat((np.asarray(one).astype(np.uint8) * 255)[:, :, None], 3, 2)
    for name, want in cases.items():
        got = assets.load_image_native(str(tmp_path / name))
        assert np.array_equal(got, want.astype(F32) / F32(255.0)), name
    # 16-bit grey
    g16 = _rng_image(19, 23, 1, np.uint16)[:, :, 0]
    Image.fromarray(g16).save(tmp_path / "g16.png")
    got = assets.load_image_native(str(tmp_path / "g16.png"))
    assert np.array_equal(got, np.repeat((g16.astype(F32) / F32(65535.0))[:, :, None], 3, 2))


def test_png_16bit_rgb_and_interlaced(tmp_path):
    import cv2

    rgb16 = _rng_image(33, 41, 3, np.uint16)
    assert cv2.imwrite(str(tmp_path / "rgb16.png"), rgb16[:, :, ::-1])
    got = assets.load_image_native(str(tmp_path / "rgb16.png"))
    assert np.array_equal(got, rgb16.astype(F32) / F32(65535.0))
    # Adam7: hand-assembled with zlib (PIL / OpenCV cannot write interlaced files)
    import struct
    import zlib

    rgb = _rng_image(21, 30, 3, np.uint8)
    h, w, _ = rgb.shape
    x0, y0, dx, dy = [0, 4, 0, 2, 0, 1, 0], [0, 0, 4, 0, 2, 0, 1], [8, 8, 4, 4, 2, 2, 1], [8, 8, 8, 4, 4, 2, 2]
    raw = b""
    for p in range(7):
        sub = rgb[y0[p]::dy[p], x0[p]::dx[p]]
        if sub.size == 0:
            continue
        for row in sub:
            raw += b"\x00" + row.tobytes()

    def chunk(t, body):
        return struct.pack(">I", len(body)) + t + body + struct.pack(">I", zlib.crc32(t + body))

    png = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 1)) + \
        chunk(b"IDAT", zlib.compress(raw)) + chunk(b"IEND", b"")
    (tmp_path / "adam7.png").write_bytes(png)
    from PIL import Image
    assert np.array_equal(np.asarray(Image.open(tmp_path / "adam7.png").convert("RGB")), rgb)  # the file is valid
    assert np.array_equal(assets.load_image_native(str(tmp_path / "adam7.png")), rgb.astype(F32) / F32(255.0))


@pytest.mark.parametrize("compression", ["raw", "tiff_lzw", "tiff_adobe_deflate", "packbits"])
def test_tiff_compressions(tmp_path, compression):
    from PIL import Image

    rgb = _rng_image(67, 91, 3, np.uint8)
    path = str(tmp_path / f"{compression}.tif")
    Image.fromarray(rgb).save(path, compression=compression)
    assert np.array_equal(assets.load_image_native(path), rgb.astype(F32) / F32(255.0))
    grey = rgb[:, :, 1]
    Image.fromarray(grey).save(path, compression=compression)
    assert np.array_equal(assets.load_image_native(path), np.repeat((grey.astype(F32) / F32(255.0))[:, :, None], 3, 2))


def test_tiff_large_lzw_predictor_and_16bit(tmp_path):
    import cv2
    from PIL import Image

    rgb = _rng_image(300, 400, 3, np.uint8)  # large enough for the LZW table to fill and reset many times
    path = str(tmp_path / "big.tif")
    Image.fromarray(rgb).save(path, compression="tiff_lzw", tiffinfo={317: 2})
    im = Image.open(path)
    assert im.tag_v2.get(317) == 2
    assert np.array_equal(assets.load_image_native(path), rgb.astype(F32) / F32(255.0))
    rgb16 = _rng_image(64, 80, 3, np.uint16)
    p16 = str(tmp_path / "rgb16.tif")
    assert cv2.imwrite(p16, rgb16[:, :, ::-1])  # OpenCV writes LZW + predictor for 16-bit TIFF
    assert np.array_equal(assets.load_image_native(p16), rgb16.astype(F32) / F32(65535.0))


@pytest.mark.parametrize("kwargs", [
    dict(subsampling=0), dict(subsampling=1), dict(subsampling=2), dict(subsampling=0, progressive=True),
    dict(subsampling=2, progressive=True), dict(subsampling=2, restart_marker_blocks=3), dict(grey=True),
    dict(grey=True, progressive=True), dict(subsampling=0, quality=100), dict(subsampling=2, optimize=True),
])
def test_jpeg_variants(tmp_path, kwargs):
    from PIL import Image

    kwargs = dict(kwargs)
    grey = kwargs.pop("grey", False)
    # smooth content, like photographs: the decoders differ only in rounding
    y, x = np.mgrid[0:75, 0:117]
    img = np.stack([127 + 100 * np.sin(x / 9.0) * np.cos(y / 11.0), 127 + 90 * np.cos(x / 13.0 + y / 7.0),
                    127 + 110 * np.sin((x + y) / 17.0)], axis=2).astype(np.uint8)
    im = Image.fromarray(img[:, :, 0] if grey else img)
    path = str(tmp_path / "t.jpg")
    kwargs.setdefault("quality", 90)
    im.save(path, **kwargs)
    got = np.rint(assets.load_image_native(path) * 255.0)
    want = np.asarray(Image.open(path).convert("RGB")).astype(np.float64)
    d = np.abs(got - want)
    assert got.shape == want.shape
    # chroma upsampling filters differ slightly between libjpeg-turbo builds; rounding is at most a few levels
    assert d.max() <= 4 and d.mean() < 0.5, (d.max(), d.mean())


def test_radiance_hdr(tmp_path):
    import cv2

    rng = np.random.default_rng(3)
    img = (rng.random((40, 64, 3)) ** 4 * 50.0).astype(F32)
    img[5:9, 10:30] = 0.0
    img[20:25] = F32(0.75)  # runs
    path = str(tmp_path / "t.hdr")
    assert cv2.imwrite(path, img[:, :, ::-1])
    want = cv2.imread(path, cv2.IMREAD_UNCHANGED)[:, :, ::-1]
    assert np.array_equal(assets.load_image_native(path), want)


def _exr_flags():
    import cv2

    return cv2, {
        "none": 0, "rle": 1, "zips": 2, "zip": 3, "piz": 4, "pxr24": 5, "b44": 6, "b44a": 7,
    }


@pytest.mark.parametrize("compression", ["none", "rle", "zips", "zip", "piz", "pxr24", "b44", "b44a"])
@pytest.mark.parametrize("half", [False, True])
def test_openexr(tmp_path, compression, half):
    cv2, flags = _exr_flags()
    if not hasattr(cv2, "IMWRITE_EXR_COMPRESSION"):
        pytest.skip("this OpenCV cannot choose the EXR compression")
    img = assets.synth_hdri("studio", 98, 50)  # not multiples of 4: partial B44 blocks
    rng = np.random.default_rng(5)
    img = (img * (1.0 + 0.05 * rng.random(img.shape))).astype(F32)
    path = str(tmp_path / f"{compression}_{half}.exr")
    params = [cv2.IMWRITE_EXR_COMPRESSION, flags[compression], cv2.IMWRITE_EXR_TYPE,
              cv2.IMWRITE_EXR_TYPE_HALF if half else cv2.IMWRITE_EXR_TYPE_FLOAT]
    try:
        ok = cv2.imwrite(path, img[:, :, ::-1], params)
    except cv2.error as e:  # OpenCV built without the codec
        pytest.skip(str(e))
    assert ok
    want = cv2.imread(path, cv2.IMREAD_UNCHANGED)[:, :, ::-1].astype(F32)
    got = assets.load_image_native(path)
    assert np.array_equal(got, want)
    if not half and compression != "pxr24":  # PXR24 keeps 24 bits of a float; B44 only packs half channels
        assert np.array_equal(got, img)


def test_bmp_tga_pnm_farbfeld(tmp_path):
    import struct

    from PIL import Image

    rgb = _rng_image(29, 45, 3, np.uint8)  # odd width: BMP row padding, TGA / PBM bit packing
    want = rgb.astype(F32) / F32(255.0)
    grey = rgb[:, :, 1]
    wantg = np.repeat((grey.astype(F32) / F32(255.0))[:, :, None], 3, 2)
    im = Image.fromarray(rgb)
    cases = []
    im.save(tmp_path / "rgb24.bmp")
    cases.append(("rgb24.bmp", want))
    Image.fromarray(np.dstack([rgb, grey])).save(tmp_path / "rgba32.bmp")
    cases.append(("rgba32.bmp", want))
    pal = im.quantize(200)
    pal.save(tmp_path / "pal8.bmp")
    cases.append(("pal8.bmp", np.asarray(pal.convert("RGB")).astype(F32) / F32(255.0)))
    pal16 = im.quantize(16)
    pal16.save(tmp_path / "pal4.bmp", bits=4)
    cases.append(("pal4.bmp", np.asarray(Image.open(tmp_path / "pal4.bmp").convert("RGB")).astype(F32) / F32(255.0)))
    Image.fromarray(grey).save(tmp_path / "grey8.bmp")
    cases.append(("grey8.bmp", wantg))
    bit = Image.fromarray((grey > 128).astype(np.uint8) * 255).convert("1")
    bit.save(tmp_path / "bit.bmp")
    wantb = np.repeat((np.asarray(bit).astype(F32))[:, :, None], 3, 2)
    cases.append(("bit.bmp", wantb))
    for rle in (False, True):
        tag = "rle" if rle else "raw"
        im.save(tmp_path / f"rgb_{tag}.tga", compression="tga_rle" if rle else None)
        cases.append((f"rgb_{tag}.tga", want))
        Image.fromarray(np.dstack([rgb, grey])).save(tmp_path / f"rgba_{tag}.tga", compression="tga_rle" if rle else None)
        cases.append((f"rgba_{tag}.tga", want))
        Image.fromarray(grey).save(tmp_path / f"grey_{tag}.tga", compression="tga_rle" if rle else None)
        cases.append((f"grey_{tag}.tga", wantg))
        pal.save(tmp_path / f"pal_{tag}.tga", compression="tga_rle" if rle else None)
        cases.append((f"pal_{tag}.tga", np.asarray(pal.convert("RGB")).astype(F32) / F32(255.0)))
    im.save(tmp_path / "flip.tga", orientation=1)  # top-down rows
    cases.append(("flip.tga", want))
    im.save(tmp_path / "p6.ppm")
    cases.append(("p6.ppm", want))
    Image.fromarray(grey).save(tmp_path / "p5.pgm")
    cases.append(("p5.pgm", wantg))
    bit.save(tmp_path / "p4.pbm")
    cases.append(("p4.pbm", wantb))
    (tmp_path / "p3.ppm").write_text("P3\n# comment\n%d %d\n255\n" % (rgb.shape[1], rgb.shape[0]) +
                                     "\n".join(" ".join(str(v) for v in row.reshape(-1)) for row in rgb) + "\n")
    cases.append(("p3.ppm", want))
    (tmp_path / "p2.pgm").write_text("P2 %d %d 255\n" % (grey.shape[1], grey.shape[0]) + " ".join(str(v) for v in grey.reshape(-1)))
    cases.append(("p2.pgm", wantg))
    (tmp_path / "p1.pbm").write_text("P1\n%d %d\n" % (grey.shape[1], grey.shape[0]) +
                                     "\n".join("".join("1" if v <= 128 else "0" for v in row) for row in grey))
    cases.append(("p1.pbm", wantb))
    g16 = _rng_image(11, 13, 3, np.uint16)
    (tmp_path / "p6_16.ppm").write_bytes(b"P6\n13 11\n65535\n" + g16.astype(">u2").tobytes())
    cases.append(("p6_16.ppm", g16.astype(F32) / F32(65535.0)))
    rgba16 = np.dstack([g16, g16[:, :, 0]])
    (tmp_path / "t.ff").write_bytes(b"farbfeld" + struct.pack(">II", 13, 11) + rgba16.astype(">u2").tobytes())
    cases.append(("t.ff", g16.astype(F32) / F32(65535.0)))
    for name, w in cases:
        got = assets.load_image_native(str(tmp_path / name))
        assert got.shape == w.shape, name
        assert np.array_equal(got, w), name
    # 16-bit 5-5-5 BMP: channel expansion rounds (v * 255 / 31), PIL truncates: at most one level apart
    v = (rgb >> 3).astype(np.uint16)
    px = (v[:, :, 0] << 10) | (v[:, :, 1] << 5) | v[:, :, 2]
    rows = b"".join(row.astype("<u2").tobytes() + b"\0" * ((4 - (2 * row.size) % 4) % 4) for row in px[::-1])
    hdr = struct.pack("<2sIHHI", b"BM", 54 + len(rows), 0, 0, 54) + struct.pack("<IiiHHIIiiII", 40, px.shape[1], px.shape[0], 1, 16, 0,
                                                                                len(rows), 2835, 2835, 0, 0)
    (tmp_path / "rgb555.bmp").write_bytes(hdr + rows)
    got = np.rint(assets.load_image_native(str(tmp_path / "rgb555.bmp")) * 255.0)
    assert np.array_equal(got, np.rint(v.astype(np.float64) * 255.0 / 31.0))
    ref = np.asarray(Image.open(tmp_path / "rgb555.bmp").convert("RGB")).astype(np.float64)
    assert np.abs(got - ref).max() <= 1


def _write_tiled_tiff(path, img, tile, deflate, predictor, big_endian=False):
    """A minimal tiled TIFF writer (chunky RGB, 8 or 16 bit, none / Deflate, optional horizontal differencing); PIL
    (libtiff) reads the files it writes, which is checked before the library's decoder is."""
    import struct
    import zlib

    h, w, spp = img.shape
    tw, tl = tile
    bps = img.dtype.itemsize
    e = ">" if big_endian else "<"
    nx, ny = (w + tw - 1) // tw, (h + tl - 1) // tl
    blobs = []
    for ty in range(ny):
        for tx in range(nx):
            t = np.zeros((tl, tw, spp), img.dtype)
            part = img[ty * tl:(ty + 1) * tl, tx * tw:(tx + 1) * tw]
            t[:part.shape[0], :part.shape[1]] = part
            if predictor:
                d = t.copy()
                d[:, 1:] = t[:, 1:] - t[:, :-1]  # wraps modulo 2^bits
                t = d
            raw = t.astype(t.dtype.newbyteorder(e)).tobytes()
            blobs.append(zlib.compress(raw) if deflate else raw)
    entries = []  # (tag, type, count, values)
    def add(tag, typ, vals):
        entries.append((tag, typ, vals))
    add(256, 4, [w]); add(257, 4, [h]); add(258, 3, [8 * bps] * spp); add(259, 3, [8 if deflate else 1]); add(262, 3, [2])
    add(277, 3, [spp]); add(284, 3, [1]); add(317, 3, [2 if predictor else 1]); add(322, 4, [tw]); add(323, 4, [tl])
    add(324, 4, None); add(325, 4, [len(b) for b in blobs])
    entries.sort(key=lambda t: t[0])
    ifd_at = 8
    ifd_size = 2 + 12 * len(entries) + 4
    extra_at = ifd_at + ifd_size
    # first pass: sizes of out-of-line arrays
    def size_of(typ, n):
        return n * (2 if typ == 3 else 4)
    extra = b""
    placed = {}
    data_at = None
    for tag, typ, vals in entries:
        n = len(blobs) if vals is None else len(vals)
        if size_of(typ, n) > 4:
            placed[tag] = extra_at + len(extra)
            extra += b"\0" * size_of(typ, n)
    data_at = extra_at + len(extra)
    offs, at = [], data_at
    for b in blobs:
        offs.append(at)
        at += len(b)
    out = (b"MM" if big_endian else b"II") + struct.pack(e + "HI", 42, ifd_at) + struct.pack(e + "H", len(entries))
    extra = bytearray(extra)
    for tag, typ, vals in entries:
        if vals is None:
            vals = offs
        n = len(vals)
        body = b"".join(struct.pack(e + ("H" if typ == 3 else "I"), v) for v in vals)
        if len(body) > 4:
            o = placed[tag] - extra_at
            extra[o:o + len(body)] = body
            out += struct.pack(e + "HHII", tag, typ, n, placed[tag])
        else:
            out += struct.pack(e + "HHI", tag, typ, n) + body.ljust(4, b"\0")
    out += struct.pack(e + "I", 0) + bytes(extra) + b"".join(blobs)
    with open(path, "wb") as f:
        f.write(out)


@pytest.mark.parametrize("dtype,deflate,predictor,big", [(np.uint8, False, False, False), (np.uint8, True, True, False),
                                                         (np.uint16, True, True, False), (np.uint16, True, True, True),
                                                         (np.uint16, False, False, True)])
def test_tiff_tiled(tmp_path, dtype, deflate, predictor, big):
    from PIL import Image

    img = _rng_image(70, 101, 3, dtype)  # edge tiles are partial in both directions
    path = str(tmp_path / "tiled.tif")
    _write_tiled_tiff(path, img, (32, 16), deflate, predictor, big)
    scale = F32(255.0) if dtype == np.uint8 else F32(65535.0)
    if dtype == np.uint8:  # PIL has no 16-bit RGB mode; libtiff still validates the 8-bit files
        assert np.array_equal(np.asarray(Image.open(path).convert("RGB")), img)
    else:
        import cv2
        ref = cv2.imread(path, cv2.IMREAD_UNCHANGED)
        assert ref is not None and np.array_equal(ref[:, :, ::-1], img)
    assert np.array_equal(assets.load_image_native(path), img.astype(F32) / scale)


def _write_tiled_exr(path, img, tile, compression, half):
    """A minimal single-part tiled OpenEXR writer (ONE_LEVEL; compression 0 = none or 3 = ZIP), enough to exercise the
    decoder's tile path; the files it writes are checked with OpenCV's (OpenEXR's) reader first."""
    import struct
    import zlib

    h, w, _ = img.shape
    tw, th = tile
    dt = np.float16 if half else np.float32

    def attr(name, typ, body):
        return name.encode() + b"\0" + typ.encode() + b"\0" + struct.pack("<i", len(body)) + body

    chlist = b"".join(n.encode() + b"\0" + struct.pack("<iBBBBii", 1 if half else 2, 0, 0, 0, 0, 1, 1) for n in "BGR") + b"\0"
    box = struct.pack("<iiii", 0, 0, w - 1, h - 1)
    header = struct.pack("<II", 20000630, 2 | 0x200)
    header += attr("channels", "chlist", chlist) + attr("compression", "compression", bytes([compression]))
    header += attr("dataWindow", "box2i", box) + attr("displayWindow", "box2i", box) + attr("lineOrder", "lineOrder", b"\0")
    header += attr("pixelAspectRatio", "float", struct.pack("<f", 1.0)) + attr("screenWindowCenter", "v2f", struct.pack("<ff", 0, 0))
    header += attr("screenWindowWidth", "float", struct.pack("<f", 1.0)) + attr("tiles", "tiledesc", struct.pack("<IIB", tw, th, 0))
    header += b"\0"
    nx, ny = (w + tw - 1) // tw, (h + th - 1) // th
    chunks = []
    for ty in range(ny):
        for tx in range(nx):
            t = img[ty * th:(ty + 1) * th, tx * tw:(tx + 1) * tw]
            raw = b"".join(t[y, :, c].astype(dt).tobytes() for y in range(t.shape[0]) for c in (2, 1, 0))  # B, G, R per line
            data = raw
            if compression == 3:
                a = np.frombuffer(raw, np.uint8)
                r = np.concatenate([a[0::2], a[1::2]]).astype(np.int32)
                d = r.copy()
                d[1:] = (r[1:] - r[:-1] + 128 + 256) & 255
                z = zlib.compress(d.astype(np.uint8).tobytes())
                data = z if len(z) < len(raw) else raw
            chunks.append(struct.pack("<iiiii", tx, ty, 0, 0, len(data)) + data)
    table_at = len(header)
    offs, at = [], table_at + 8 * len(chunks)
    for c in chunks:
        offs.append(at)
        at += len(c)
    with open(path, "wb") as f:
        f.write(header + b"".join(struct.pack("<Q", o) for o in offs) + b"".join(chunks))


@pytest.mark.parametrize("compression", [0, 3])
@pytest.mark.parametrize("half", [False, True])
def test_openexr_tiled(tmp_path, compression, half):
    import cv2

    img = assets.synth_hdri("indoor", 83, 37).astype(F32)  # neither dimension a multiple of the tile size
    img = (img * (1.0 + 0.05 * np.random.default_rng(9).random(img.shape))).astype(F32)
    path = str(tmp_path / "tiled.exr")
    _write_tiled_exr(path, img, (32, 16), compression, half)
    ref = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    assert ref is not None, "OpenEXR rejects the test file"
    want = ref[:, :, 2::-1].astype(F32)
    assert np.array_equal(want, img.astype(np.float16).astype(F32) if half else img)
    assert np.array_equal(assets.load_image_native(path), want)


def _rle8_bmp(idx, palette, absolute_every=0):
    """A BMP with BI_RLE8 pixel data, written by hand (Pillow reads RLE BMPs but does not write them)."""
    import struct
    h, w = idx.shape
    body = bytearray()
    for y in range(h - 1, -1, -1):  # bottom-up
        row, x = idx[y], 0
        while x < w:
            if absolute_every and (x // 7) % absolute_every == 0 and w - x >= 3:
                n = min(5, w - x)  # absolute run of n >= 3 literal indices, padded to 16 bits
                body += bytes([0, n]) + bytes(int(v) for v in row[x:x + n]) + (b"\0" if n & 1 else b"")
                x += n
                continue
            n = 1
            while x + n < w and n < 255 and row[x + n] == row[x]:
                n += 1
            body += bytes([n, int(row[x])])
            x += n
        body += b"\0\0"
    body += b"\0\1"
    pal = b"".join(bytes([int(c[2]), int(c[1]), int(c[0]), 0]) for c in palette)
    off = 14 + 40 + len(pal)
    return (b"BM" + struct.pack("<IHHI", off + len(body), 0, 0, off) +
            struct.pack("<IiiHHIIiiII", 40, w, h, 1, 8, 1, len(body), 2835, 2835, len(palette), 0) + pal + bytes(body))


def test_gif_ico_dds_and_rle_bmp(tmp_path):
    # the rest of the `image` crate's default decoder set (image 0.24.3: gif, ico, dds / dxt; bmp with RLE)
    from PIL import Image

    y, x = np.mgrid[0:61, 0:83]
    rgb = np.stack([(np.sin(x / 7.0) + 1) * 120, (np.cos(y / 5.0) + 1) * 120, (x + y) % 256], -1).astype(np.uint8)
    pil = lambda p: np.asarray(Image.open(p).convert("RGB"), F32) / F32(255.0)  # noqa: E731
    # GIF: 8-bit and 3-bit code sizes, interlaced rows, a transparent index (keeps its palette colour), > 4096 codes
    Image.fromarray(rgb).quantize(200).save(tmp_path / "a.gif")
    Image.fromarray(rgb).quantize(200).save(tmp_path / "b.gif", interlace=True)
    Image.fromarray(rgb).quantize(200).save(tmp_path / "c.gif", transparency=3)
    Image.fromarray(rgb).quantize(5).save(tmp_path / "d.gif")
    Image.fromarray(np.repeat(np.repeat(rgb, 9, 0), 7, 1)).quantize(256).save(tmp_path / "e.gif")
    for n in "abcde":
        assert np.array_equal(assets.load_image_native(str(tmp_path / f"{n}.gif")), pil(tmp_path / f"{n}.gif")), n
    # ICO: PNG payload, BMP payload (XOR + AND bitmaps), several entries (the largest wins)
    icon = Image.fromarray(np.dstack([rgb[:48, :48], np.full((48, 48), 255, np.uint8)]))
    icon.save(tmp_path / "p.ico", sizes=[(48, 48)])
    icon.save(tmp_path / "q.ico", sizes=[(48, 48)], bitmap_format="bmp")
    icon.save(tmp_path / "r.ico", sizes=[(16, 16), (48, 48), (32, 32)], bitmap_format="bmp")
    for n in "pqr":
        got = assets.load_image_native(str(tmp_path / f"{n}.ico"))
        assert got.shape == (48, 48, 3) and np.array_equal(got, pil(tmp_path / f"{n}.ico")), n
    # DDS: DXT1 / DXT3 / DXT5 colour blocks; decoders differ in how they widen 5:6:5 and round the interpolants
    for fmt in ("DXT1", "DXT3", "DXT5"):
        Image.fromarray(np.dstack([rgb[:58, :79], np.full((58, 79), 255, np.uint8)])).save(tmp_path / f"{fmt}.dds", pixel_format=fmt)
        got = assets.load_image_native(str(tmp_path / f"{fmt}.dds"))
        d = np.abs(np.rint(got * 255.0) - np.rint(pil(tmp_path / f"{fmt}.dds") * 255.0))
        assert got.shape == (58, 79, 3) and d.max() <= 2, fmt
    # a hand-made DXT1 block pair: endpoints 0xFFFF / 0x0000 (4-colour mode) and the 3-colour + black mode
    import struct
    hdr = bytearray(128)
    hdr[0:4] = b"DDS "
    struct.pack_into("<III", hdr, 4, 124, 0x1007, 4)
    struct.pack_into("<I", hdr, 16, 8)
    struct.pack_into("<II4s", hdr, 76, 32, 4, b"DXT1")
    blocks = struct.pack("<HHI", 0xFFFF, 0x0000, 0xE4E4E4E4) + struct.pack("<HHI", 0x0000, 0xFFFF, 0xE4E4E4E4)
    (tmp_path / "hand.dds").write_bytes(bytes(hdr) + blocks)
    got = np.rint(assets.load_image_native(str(tmp_path / "hand.dds")) * 255.0)
    assert got[0, :4, 0].tolist() == [255, 0, 170, 85] and got[0, 4:, 0].tolist() == [0, 255, 128, 0]
    # BMP RLE8: encoded runs only, then with absolute runs mixed in
    pal_img = Image.fromarray(rgb).quantize(64)
    idx = np.asarray(pal_img)
    palette = np.asarray(pal_img.getpalette()[:64 * 3], np.uint8).reshape(-1, 3)
    for k, every in enumerate((0, 2)):
        path = tmp_path / f"rle{k}.bmp"
        path.write_bytes(_rle8_bmp(idx, palette, every))
        want = palette[idx].astype(F32) / F32(255.0)
        assert np.array_equal(pil(path), want)  # the file is a valid RLE8 BMP
        assert np.array_equal(assets.load_image_native(str(path)), want)
    for name, why in [("a.gif", "gif"), ("q.ico", "ico|bmp"), ("DXT5.dds", "dds")]:
        cut = tmp_path / ("cut_" + name)
        cut.write_bytes((tmp_path / name).read_bytes()[:150])
        with pytest.raises(_lib.VoidrayError, match=why):
            assets.load_image_native(str(cut))


def test_errors_are_reported_not_fatal(tmp_path):
    with pytest.raises(_lib.VoidrayError, match="cannot open"):
        assets.load_image_native(str(tmp_path / "missing.png"))
    with pytest.raises(_lib.VoidrayError, match="not a regular file"):
        assets.load_image_native(str(tmp_path))  # a directory
    bad = tmp_path / "bad.bin"
    bad.write_bytes(b"not an image at all")
    with pytest.raises(_lib.VoidrayError, match="unrecognised"):
        assets.load_image_native(str(bad))
    trunc = tmp_path / "trunc.png"
    trunc.write_bytes(open(os.path.join(ASSETS, "uv_test.png"), "rb").read()[:900])
    with pytest.raises(_lib.VoidrayError, match="png"):
        assets.load_image_native(str(trunc))
    tj = tmp_path / "trunc.jpg"
    tj.write_bytes(open(os.path.join(ASSETS, "mushroom_albedo.jpg"), "rb").read()[:100000])
    with pytest.raises(_lib.VoidrayError, match="jpeg"):
        assets.load_image_native(str(tj))
