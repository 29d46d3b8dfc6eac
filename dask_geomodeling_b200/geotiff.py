"""The on-disk formats either side of the raster path: tiled GeoTIFF and VRT mosaics.

The reference writes and reads its files through GDAL (`raster/sinks.py:78-124`,
`raster/sources.py:444-452`), which this build does not link.  What the reference asks of GDAL
there is small and fully specified by TIFF 6.0 + the GeoTIFF 1.1 key directory:

* ``write_geotiff``: what ``driver.Create(path, w, h, 1, type, ["COMPRESS=DEFLATE", "TILED=YES"])``
  followed by ``SetGeoTransform / SetSpatialRef / SetNoDataValue / WriteArray`` leaves on disk --
  256 x 256 tiles, Adobe-deflate (code 8), ModelPixelScale + ModelTiepoint, a GeoKey directory with
  the EPSG code, the ``GDAL_NODATA`` ASCII tag (42113); BigTIFF when the file passes 4 GB.
* ``GeoTiff``: header of a TIFF / BigTIFF (both byte orders; strips or tiles; uncompressed or
  deflate / LZW; horizontal and floating-point predictors; chunky or planar samples) and ``read_window``, which inflates only
  the tiles a request touches.
* ``write_vrt`` / ``Mosaic``: the VRT ``gdal.BuildVRT`` makes of equally-gridded tiles
  (`raster/sinks.py:126-145`) and its reader (sources pasted by their ``DstRect``).

Host-side I/O ("I/O-bound"): tiles are inflated / deflated on a thread pool (zlib
releases the GIL); the arrays it hands over enter the CUDA path through the same window upload +
nearest-neighbour gather as ``MemorySource``.
"""
import math
import os
import re
import struct
import zlib
from concurrent.futures import ThreadPoolExecutor
from xml.etree import ElementTree
from xml.sax.saxutils import escape

import numpy as np

from . import utils

__all__ = ["write_geotiff", "GeoTiff", "write_vrt", "Mosaic", "open_raster"]

TILE = 256          # GDAL's default block size for TILED=YES
_BYTE, _ASCII, _SHORT, _LONG, _RATIONAL = 1, 2, 3, 4, 5
_DOUBLE, _LONG8 = 12, 16
_TYPE_FORMAT = {1: "B", 2: "c", 3: "H", 4: "I", 5: "II", 6: "b", 7: "B", 8: "h", 9: "i", 10: "ii",
                11: "f", 12: "d", 16: "Q", 17: "q", 18: "Q"}
_SAMPLE_FORMAT = {"u": 1, "i": 2, "f": 3}
_GDAL_TYPE_NAMES = {"u1": "Byte", "i1": "Int8", "u2": "UInt16", "i2": "Int16", "u4": "UInt32",
                    "i4": "Int32", "u8": "UInt64", "i8": "Int64", "f4": "Float32", "f8": "Float64"}
_POOL = None


def _pool():
    global _POOL
    if _POOL is None:
        _POOL = ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1))
    return _POOL


def _epsg_code(projection):
    match = re.fullmatch(r"EPSG:(\d+)", str(projection).strip(), flags=re.IGNORECASE)
    return int(match.group(1)) if match else None


def _nodata_text(value, dtype):
    """GDAL prints the band's no data value with %.18g (gtiffdataset: WriteNoDataValue)."""
    if dtype.kind == "f":
        value = float(value)
        return "nan" if math.isnan(value) else "%.18g" % value
    return str(int(value))


def _geo_keys(projection):
    """(GeoKeyDirectory shorts, GeoAsciiParams text or None) for an EPSG code or a WKT string."""
    code = _epsg_code(projection)
    if code is not None:
        geographic = utils.is_geographic(projection)
        keys = [(1024, 0, 1, 2 if geographic else 1), (1025, 0, 1, 1),
                (2048 if geographic else 3072, 0, 1, code)]
        ascii_params = None
    else:   # a WKT: user-defined CRS, the definition travels as an "ESRI PE String" citation
        ascii_params = "ESRI PE String = {}|".format(projection)
        keys = [(1024, 0, 1, 1), (1025, 0, 1, 1), (1026, 34737, len(ascii_params), 0), (3072, 0, 1, 32767)]
    directory = [1, 1, 0, len(keys)]
    for key in keys:
        directory.extend(key)
    return directory, ascii_params


def _predict(block, predictor):
    """Forward predictor of one (rows, width) block of little-endian samples -> bytes."""
    if predictor == 2:       # horizontal differencing of the samples' bit patterns
        words = block.view(np.dtype("<u{}".format(block.dtype.itemsize)))
        out = words.copy()
        out[:, 1:] = words[:, 1:] - words[:, :-1]
        return out.tobytes()
    # floating-point predictor (TIFF Technical Note 3): byte planes, most significant first, then
    # byte-wise horizontal differencing over the whole row
    rows, width = block.shape
    planes = block.astype(block.dtype.newbyteorder(">")).view("u1").reshape(rows, width, block.dtype.itemsize)
    line = np.ascontiguousarray(planes.transpose(0, 2, 1)).reshape(rows, -1)
    out = line.copy()
    out[:, 1:] = line[:, 1:] - line[:, :-1]
    return out.tobytes()


def write_geotiff(path, values, geo_transform, projection, no_data_value=None, compress=True, tile=TILE,
                  bigtiff=None, predictor=1):
    """Write ``values`` ((h, w) or (bands, h, w)) as a tiled GeoTIFF; see the module docstring.
    ``bigtiff``: None = when the file would pass 4 GB (GDAL's BIGTIFF=IF_NEEDED), True / False = forced.
    ``predictor``: 1 = none (what the reference's sink asks for), 2 = horizontal differencing,
    3 = floating-point predictor (float rasters)."""
    values = np.asarray(values)
    if values.ndim == 2:
        values = values[np.newaxis]
    if values.ndim != 3:
        raise ValueError("expected a (bands, height, width) array")
    if values.dtype == bool:
        values = values.view("u1")
    dtype = values.dtype.newbyteorder("<")
    if dtype.str[1:] not in _GDAL_TYPE_NAMES:
        raise ValueError("Unsupported dtype '{}' for GeoTIFF".format(values.dtype))
    bands, height, width = values.shape
    tiles_x, tiles_y = -(-width // tile), -(-height // tile)
    if predictor not in (1, 2, 3) or (predictor == 3 and dtype.kind != "f"):
        raise ValueError("predictor {} does not apply to dtype '{}'".format(predictor, values.dtype))

    def encode(index):
        band, rest = divmod(index, tiles_y * tiles_x)
        ty, tx = divmod(rest, tiles_x)
        block = np.zeros((tile, tile), dtype=dtype)
        part = values[band, ty * tile:(ty + 1) * tile, tx * tile:(tx + 1) * tile]
        block[:part.shape[0], :part.shape[1]] = part
        raw = block.tobytes() if predictor == 1 else _predict(block, predictor)
        return zlib.compress(raw, 6) if compress else raw

    n_tiles = bands * tiles_y * tiles_x
    chunks = list(_pool().map(encode, range(n_tiles))) if n_tiles > 1 else [encode(0)]
    big = sum(len(c) for c in chunks) + 16 * n_tiles + 4096 > 0xFFFF0000 if bigtiff is None else bool(bigtiff)

    tags = [
        (256, _LONG, [width]), (257, _LONG, [height]),
        (258, _SHORT, [dtype.itemsize * 8] * bands),
        (259, _SHORT, [8 if compress else 1]), (262, _SHORT, [1]),
        (277, _SHORT, [bands]), (284, _SHORT, [1 if bands == 1 else 2]),
        (322, _LONG, [tile]), (323, _LONG, [tile]),
        (324, _LONG8 if big else _LONG, None), (325, _LONG8 if big else _LONG, [len(c) for c in chunks]),
        (339, _SHORT, [_SAMPLE_FORMAT[dtype.kind]] * bands),
    ]
    if bands > 1:
        tags.append((338, _SHORT, [0] * (bands - 1)))
    if predictor != 1:
        tags.append((317, _SHORT, [predictor]))
    p, a, _, q, _, d = [float(x) for x in geo_transform]
    tags.append((33550, _DOUBLE, [abs(a), abs(d), 0.0]))
    tags.append((33922, _DOUBLE, [0.0, 0.0, 0.0, p, q, 0.0]))
    if projection is not None:
        directory, ascii_params = _geo_keys(utils.get_epsg_or_wkt(projection))
        tags.append((34735, _SHORT, directory))
        if ascii_params is not None:
            tags.append((34737, _ASCII, ascii_params))
    if no_data_value is not None:
        tags.append((42113, _ASCII, _nodata_text(no_data_value, dtype)))
    tags.sort(key=lambda t: t[0])

    head = 16 if big else 8
    offsets, at = [], head
    for chunk in chunks:
        offsets.append(at)
        at += len(chunk) + (len(chunk) & 1)
    ifd_at = at
    entry, word, count_fmt = (20, 8, "<Q") if big else (12, 4, "<H")
    ifd_size = struct.calcsize(count_fmt) + entry * len(tags) + word
    extra_at = ifd_at + ifd_size
    ifd, extra = bytearray(), bytearray()
    ifd += struct.pack(count_fmt, len(tags))
    for tag, kind, data in tags:
        if tag == 324:
            data = offsets
        if kind == _ASCII:
            payload = data.encode("ascii", "replace") + b"\0"
            count = len(payload)
        else:
            payload = struct.pack("<%d%s" % (len(data), _TYPE_FORMAT[kind]), *data)
            count = len(data)
        ifd += struct.pack("<HHQ" if big else "<HHI", tag, kind, count)
        if len(payload) <= word:
            ifd += payload.ljust(word, b"\0")
        else:
            ifd += struct.pack("<Q" if big else "<I", extra_at + len(extra))
            extra += payload + (b"\0" if len(payload) & 1 else b"")
    ifd += b"\0" * word    # no further directory
    tmp = path + ".part"
    with open(tmp, "wb") as f:
        f.write(struct.pack("<2sHHHQ", b"II", 43, 8, 0, ifd_at) if big else struct.pack("<2sHI", b"II", 42, ifd_at))
        for chunk in chunks:
            f.write(chunk)
            if len(chunk) & 1:
                f.write(b"\0")
        f.write(ifd)
        f.write(extra)
    os.replace(tmp, path)    # a reader never sees half a file


def _lzw_decode(data):
    """TIFF's LZW (compression 5): MSB-first codes of 9..12 bits, ClearCode 256, EndOfInformation
    257, the code width grows one entry early.  Pure Python -- a slow path for files of other
    writers (GDAL's own default is uncompressed, the sink of this package writes deflate)."""
    table = [bytes([i]) for i in range(256)] + [b"", b""]
    out = bytearray()
    bits = value = 0
    width, previous = 9, None
    for byte in data:
        value = (value << 8) | byte
        bits += 8
        while bits >= width:
            bits -= width
            code = (value >> bits) & ((1 << width) - 1)
            if code == 257:
                return bytes(out)
            if code == 256:
                del table[258:]
                width, previous = 9, None
                continue
            if previous is None:
                entry = table[code]
            elif code < len(table):
                entry = table[code]
                table.append(previous + entry[:1])
            else:
                entry = previous + previous[:1]
                table.append(entry)
            out += entry
            previous = entry
            if len(table) >= (1 << width) - 1 and width < 12:
                width += 1
        value &= (1 << bits) - 1
    return bytes(out)


class GeoTiff(object):
    """Header of a (Big)TIFF file; pixels are read per window."""

    def __init__(self, path):
        self.path = path
        with open(path, "rb") as f:
            head = f.read(16)
            if head[:2] == b"II":
                self._e = "<"
            elif head[:2] == b"MM":
                self._e = ">"
            else:
                raise IOError("'{}' is not a TIFF file".format(path))
            magic = struct.unpack(self._e + "H", head[2:4])[0]
            if magic == 42:
                self._big, ifd_at = False, struct.unpack(self._e + "I", head[4:8])[0]
            elif magic == 43:
                self._big, ifd_at = True, struct.unpack(self._e + "Q", head[8:16])[0]
            else:
                raise IOError("'{}' is not a TIFF file".format(path))
            self.tags = self._read_directory(f, ifd_at)
        t = self.tags
        self.width, self.height = int(t[256][0]), int(t[257][0])
        self.bands = int(t.get(277, [1])[0])
        bits = int(t.get(258, [1])[0])
        kind = {1: "u", 2: "i", 3: "f"}.get(int(t.get(339, [1])[0]), "u")
        if bits % 8 or any(int(b) != bits for b in t.get(258, [bits])):
            raise NotImplementedError("TIFF samples of {} bits".format(t.get(258)))
        self.dtype = np.dtype("{}{}".format(kind, bits // 8))
        self.compression = int(t.get(259, [1])[0])
        if self.compression not in (1, 5, 8, 32946):
            raise NotImplementedError("TIFF compression scheme {} (only none, LZW and deflate)".format(self.compression))
        self.predictor = int(t.get(317, [1])[0])
        if self.predictor not in (1, 2, 3) or (self.predictor == 3 and self.dtype.kind != "f"):
            raise NotImplementedError("TIFF predictor {}".format(self.predictor))
        self.planar = int(t.get(284, [1])[0]) if self.bands > 1 else 2
        if 322 in t:
            self.block_w, self.block_h = int(t[322][0]), int(t[323][0])
            self._offsets, self._counts = t[324], t[325]
        else:
            self.block_w = self.width
            self.block_h = min(int(t.get(278, [self.height])[0]), self.height)
            self._offsets, self._counts = t[273], t[279]
        self.blocks_x = -(-self.width // self.block_w)
        self.blocks_y = -(-self.height // self.block_h)
        self.geo_transform = self._geo_transform()
        self.projection = self._projection()
        self.no_data_value = self._no_data_value()

    shape = property(lambda self: (self.bands, self.height, self.width))

    def _read_directory(self, f, at):
        e = self._e
        f.seek(at)
        if self._big:
            n = struct.unpack(e + "Q", f.read(8))[0]
            raw, size, word, fmt = f.read(20 * n), 20, 8, e + "HHQ"
        else:
            n = struct.unpack(e + "H", f.read(2))[0]
            raw, size, word, fmt = f.read(12 * n), 12, 4, e + "HHI"
        tags = {}
        for i in range(n):
            rec = raw[i * size:(i + 1) * size]
            tag, kind, count = struct.unpack(fmt, rec[:size - word])
            if kind not in _TYPE_FORMAT:
                continue
            item = _TYPE_FORMAT[kind]
            nbytes = struct.calcsize("=" + item) * count
            if nbytes <= word:
                payload = rec[size - word:size - word + nbytes]
            else:
                where = struct.unpack(e + ("Q" if self._big else "I"), rec[size - word:])[0]
                f.seek(where)
                payload = f.read(nbytes)
            if kind == _ASCII:
                tags[tag] = payload.split(b"\0")[0].decode("latin-1")
            elif len(item) == 1 and count > 64:
                tags[tag] = np.frombuffer(payload, dtype=np.dtype(e + {"B": "u1", "H": "u2", "I": "u4", "Q": "u8", "b": "i1",
                                          "h": "i2", "i": "i4", "q": "i8", "f": "f4", "d": "f8"}[item]))
            else:
                tags[tag] = struct.unpack(e + item * count, payload)
        return tags

    def _geo_transform(self):
        t = self.tags
        if 33550 in t and 33922 in t:
            sx, sy = float(t[33550][0]), float(t[33550][1])
            i, j, _, x, y, _ = [float(v) for v in t[33922][:6]]
            gt = [x - i * sx, sx, 0.0, y + j * sy, 0.0, -sy]
        elif 34264 in t:
            m = [float(v) for v in t[34264]]
            gt = [m[3], m[0], m[1], m[7], m[4], m[5]]
        else:
            return None
        if self._geo_key(1025) == 2:       # PixelIsPoint: the tiepoint is a cell centre
            gt[0] -= 0.5 * gt[1]
            gt[3] -= 0.5 * gt[5]
        return tuple(gt)

    def _geo_key(self, wanted):
        directory = self.tags.get(34735)
        if directory is None:
            return None
        for k in range(int(directory[3])):
            key, where, count, value = [int(v) for v in directory[4 + 4 * k:8 + 4 * k]]
            if key != wanted:
                continue
            if where == 0:
                return value
            if where == 34737:
                return self.tags.get(34737, "")[value:value + count].rstrip("|")
            if where == 34736:
                return float(self.tags[34736][value])
        return None

    def _projection(self):
        for key in (3072, 2048):
            code = self._geo_key(key)
            if isinstance(code, int) and 0 < code < 32767:
                return "EPSG:{}".format(code)
        citation = self._geo_key(1026)
        if isinstance(citation, str) and citation.startswith("ESRI PE String = "):
            return citation[len("ESRI PE String = "):]
        return None

    def _no_data_value(self):
        text = self.tags.get(42113)
        if not text:
            return None
        try:
            return float(text.strip())
        except ValueError:
            return None

    def metadata(self, band):
        """The ``metadata`` item of a band in GDAL's metadata tag (42112), if any."""
        text = self.tags.get(42112)
        if not text:
            return None
        try:
            root = ElementTree.fromstring(text)
        except ElementTree.ParseError:
            return None
        for item in root.iter("Item"):
            if item.get("name") == "metadata" and int(item.get("sample", -1)) == band:
                return item.text
        return None

    def _block(self, f_path, index, rows, samples):
        """Decoded block ``index`` as (rows, block_w, samples) in native byte order."""
        offset, count = int(self._offsets[index]), int(self._counts[index])
        with open(f_path, "rb") as f:
            f.seek(offset)
            raw = f.read(count)
        if self.compression == 5:
            raw = _lzw_decode(raw)
        elif self.compression != 1:
            raw = zlib.decompress(raw)
        if self.predictor == 3:   # floating-point predictor: undo the byte differencing and the byte planes
            size = self.dtype.itemsize
            line = np.frombuffer(raw, dtype="u1", count=rows * self.block_w * samples * size)
            line = np.cumsum(line.reshape(rows, -1, samples), axis=1, dtype="u1").reshape(rows, size, -1)
            block = np.ascontiguousarray(line.transpose(0, 2, 1)).view(self.dtype.newbyteorder(">"))
            return block.reshape(rows, self.block_w, samples).astype(self.dtype)
        block = np.frombuffer(raw, dtype=self.dtype.newbyteorder(self._e),
                              count=rows * self.block_w * samples).reshape(rows, self.block_w, samples)
        if self.predictor == 2:    # horizontal differencing, on the samples' bit patterns
            words = np.dtype("u{}".format(self.dtype.itemsize))
            block = np.cumsum(block.view(words.newbyteorder(self._e)), axis=1, dtype=words).view(
                self.dtype.newbyteorder("="))
        return block

    def read_window(self, b0, b1, r0, r1, c0, c1):
        """Samples [b0:b1, r0:r1, c0:c1] as a C-contiguous native array; only the blocks the window
        touches are read and inflated."""
        out = np.empty((b1 - b0, r1 - r0, c1 - c0), dtype=self.dtype)
        if out.size == 0:
            return out
        bx0, bx1 = c0 // self.block_w, (c1 - 1) // self.block_w + 1
        by0, by1 = r0 // self.block_h, (r1 - 1) // self.block_h + 1
        per_plane = self.blocks_x * self.blocks_y
        planes = range(b0, b1) if self.planar == 2 else [None]
        jobs = [(plane, by, bx) for plane in planes for by in range(by0, by1) for bx in range(bx0, bx1)]
        tiled = 322 in self.tags

        def paste(job):
            plane, by, bx = job
            rows = self.block_h if tiled else min(self.block_h, self.height - by * self.block_h)
            index = by * self.blocks_x + bx + (0 if plane is None else plane * per_plane)
            block = self._block(self.path, index, rows, self.bands if plane is None else 1)
            y0, x0 = by * self.block_h, bx * self.block_w
            ya, yb = max(r0, y0), min(r1, y0 + rows)
            xa, xb = max(c0, x0), min(c1, x0 + self.block_w)
            part = block[ya - y0:yb - y0, xa - x0:xb - x0]
            if plane is None:
                out[:, ya - r0:yb - r0, xa - c0:xb - c0] = np.moveaxis(part[:, :, b0:b1], 2, 0)
            else:
                out[plane - b0, ya - r0:yb - r0, xa - c0:xb - c0] = part[:, :, 0]

        if len(jobs) > 1:
            list(_pool().map(paste, jobs))
        else:
            paste(jobs[0])
        return out

    def read(self):
        return self.read_window(0, self.bands, 0, self.height, 0, self.width)


def write_vrt(target, paths):
    """Mosaic of GeoTIFF tiles on one grid as a VRT next to them (what ``gdal.BuildVRT`` writes for
    the tiles of a RasterFileSink: union of the extents, the tiles' own resolution)."""
    tiles = [GeoTiff(p) for p in sorted(paths)]
    first = tiles[0]
    if first.geo_transform is None:
        raise IOError("'{}' carries no georeference".format(first.path))
    a, d = first.geo_transform[1], first.geo_transform[5]
    for t in tiles[1:]:
        if t.geo_transform is None or t.dtype != first.dtype or t.bands != first.bands or \
                not math.isclose(t.geo_transform[1], a, rel_tol=1e-9) or \
                not math.isclose(t.geo_transform[5], d, rel_tol=1e-9):
            raise IOError("'{}' does not match the grid or type of '{}'".format(t.path, first.path))
    x_min = min(t.geo_transform[0] for t in tiles)
    y_max = max(t.geo_transform[3] for t in tiles)
    x_max = max(t.geo_transform[0] + t.width * a for t in tiles)
    y_min = min(t.geo_transform[3] + t.height * d for t in tiles)
    width, height = int(round((x_max - x_min) / a)), int(round((y_min - y_max) / d))
    type_name = _GDAL_TYPE_NAMES[first.dtype.str[1:]]
    base = os.path.dirname(os.path.abspath(target))
    lines = ['<VRTDataset rasterXSize="{}" rasterYSize="{}">'.format(width, height)]
    if first.projection is not None:
        lines.append("  <SRS>{}</SRS>".format(escape(first.projection)))
    lines.append("  <GeoTransform>{}</GeoTransform>".format(
        ", ".join("%.16e" % v for v in (x_min, a, 0.0, y_max, 0.0, d))))
    for band in range(1, first.bands + 1):
        lines.append('  <VRTRasterBand dataType="{}" band="{}">'.format(type_name, band))
        nodata = first.tags.get(42113)
        if nodata:
            lines.append("    <NoDataValue>{}</NoDataValue>".format(nodata.strip()))
        for t in tiles:
            rel = os.path.relpath(os.path.abspath(t.path), base)
            x_off = int(round((t.geo_transform[0] - x_min) / a))
            y_off = int(round((t.geo_transform[3] - y_max) / d))
            lines += [
                "    <ComplexSource>",
                '      <SourceFilename relativeToVRT="1">{}</SourceFilename>'.format(escape(rel)),
                "      <SourceBand>{}</SourceBand>".format(band),
                '      <SourceProperties RasterXSize="{}" RasterYSize="{}" DataType="{}" BlockXSize="{}" BlockYSize="{}" />'.format(
                    t.width, t.height, type_name, t.block_w, t.block_h),
                '      <SrcRect xOff="0" yOff="0" xSize="{}" ySize="{}" />'.format(t.width, t.height),
                '      <DstRect xOff="{}" yOff="{}" xSize="{}" ySize="{}" />'.format(x_off, y_off, t.width, t.height),
            ]
            if nodata:
                lines.append("      <NODATA>{}</NODATA>".format(nodata.strip()))
            lines.append("    </ComplexSource>")
        lines.append("  </VRTRasterBand>")
    lines.append("</VRTDataset>")
    with open(target, "w") as f:
        f.write("\n".join(lines) + "\n")


class Mosaic(object):
    """Reader of a VRT whose sources are unscaled rectangles of GeoTIFF files."""

    def __init__(self, path):
        self.path = path
        root = ElementTree.parse(path).getroot()
        if root.tag != "VRTDataset":
            raise IOError("'{}' is not a VRT".format(path))
        self.width, self.height = int(root.get("rasterXSize")), int(root.get("rasterYSize"))
        srs = root.find("SRS")
        self.projection = None if srs is None or not srs.text else utils.get_epsg_or_wkt(srs.text)
        gt = root.find("GeoTransform")
        self.geo_transform = None if gt is None else tuple(float(v) for v in gt.text.split(","))
        bands = root.findall("VRTRasterBand")
        self.bands = len(bands)
        names = {v: k for k, v in _GDAL_TYPE_NAMES.items()}
        self.dtype = np.dtype(names[bands[0].get("dataType")])
        nodata = bands[0].find("NoDataValue")
        self.no_data_value = None if nodata is None else float(nodata.text)
        base = os.path.dirname(os.path.abspath(path))
        self._files = {}
        self._sources = []      # (band index, file, source band index, src rect, dst rect)
        for index, band in enumerate(bands):
            for node in list(band.findall("ComplexSource")) + list(band.findall("SimpleSource")):
                name = node.find("SourceFilename")
                file_path = name.text if name.get("relativeToVRT") != "1" else os.path.join(base, name.text)
                src, dst = node.find("SrcRect"), node.find("DstRect")
                rect = lambda r: tuple(int(round(float(r.get(k)))) for k in ("xOff", "yOff", "xSize", "ySize"))
                if rect(src)[2:] != rect(dst)[2:]:
                    raise NotImplementedError("VRT sources that are rescaled")
                self._sources.append((index, file_path, int(node.find("SourceBand").text) - 1, rect(src), rect(dst)))

    shape = property(lambda self: (self.bands, self.height, self.width))

    def metadata(self, band):
        return None

    def _file(self, path):
        if path not in self._files:
            self._files[path] = GeoTiff(path)
        return self._files[path]

    def read_window(self, b0, b1, r0, r1, c0, c1):
        fill = 0 if self.no_data_value is None else self.no_data_value
        out = np.full((b1 - b0, r1 - r0, c1 - c0), fill, dtype=self.dtype)
        for band, path, src_band, (sx, sy, _, _), (dx, dy, w, h) in self._sources:
            if not b0 <= band < b1:
                continue
            xa, xb, ya, yb = max(c0, dx), min(c1, dx + w), max(r0, dy), min(r1, dy + h)
            if xa >= xb or ya >= yb:
                continue
            part = self._file(path).read_window(src_band, src_band + 1, ya - dy + sy, yb - dy + sy,
                                                xa - dx + sx, xb - dx + sx)
            out[band - b0, ya - r0:yb - r0, xa - c0:xb - c0] = part[0]
        return out

    def read(self):
        return self.read_window(0, self.bands, 0, self.height, 0, self.width)


def open_raster(path):
    """GeoTiff or Mosaic, by content."""
    with open(path, "rb") as f:
        head = f.read(64)
    if head[:2] in (b"II", b"MM"):
        return GeoTiff(path)
    if b"<VRTDataset" in head:
        return Mosaic(path)
    raise IOError("'{}' is neither a TIFF nor a VRT file".format(path))
