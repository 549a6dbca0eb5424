"""Minimal bit-stream container.

The reference keeps the coded strings in a Python dict and has no on-disk format; this is the small header that
makes a compress() result portable: architecture, original and padded image size, latent shape, and the lengths of
the per-image y / z strings, followed by the strings themselves.  Layout (little-endian):

    magic  b"ICMB" | u8 version = 1 | u8 len(arch) | arch ascii
    u32 H | u32 W            original (unpadded) image size
    u16 pad_left | u16 pad_right | u16 pad_top | u16 pad_bottom
    u32 zh | u32 zw          the "shape" entry of compress()
    u32 n_images
    n_images x (u32 len_y, u32 len_z)
    y_0 z_0 y_1 z_1 ...      the rANS strings, byte-identical to the reference's for the same (y, mu, sigma)
"""
import struct

MAGIC = b"ICMB"
VERSION = 1


def pack(arch, strings, shape, image_size, pads=(0, 0, 0, 0)):
    """strings = [[y_0..], [z_0..]] as returned by compress(); shape = (zh, zw); image_size = (H, W) before padding."""
    y_strings, z_strings = strings
    if len(y_strings) != len(z_strings):
        raise ValueError("need one y-string and one z-string per image")
    name = arch.encode("ascii")
    if not 0 < len(name) < 256:
        raise ValueError("bad architecture name")
    out = [MAGIC, struct.pack("<BB", VERSION, len(name)), name, struct.pack("<II", int(image_size[0]), int(image_size[1])),
           struct.pack("<HHHH", *(int(p) for p in pads)), struct.pack("<II", int(shape[0]), int(shape[1])), struct.pack("<I", len(y_strings))]
    for y, z in zip(y_strings, z_strings):
        out.append(struct.pack("<II", len(y), len(z)))
    for y, z in zip(y_strings, z_strings):
        out.append(bytes(y))
        out.append(bytes(z))
    return b"".join(out)


def unpack(blob):
    """-> dict(arch, strings=[[y..],[z..]], shape=(zh, zw), image_size=(H, W), pads=(l, r, t, b)); ValueError on damage."""
    mv = memoryview(blob)
    if len(mv) < 6 or bytes(mv[:4]) != MAGIC:
        raise ValueError("not an ICMB container")
    version, n = struct.unpack_from("<BB", mv, 4)
    if version != VERSION:
        raise ValueError(f"unsupported container version {version}")
    o = 6
    need = o + n + 8 + 8 + 8 + 4
    if len(mv) < need:
        raise ValueError("truncated container header")
    arch = bytes(mv[o:o + n]).decode("ascii")
    o += n
    H, W = struct.unpack_from("<II", mv, o); o += 8
    pads = struct.unpack_from("<HHHH", mv, o); o += 8
    zh, zw = struct.unpack_from("<II", mv, o); o += 8
    (count,) = struct.unpack_from("<I", mv, o); o += 4
    if len(mv) < o + 8 * count:
        raise ValueError("truncated container index")
    lens = [struct.unpack_from("<II", mv, o + 8 * i) for i in range(count)]
    o += 8 * count
    if len(mv) != o + sum(a + b for a, b in lens):
        raise ValueError("container size does not match its index")
    ys, zs = [], []
    for ly, lz in lens:
        ys.append(bytes(mv[o:o + ly])); o += ly
        zs.append(bytes(mv[o:o + lz])); o += lz
    return {"arch": arch, "strings": [ys, zs], "shape": (zh, zw), "image_size": (H, W), "pads": tuple(pads)}
