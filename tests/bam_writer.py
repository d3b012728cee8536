"""Minimal BAM writer for tests (SAM/BAM specification §4): BGZF blocks + header + alignment records with auxiliary fields."""
import struct
import zlib

CIGAR_OPS = "MIDNSHP=X"


def bgzf_block(data):
    assert len(data) <= 65280
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    c = co.compress(data) + co.flush()
    total = 18 + len(c) + 8
    hdr = b"\x1f\x8b\x08\x04" + b"\x00\x00\x00\x00" + b"\x00\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, total - 1)
    return hdr + c + struct.pack("<II", zlib.crc32(data) & 0xffffffff, len(data))


def _aux(tags):
    out = b""
    for tag, ty, val in tags:
        out += tag.encode() + ty.encode()
        if ty == "Z":
            out += val.encode() + b"\x00"
        elif ty == "i":
            out += struct.pack("<i", val)
        elif ty == "C":
            out += struct.pack("<B", val)
        elif ty == "A":
            out += val.encode()
        elif ty == "B":                                  # ("B", (subtype, [values]))
            st, vals = val
            out += st.encode() + struct.pack("<I", len(vals)) + b"".join(struct.pack({"c": "<b", "C": "<B", "s": "<h", "S": "<H", "i": "<i", "I": "<I", "f": "<f"}[st], v) for v in vals)
        else:
            raise ValueError(ty)
    return out


def record(ref, pos, flag, mapq, qname, cigar, l_seq, tags):
    """pos 1-based; cigar [(length, op char)]."""
    name = qname.encode() + b"\x00"
    cig = b"".join(struct.pack("<I", (n << 4) | CIGAR_OPS.index(op)) for n, op in cigar)
    body = struct.pack("<iiBBHHHiiii", ref, pos - 1, len(name), mapq, 4680, len(cigar), flag, l_seq, -1, -1, 0)
    body += name + cig + b"\x11" * ((l_seq + 1) // 2) + b"\xff" * l_seq + _aux(tags)
    return struct.pack("<i", len(body)) + body


def write_bam(path, refs, records, block=4000, text=b"@HD\tVN:1.0\tSO:coordinate\n"):
    """refs [(name, length)], records = bytes from record().  Small blocks so that records straddle block boundaries."""
    d = b"BAM\x01" + struct.pack("<i", len(text)) + text + struct.pack("<i", len(refs))
    for name, ln in refs:
        nm = name.encode() + b"\x00"
        d += struct.pack("<i", len(nm)) + nm + struct.pack("<i", ln)
    d += b"".join(records)
    with open(path, "wb") as f:
        for o in range(0, len(d), block):
            f.write(bgzf_block(d[o:o + block]))
        f.write(bgzf_block(b""))
