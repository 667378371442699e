"""Minimal FLAC *writer* for tests: VERBATIM or CONSTANT subframes only, optional mid/side stereo,
correct CRC-8 / CRC-16 / MD5.  Lets the decoder's container, channel and sample-size handling be
tested anywhere (the Rice / LPC paths are covered by the reference's sample files where available)."""
import hashlib


class _W:
    def __init__(self):
        self.bits = []

    def put(self, v, n):
        v &= (1 << n) - 1
        self.bits.extend((v >> (n - 1 - i)) & 1 for i in range(n))

    def align(self):
        while len(self.bits) % 8:
            self.bits.append(0)

    def bytes(self):
        self.align()
        return bytes(int("".join(map(str, self.bits[i:i + 8])), 2) for i in range(0, len(self.bits), 8))


def _crc(data, poly, width):
    c, top, mask = 0, 1 << (width - 1), (1 << width) - 1
    for b in data:
        c ^= b << (width - 8)
        for _ in range(8):
            c = ((c << 1) ^ poly) & mask if c & top else (c << 1) & mask
    return c


def write_flac(pcm, sr, bps=16, blocksize=1000, mid_side=False):
    """pcm: list of per-channel integer lists -> FLAC bytes."""
    channels, n = len(pcm), len(pcm[0])
    width = (bps + 7) // 8
    md5 = hashlib.md5(b"".join(int(pcm[c][i]).to_bytes(width, "little", signed=True)
                               for i in range(n) for c in range(channels))).digest()
    w = _W()
    w.put(blocksize, 16); w.put(blocksize, 16); w.put(0, 24); w.put(0, 24)
    w.put(sr, 20); w.put(channels - 1, 3); w.put(bps - 1, 5); w.put(n, 36)
    info = w.bytes() + md5
    out = b"fLaC" + bytes([0x80]) + len(info).to_bytes(3, "big") + info
    ss_code = {8: 1, 12: 2, 16: 4, 20: 5, 24: 6, 32: 7}[bps]
    for fi, s in enumerate(range(0, n, blocksize)):
        bs = min(blocksize, n - s)
        w = _W()
        w.put(0x3FFE, 14); w.put(0, 1); w.put(0, 1)
        w.put(7, 4); w.put(0, 4)                               # 16-bit block size follows; sample rate from STREAMINFO
        w.put(10 if (mid_side and channels == 2) else channels - 1, 4); w.put(ss_code, 3); w.put(0, 1)
        assert fi < 128
        w.put(fi, 8)
        w.put(bs - 1, 16)
        hdr = w.bytes()
        w2 = _W()
        w2.bits = [int(b) for byte in hdr + bytes([_crc(hdr, 0x07, 8)]) for b in f"{byte:08b}"]
        block = [pcm[c][s:s + bs] for c in range(channels)]
        sizes = [bps] * channels
        if mid_side and channels == 2:
            left, right = block
            block = [[(a + b) >> 1 for a, b in zip(left, right)], [a - b for a, b in zip(left, right)]]
            sizes = [bps, bps + 1]
        for ch, size in zip(block, sizes):
            if len(set(ch)) == 1:
                w2.put(0, 1); w2.put(0, 6); w2.put(0, 1); w2.put(ch[0], size)
            else:
                w2.put(0, 1); w2.put(1, 6); w2.put(0, 1)
                for v in ch:
                    w2.put(v, size)
        body = w2.bytes()
        out += body + _crc(body, 0x8005, 16).to_bytes(2, "big")
    return out
