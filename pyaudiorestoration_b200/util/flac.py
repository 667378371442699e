"""A self-contained FLAC decoder (numpy + a tight Python loop over residual samples).

Size and speed: the bit tables cover a bounded window of the file (~100 MB of memory however long it is), but the
residual loop is Python: roughly 10 s of mono 44.1 kHz audio per second.  Fine for the reference's sample fixtures and
short takes; ``util.io_ops.read_file`` uses ``soundfile`` instead whenever that package is importable.

The reference reads audio through soundfile/libsndfile (util/io_ops.py:7-16), which is not available
in this image; its sample fixtures (samples/*.flac) and the "*.flac *.wav" file dialogs of its tools
need a decoder either side of the hot path.  Implements the FLAC format
subset libFLAC produces: fixed and variable block sizes, CONSTANT / VERBATIM / FIXED / LPC subframes,
Rice and Rice2 residual coding with escape partitions, wasted bits, all stereo decorrelation modes,
8-32 bits per sample.  Frame CRC-16 values are checked; ``read_flac(..., verify_md5=True)`` also
checks the STREAMINFO MD5 of the decoded audio, which pins the decode bit-exactly.
"""
import hashlib
import struct

import numpy as np

_FIXED_COEFFS = ((), (1,), (2, -1), (3, -3, 1), (4, -6, 4, -1))
_BLOCK_SIZES = {1: 192, 2: 576, 3: 1152, 4: 2304, 5: 4608, 8: 256, 9: 512, 10: 1024, 11: 2048, 12: 4096, 13: 8192,
                14: 16384, 15: 32768}
_SAMPLE_SIZES = {1: 8, 2: 12, 4: 16, 5: 20, 6: 24, 7: 32}


def _crc16_table():
    tab = []
    for i in range(256):
        c = i << 8
        for _ in range(8):
            c = ((c << 1) ^ 0x8005) & 0xFFFF if c & 0x8000 else (c << 1) & 0xFFFF
        tab.append(c)
    return tab


_CRC16 = _crc16_table()


def _crc16(data):
    c = 0
    for b in data:
        c = ((c << 8) & 0xFFFF) ^ _CRC16[(c >> 8) ^ b]
    return c


class _Bits:
    """Bit reader over a byte string with a BOUNDED random-access window: for the bytes of the current window
    ``win[p]`` holds the 32 bits that start at (window-relative) bit p and ``next_one[p]`` the position of the first
    1 bit at or after p.  The window (``WINDOW`` bytes, ~64 bytes of tables per byte) is rebuilt when the cursor gets
    near its end, so memory stays ~100 MB however long the file is; ``ensure`` makes room for a whole frame."""

    WINDOW = 1 << 19

    def __init__(self, data):
        self.data = data
        self.nbits = len(data) * 8
        self.pos = 0                 # absolute bit position
        self.base = 0                # absolute bit position of window bit 0
        self.limit = 0               # window length in bits that is safe to read 40 bits at
        self._load(0, self.WINDOW)

    def _load(self, byte0, nbytes):
        raw = np.frombuffer(self.data, dtype=np.uint8, count=max(0, min(nbytes, len(self.data) - byte0)), offset=byte0)
        n = len(raw)
        padded = np.concatenate([raw, np.zeros(8, np.uint8)]).astype(np.uint64)
        # 40-bit big-endian window at every byte, then one shifted copy per bit offset
        w40 = (padded[0:n] << np.uint64(32)) | (padded[1:n + 1] << np.uint64(24)) | (padded[2:n + 2] << np.uint64(16)) | \
              (padded[3:n + 3] << np.uint64(8)) | padded[4:n + 4]
        win = np.empty((n, 8), dtype=np.uint32)
        for sh in range(8):
            win[:, sh] = ((w40 >> np.uint64(8 - sh)) & np.uint64(0xFFFFFFFF)).astype(np.uint32)
        self.win = win.reshape(-1)
        nb = n * 8
        bits = np.unpackbits(raw)
        idx = np.where(bits == 1, np.arange(nb, dtype=np.int32), np.int32(nb))
        self.next_one = np.minimum.accumulate(idx[::-1])[::-1] if nb else idx
        self.base = byte0 * 8
        self.wbits = nb
        # reads near the end of the window are only safe when the window ends with the data
        self.limit = nb if byte0 + n >= len(self.data) else nb - 64

    def ensure(self, need_bits):
        """Make the window cover ``need_bits`` bits from the cursor (a whole frame) without another reload."""
        rel = self.pos - self.base
        if rel + need_bits > self.limit and self.base + self.wbits < self.nbits:
            self._load(self.pos // 8, max(self.WINDOW, (need_bits + 7) // 8 + 64))

    def _rel(self, k):
        rel = self.pos - self.base
        if rel + k > self.limit and self.base + self.wbits < self.nbits:
            self._load(self.pos // 8, self.WINDOW)
            rel = self.pos - self.base
        return rel

    def u(self, k):
        """Unsigned k-bit field (k <= 32) at the cursor."""
        if k == 0:
            return 0
        v = int(self.win[self._rel(k)]) >> (32 - k)
        self.pos += k
        return v

    def u_long(self, k):
        v = 0
        while k > 32:
            v = (v << 32) | self.u(32)
            k -= 32
        return (v << k) | self.u(k)

    def s(self, k):
        v = self.u(k) if k <= 32 else self.u_long(k)
        return v - (1 << k) if k and v >> (k - 1) else v

    def unary(self):
        rel = self._rel(64)
        p = int(self.next_one[rel])
        if p >= self.wbits and self.base + self.wbits < self.nbits:
            raise ValueError("FLAC: unary code longer than the decoder's window")
        q = p - rel
        self.pos += q + 1
        return q


def _residual(br, blocksize, order, out):
    method = br.u(2)
    if method > 1:
        raise ValueError("FLAC: reserved residual coding method")
    pbits = 4 if method == 0 else 5
    esc = (1 << pbits) - 1
    porder = br.u(4)
    nparts = 1 << porder
    i = order
    for part in range(nparts):
        count = (blocksize >> porder) - (order if part == 0 else 0)
        k = br.u(pbits)
        if k == esc:
            raw_bits = br.u(5)
            for _ in range(count):
                out[i] = br.s(raw_bits) if raw_bits else 0
                i += 1
            continue
        # the caller's ensure() put the whole frame inside the window: window-relative positions from here on
        win, next_one, top = br.win, br.next_one, br.wbits
        pos = br.pos - br.base
        shift = 32 - k
        if k:
            for _ in range(count):
                p = int(next_one[pos])
                u = ((p - pos) << k) | (int(win[p + 1]) >> shift)
                pos = p + 1 + k
                out[i] = (u >> 1) ^ -(u & 1)
                i += 1
        else:
            for _ in range(count):
                p = int(next_one[pos])
                u = p - pos
                pos = p + 1
                out[i] = (u >> 1) ^ -(u & 1)
                i += 1
        if pos > top:
            raise ValueError("FLAC: residual runs past the decoder's window (corrupt stream?)")
        br.pos = br.base + pos


def _predict(out, order, coeffs, shift):
    """out[order:] holds residuals; in-place LPC synthesis (integer arithmetic, exact)."""
    n = len(out)
    if order == 0:
        return
    vals = out.tolist()
    if order == 1 and shift == 0 and coeffs[0] == 1:
        out[:] = np.cumsum(out, dtype=np.int64)
        return
    c = list(coeffs)
    for i in range(order, n):
        acc = 0
        for j in range(order):
            acc += c[j] * vals[i - 1 - j]
        vals[i] += acc >> shift
    out[:] = vals


def _subframe(br, blocksize, bps):
    if br.u(1):
        raise ValueError("FLAC: subframe padding bit set")
    kind = br.u(6)
    wasted = 0
    if br.u(1):
        wasted = br.unary() + 1
        bps -= wasted
    out = np.zeros(blocksize, dtype=np.int64)
    if kind == 0:
        out[:] = br.s(bps)
    elif kind == 1:
        for i in range(blocksize):
            out[i] = br.s(bps)
    elif 8 <= kind <= 12:
        order = kind - 8
        for i in range(order):
            out[i] = br.s(bps)
        _residual(br, blocksize, order, out)
        _predict(out, order, _FIXED_COEFFS[order], 0)
    elif kind >= 32:
        order = kind - 31
        for i in range(order):
            out[i] = br.s(bps)
        precision = br.u(4) + 1
        if precision == 16:
            raise ValueError("FLAC: invalid LPC precision")
        shift = br.s(5)
        if shift < 0:
            raise ValueError("FLAC: negative LPC shift")
        coeffs = [br.s(precision) for _ in range(order)]
        _residual(br, blocksize, order, out)
        _predict(out, order, coeffs, shift)
    else:
        raise ValueError(f"FLAC: reserved subframe type {kind}")
    if wasted:
        out <<= wasted
    return out


def _utf8_number(br):
    first = br.u(8)
    if first < 0x80:
        return first
    n = 0
    while first & (0x80 >> n):
        n += 1
    v = first & (0x7F >> n)
    for _ in range(n - 1):
        v = (v << 6) | (br.u(8) & 0x3F)
    return v


def decode_flac(data, verify_md5=False):
    """``(samples int32 [frames, channels], samplerate, bits_per_sample)`` of a FLAC byte string."""
    if data[:4] != b"fLaC":
        raise ValueError("not a FLAC stream")
    pos = 4
    info = None
    while True:
        hdr = data[pos]
        size = int.from_bytes(data[pos + 1:pos + 4], "big")
        body = data[pos + 4:pos + 4 + size]
        pos += 4 + size
        if hdr & 0x7F == 0:
            v = int.from_bytes(body[10:18], "big")
            info = {"sr": v >> 44, "channels": ((v >> 41) & 7) + 1, "bps": ((v >> 36) & 31) + 1,
                    "total": v & ((1 << 36) - 1), "md5": body[18:34]}
        if hdr & 0x80:
            break
    if info is None:
        raise ValueError("FLAC: no STREAMINFO block")
    channels, bps_stream = info["channels"], info["bps"]
    frames = data[pos:]
    br = _Bits(frames)
    blocks = []
    total = 0
    nbytes = len(frames)
    while br.pos + 16 <= br.nbits and (info["total"] == 0 or total < info["total"]):
        start = br.pos // 8
        if br.u(14) != 0x3FFE:
            raise ValueError("FLAC: lost frame sync")
        br.u(1)
        br.u(1)                                   # blocking strategy (only changes the meaning of the number)
        bs_code, sr_code = br.u(4), br.u(4)
        ch_assign, ss_code = br.u(4), br.u(3)
        br.u(1)
        _utf8_number(br)
        if bs_code == 6:
            blocksize = br.u(8) + 1
        elif bs_code == 7:
            blocksize = br.u(16) + 1
        elif bs_code in _BLOCK_SIZES:
            blocksize = _BLOCK_SIZES[bs_code]
        else:
            raise ValueError("FLAC: reserved block size code")
        if sr_code == 12:
            br.u(8)
        elif sr_code in (13, 14):
            br.u(16)
        br.u(8)                                   # CRC-8 of the header
        bps = _SAMPLE_SIZES.get(ss_code, bps_stream)
        br.ensure(blocksize * (ch_assign + 1 if ch_assign < 8 else 2) * 64 + 4096)     # the whole frame inside the window
        if ch_assign < 8:
            subs = [_subframe(br, blocksize, bps) for _ in range(ch_assign + 1)]
        elif ch_assign == 8:                      # left / side
            left = _subframe(br, blocksize, bps)
            side = _subframe(br, blocksize, bps + 1)
            subs = [left, left - side]
        elif ch_assign == 9:                      # side / right
            side = _subframe(br, blocksize, bps + 1)
            right = _subframe(br, blocksize, bps)
            subs = [right + side, right]
        elif ch_assign == 10:                     # mid / side
            mid = _subframe(br, blocksize, bps)
            side = _subframe(br, blocksize, bps + 1)
            mid = (mid << 1) | (side & 1)
            subs = [(mid + side) >> 1, (mid - side) >> 1]
        else:
            raise ValueError("FLAC: reserved channel assignment")
        br.pos = (br.pos + 7) & ~7
        end = br.pos // 8
        if end + 2 > nbytes:
            raise ValueError("FLAC: truncated frame")
        crc = br.u(16)
        if _crc16(frames[start:end]) != crc:
            raise ValueError("FLAC: frame CRC-16 mismatch")
        blocks.append(np.stack(subs, axis=1))
        total += blocksize
    pcm = np.concatenate(blocks, axis=0) if blocks else np.zeros((0, channels), np.int64)
    if info["total"]:
        pcm = pcm[:info["total"]]
    if verify_md5 and any(info["md5"]):
        width = (bps_stream + 7) // 8
        if width == 2:
            raw = pcm.astype("<i2").tobytes()
        elif width == 4:
            raw = pcm.astype("<i4").tobytes()
        elif width == 1:
            raw = pcm.astype("i1").tobytes()
        else:
            raw = b"".join(struct.pack("<i", int(v))[:3] for v in pcm.reshape(-1))
        if hashlib.md5(raw).digest() != info["md5"]:
            raise ValueError("FLAC: MD5 of the decoded audio does not match STREAMINFO")
    return pcm.astype(np.int32), info["sr"], bps_stream


def read_flac(path, verify_md5=False):
    """``(signal float32 [frames, channels], samplerate, channels)``, scaled like soundfile's
    ``read(dtype='float32', always_2d=True)`` (integer PCM / 2**(bits-1))."""
    with open(path, "rb") as f:
        data = f.read()
    pcm, sr, bps = decode_flac(data, verify_md5=verify_md5)
    signal = (pcm.astype(np.float64) / float(1 << (bps - 1))).astype(np.float32)
    return signal, sr, pcm.shape[1]
