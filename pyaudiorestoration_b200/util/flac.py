"""A self-contained FLAC decoder (numpy + a tight Python loop over residual samples).

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
    """Random-access bit view of a byte string: ``win[p]`` holds the 32 bits that start at bit p,
    ``next_one[p]`` the position of the first 1 bit at or after p."""

    def __init__(self, data):
        raw = np.frombuffer(data, dtype=np.uint8)
        self.nbits = len(raw) * 8
        padded = np.concatenate([raw, np.zeros(8, np.uint8)])
        # 40-bit big-endian window at every byte, then one shifted copy per bit offset
        b = padded.astype(np.uint64)
        n = len(raw)
        w40 = (b[0:n] << np.uint64(32)) | (b[1:n + 1] << np.uint64(24)) | (b[2:n + 2] << np.uint64(16)) | \
              (b[3:n + 3] << np.uint64(8)) | b[4:n + 4]
        win = np.empty((n, 8), dtype=np.uint64)
        for s in range(8):
            win[:, s] = (w40 >> np.uint64(8 - s)) & np.uint64(0xFFFFFFFF)
        self.win = win.reshape(-1)
        bits = np.unpackbits(raw)
        idx = np.where(bits == 1, np.arange(self.nbits, dtype=np.int64), np.int64(self.nbits))
        self.next_one = np.minimum.accumulate(idx[::-1])[::-1]
        self.pos = 0

    def u(self, k):
        """Unsigned k-bit field (k <= 32) at the cursor."""
        if k == 0:
            return 0
        v = int(self.win[self.pos]) >> (32 - k)
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
        p = int(self.next_one[self.pos])
        q = p - self.pos
        self.pos = p + 1
        return q


def _residual(br, blocksize, order, out):
    method = br.u(2)
    if method > 1:
        raise ValueError("FLAC: reserved residual coding method")
    pbits = 4 if method == 0 else 5
    esc = (1 << pbits) - 1
    porder = br.u(4)
    nparts = 1 << porder
    win, next_one = br.win, br.next_one
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
        pos = br.pos
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
        br.pos = pos


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
