"""Audio file I/O either side of the hot path: the part of the reference's ``util/io_ops.py``
that ``util.resampling.run`` needs (``read_file`` :7-16, ``write_file`` :19-23) plus the float-WAV
writer ``run`` uses at util/resampling.py:235-237.

The reference delegates to soundfile/libsndfile, which is not installed in this image (SURVEY.md
8f rank 3); this is a self-contained RIFF/WAVE implementation on numpy: PCM 8/16/24/32-bit and
IEEE float 32/64 in, IEEE float32 out (RF64 header when the data chunk passes 4 GiB).
"""
import logging
import os
import struct

import numpy as np

WAVE_FORMAT_PCM = 1
WAVE_FORMAT_IEEE_FLOAT = 3
WAVE_FORMAT_EXTENSIBLE = 0xFFFE


def write_float_wav(path, signal, sr):
    """Write ``signal`` (frames,) or (frames, channels) as an IEEE-float32 WAV -- what
    ``sf.SoundFile(path, 'w+', sr, channels, subtype='FLOAT').write(signal)`` produces
    (util/resampling.py:236-237): 'fmt ' with format tag 3, a 'fact' chunk, then 'data'."""
    data = np.asarray(signal, dtype="<f4")
    if data.ndim == 1:
        data = data[:, None]
    if data.ndim != 2:
        raise ValueError("signal must be (frames,) or (frames, channels)")
    frames, channels = data.shape
    if channels < 1:
        raise ValueError("need at least one channel")
    nbytes = frames * channels * 4
    fmt = struct.pack("<HHIIHHH", WAVE_FORMAT_IEEE_FLOAT, channels, int(sr), int(sr) * channels * 4,
                      channels * 4, 32, 0)
    pad = nbytes & 1
    with open(path, "wb") as f:
        if nbytes + 64 < 0xFFFFFFFF:
            fact = struct.pack("<I", frames)
            riff_size = 4 + (8 + len(fmt)) + (8 + len(fact)) + (8 + nbytes + pad)
            f.write(b"RIFF" + struct.pack("<I", riff_size) + b"WAVE")
            f.write(b"fmt " + struct.pack("<I", len(fmt)) + fmt)
            f.write(b"fact" + struct.pack("<I", len(fact)) + fact)
            f.write(b"data" + struct.pack("<I", nbytes))
        else:
            ds64 = struct.pack("<QQQI", 0, nbytes, frames, 0)
            riff_size = 4 + (8 + len(ds64)) + (8 + len(fmt)) + (8 + nbytes + pad)
            ds64 = struct.pack("<QQQI", riff_size, nbytes, frames, 0)
            f.write(b"RF64" + struct.pack("<I", 0xFFFFFFFF) + b"WAVE")
            f.write(b"ds64" + struct.pack("<I", len(ds64)) + ds64)
            f.write(b"fmt " + struct.pack("<I", len(fmt)) + fmt)
            f.write(b"data" + struct.pack("<I", 0xFFFFFFFF))
        np.ascontiguousarray(data).tofile(f)
        if pad:
            f.write(b"\0")


def _decode_pcm(raw, bits, channels):
    if bits == 8:
        x = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
    elif bits == 16:
        x = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
    elif bits == 24:
        b = np.frombuffer(raw, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
        v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        v = np.where(v & 0x800000, v - 0x1000000, v)
        x = v.astype(np.float32) / 8388608.0
    elif bits == 32:
        x = (np.frombuffer(raw, dtype="<i4").astype(np.float64) / 2147483648.0).astype(np.float32)
    else:
        raise ValueError(f"unsupported PCM bit depth {bits}")
    return x.reshape(-1, channels)


def read_wav(path):
    """(signal[frames, channels] float32, samplerate, channels) of a RIFF/RF64 WAVE file, scaled
    like soundfile's ``read(dtype='float32')`` (integer PCM divided by 2**(bits-1))."""
    with open(path, "rb") as f:
        head = f.read(12)
        if len(head) < 12 or head[:4] not in (b"RIFF", b"RF64") or head[8:12] != b"WAVE":
            raise ValueError(f"{path}: not a RIFF/WAVE file")
        rf64 = head[:4] == b"RF64"
        fmt = None
        data_size64 = None
        while True:
            ch = f.read(8)
            if len(ch) < 8:
                raise ValueError(f"{path}: no data chunk")
            cid, size = ch[:4], struct.unpack("<I", ch[4:])[0]
            if cid == b"ds64":
                body = f.read(size + (size & 1))
                data_size64 = struct.unpack("<Q", body[8:16])[0]
            elif cid == b"fmt ":
                body = f.read(size + (size & 1))
                tag, channels, sr, _, align, bits = struct.unpack("<HHIIHH", body[:16])
                if tag == WAVE_FORMAT_EXTENSIBLE and size >= 26:
                    tag = struct.unpack("<H", body[24:26])[0]
                fmt = (tag, channels, sr, align, bits)
            elif cid == b"data":
                if fmt is None:
                    raise ValueError(f"{path}: data chunk before fmt chunk")
                if rf64 and size == 0xFFFFFFFF and data_size64 is not None:
                    size = data_size64
                tag, channels, sr, align, bits = fmt
                raw = f.read(size)
                usable = len(raw) - len(raw) % max(align, 1)
                raw = raw[:usable]
                if tag == WAVE_FORMAT_IEEE_FLOAT and bits == 32:
                    sig = np.frombuffer(raw, dtype="<f4").reshape(-1, channels).astype(np.float32)
                elif tag == WAVE_FORMAT_IEEE_FLOAT and bits == 64:
                    sig = np.frombuffer(raw, dtype="<f8").reshape(-1, channels).astype(np.float32)
                elif tag == WAVE_FORMAT_PCM:
                    sig = _decode_pcm(raw, bits, channels)
                else:
                    raise ValueError(f"{path}: unsupported WAVE format tag {tag} / {bits} bit")
                return sig, sr, channels
            else:
                f.seek(size + (size & 1), os.SEEK_CUR)


def read_file(audio_path):
    """util/io_ops.py:7-16: ``(signal[frames, channels] float32, samplerate, channels)``.
    WAV here; FLAC through ``soundfile`` when that package is importable (the reference's own reader, libsndfile
    speed), else via ``flac.read_flac`` (self-contained numpy decoder: bounded memory, but a Python loop per residual
    sample -- meant for the sample fixtures and short takes)."""
    logging.info(f"Reading {audio_path}")
    ext = os.path.splitext(audio_path)[1].lower()
    if ext == ".flac":
        try:
            import soundfile as sf                       # not installed in the build image; used where available
            with sf.SoundFile(audio_path) as snd:
                signal = snd.read(always_2d=True, dtype="float32")
                sr, channels = snd.samplerate, snd.channels
        except ImportError:
            from . import flac
            signal, sr, channels = flac.read_flac(audio_path)
    else:
        signal, sr, channels = read_wav(audio_path)
    if len(signal) == 0:
        raise AttributeError(f"Reading {audio_path} failed, file holds no audio frames")
    return signal, sr, channels


def write_file(audio_path, signal, sr, channels, suffix="_out"):
    """util/io_ops.py:19-23."""
    signal = np.asarray(signal)
    if signal.ndim == 1:
        signal = signal[:, None]
    if signal.shape[1] != channels:
        raise ValueError("channel count does not match the signal")
    write_float_wav(f"{os.path.splitext(audio_path)[0]}{suffix}.wav", signal, sr)
    logging.info(f"Wrote {audio_path}")
