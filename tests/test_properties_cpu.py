"""Property tests (hypothesis) of the CPU oracle and of the host-side index plans -- size-independent
facts the GPU tests also lean on: linearity and reconstruction of the transform pair, monotone positions
whose count is the sum of the integer segment lengths, shard plans that tile the job, and the sample
ranges the pipelined host path uploads before it launches a chunk of frames."""
import numpy as np
from hypothesis import given, settings, strategies as st

import oracle
from oracle import oracle_np as onp
from pyaudiorestoration_b200 import dist as pdist

FAST = settings(max_examples=25, deadline=None)


@FAST
@given(st.integers(3, 7), st.integers(1, 4), st.integers(0, 2 ** 31), st.floats(-2, 2))
def test_stft_is_linear_and_invertible(log2n, overlap_pow, seed, a):
    n_fft = 2 << log2n
    hop = max(1, n_fft >> overlap_pow)
    rng = np.random.default_rng(seed)
    n = int(rng.integers(n_fft, 6 * n_fft))
    x, y = rng.standard_normal(n).astype(np.float32), rng.standard_normal(n).astype(np.float32)
    sx, sy = onp.stft_f64(x, n_fft, hop), onp.stft_f64(y, n_fft, hop)
    sz = onp.stft_f64((np.float32(a) * x + y).astype(np.float32), n_fft, hop)
    assert sx.shape == (n // hop + 1, n_fft // 2 + 1)
    assert np.max(np.abs(sz - (np.float32(a) * sx + sy))) <= 1e-5 * max(1.0, np.max(np.abs(sz)))
    if hop <= n_fft // 4:                                   # blackmanharris needs >= 4x overlap to reconstruct
        s = onp.stft_ref(onp.fix_length(x, n + n_fft // 2), n_fft, hop)
        back = onp.istft_ref(s, hop_length=hop, length=n)
        assert np.max(np.abs(back - x)) <= 1e-5 * np.max(np.abs(x))


@FAST
@given(st.integers(0, 2 ** 31), st.integers(2, 40), st.floats(0.3, 3.0))
def test_positions_are_monotone_and_counted_by_the_segments(seed, k, scale):
    rng = np.random.default_rng(seed)
    times = np.concatenate([[0.0], np.cumsum(rng.uniform(20, 400, k))])
    speeds = scale * (1 + 0.3 * rng.uniform(-1, 1, k + 1))
    pos = oracle.speed_to_pos_c(times, speeds, 10 ** 9)                 # end test never fires
    seg = onp.speed_segments(times, speeds)
    assert len(pos) == seg.sum() and np.all(np.diff(pos) > 0)
    assert np.array_equal(pos, onp.speed_to_pos(times, speeds, 10 ** 9))
    n_in = float(pos[len(pos) // 2])                                    # end test fires inside the curve
    cut = oracle.speed_to_pos_c(times, speeds, n_in)
    assert 0 < len(cut) < len(pos) and np.array_equal(cut, pos[:len(cut)])


@FAST
@given(st.integers(1, 8), st.integers(0, 6), st.integers(1, 3), st.integers(1, 200), st.integers(20000, 3_000_000))
def test_time_shards_tile_frames_and_cover_their_halos(world, log_extra, overlap_pow, nt, n):
    n_fft = 64 << log_extra
    hop = max(4, n_fft >> overlap_pow)
    if world > 1 and (n // world) < 4 * max(n_fft, nt + 2):
        world = 1
    frames = []
    for r in range(world):
        sh = pdist.TimeShard(n, n_fft, hop, nt, r, world)
        frames.append((sh.frame0, sh.frame1))
        lo, hi = sh.frame0 * hop - n_fft // 2, (sh.frame1 - 1) * hop + n_fft // 2
        assert max(lo, 0) >= sh.origin and min(hi, n) <= sh.origin + sh.local_len
        assert sh.origin % 4 == 0 and sh.s0 % hop == 0
        assert r == 0 or sh.s0 - (nt + 1) >= sh.origin
        assert r == world - 1 or sh.s1 + nt + 1 <= sh.origin + sh.local_len
    assert frames[0][0] == 0 and frames[-1][1] == n // hop + 1
    assert all(a[1] == b[0] for a, b in zip(frames, frames[1:]))


@FAST
@given(st.integers(4, 9), st.integers(0, 3), st.integers(1, 64), st.integers(1, 40))
def test_pipelined_upload_covers_every_frame_of_a_chunk(log2n, overlap_pow, per_chunk, mult):
    """The host-pointer STFT uploads samples [0, need) before it launches frames [t0, t1)
    (csrc/api.cu: need = (t1-1)*hop - n_fft/2 + n_fft + 8, everything once a frame reaches the
    reflected tail): every index a frame reads, after np.pad's reflection, must lie below `need`."""
    n_fft = 2 << log2n
    hop = max(1, n_fft >> overlap_pow)
    n = mult * n_fft + 17
    T = n // hop + 1
    half = n_fft // 2
    for t0 in range(0, T, per_chunk):
        t1 = min(T, t0 + per_chunk)
        need = (t1 - 1) * hop - half + n_fft + 8
        if need >= n - 1 or n <= 2 * n_fft:
            need = n
        idx = np.arange((t0 * hop) - half, (t1 - 1) * hop - half + n_fft)
        idx = np.where(idx < 0, -idx, idx)
        idx = np.where(idx >= n, 2 * (n - 1) - idx, idx)
        assert idx.min() >= 0 and idx.max() < need
