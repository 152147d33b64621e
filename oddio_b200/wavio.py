"""The steps either side of the hot path: WAV ingest -> Frames (examples/wav.rs:15-46)
and offline render -> 16-bit WAV (examples/offline.rs:25-45). File handling is Python's `wave` module; the
arithmetic of both steps runs on the device when the data allows it (8/16-bit integer files are uploaded as
int16 and scaled by `odb_frames_from_i16`; `render_offline_device` quantises every block with `*_sample_i16`).
The numpy versions below are the reference statements the tests compare the device against; no arithmetic of
the mix itself lives here."""
from __future__ import annotations

import wave
from typing import Callable, Tuple

import numpy as np


def quantize_i16(samples: np.ndarray) -> np.ndarray:
    """`(sample * i16::MAX as f32) as i16` (examples/offline.rs:39): f32 multiply, then Rust's float->int `as`
    cast - truncation toward zero, saturating at the i16 range, NaN -> 0."""
    x = np.asarray(samples, dtype=np.float32) * np.float32(32767.0)
    x = np.nan_to_num(x, nan=0.0, posinf=32767.0, neginf=-32768.0)
    return np.clip(np.trunc(x), -32768.0, 32767.0).astype(np.int16)


def read_wav(path: str) -> Tuple[int, np.ndarray]:
    """Integer-PCM WAV -> (rate, float32 samples of shape (n,) or (n, channels)), scaled as examples/wav.rs:30-37:
    `sample as f32 / (2^(bits-1) - 1) as f32` (8-bit files are unsigned on disk and re-centred first, as hound does)."""
    with wave.open(path, "rb") as w:
        ch, width, rate, n = w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()
        raw = w.readframes(n)
    if width == 1:
        ints = np.frombuffer(raw, dtype=np.uint8).astype(np.int32) - 128
    elif width == 2:
        ints = np.frombuffer(raw, dtype="<i2").astype(np.int32)
    elif width == 3:
        b = np.frombuffer(raw, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
        ints = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        ints = np.where(ints >= 1 << 23, ints - (1 << 24), ints)
    elif width == 4:
        ints = np.frombuffer(raw, dtype="<i4").astype(np.int32)
    else:
        raise ValueError(f"unsupported sample width {width}")
    max_value = np.float32(2 ** (8 * width - 1) - 1)
    x = ints.astype(np.float32) / max_value
    return rate, (x if ch == 1 else x.reshape(-1, ch))


def frames_from_wav(path: str, ctx=None):
    """examples/wav.rs:44-46: decode, `frame_stereo`, `Frames::from_slice` - PCM goes to HBM once. 8- and 16-bit
    files cross the bus as int16 and are scaled on the device (bit-identical to `read_wav`)."""
    from .api import Frames

    with wave.open(path, "rb") as w:
        ch, width, rate, n = w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()
        raw = w.readframes(n) if width in (1, 2) and ch in (1, 2) else None
    if raw is None:
        rate, x = read_wav(path)
        return Frames.from_slice(rate, x, ctx)
    if width == 1:
        ints = (np.frombuffer(raw, dtype=np.uint8).astype(np.int16) - 128)
    else:
        ints = np.frombuffer(raw, dtype="<i2").astype(np.int16)
    return Frames.from_i16(rate, ints if ch == 1 else ints.reshape(-1, ch), 8 * width, ctx)


def render_offline(run: Callable[[int], np.ndarray], path: str, rate: int, block_size: int, n_blocks: int, channels: int = 2) -> int:
    """examples/offline.rs:25-45: `n_blocks` callbacks of `block_size` frames through `run(block_size)` (which returns
    the float32 block of one `oddio::run` call), quantised to 16-bit PCM and appended to `path`. Returns frames written."""
    with wave.open(path, "wb") as w:
        w.setnchannels(channels)
        w.setsampwidth(2)
        w.setframerate(rate)
        for _ in range(n_blocks):
            w.writeframes(quantize_i16(run(block_size)).astype("<i2").tobytes())
    return n_blocks * block_size


def render_offline_device(signal, path: str, rate: int, block_size: int, n_blocks: int) -> int:
    """examples/offline.rs:25-45 with the quantisation on the device: `n_blocks` callbacks of `block_size` frames of
    `signal` (a SpatialScene / Mixer, possibly under Tanh / Reinhard), each leaving the GPU as 16-bit PCM."""
    interval = float(np.float32(1.0) / np.float32(rate))  # oddio::run, lib.rs:91
    with wave.open(path, "wb") as w:
        w.setnchannels(signal.channels)
        w.setsampwidth(2)
        w.setframerate(rate)
        for _ in range(n_blocks):
            w.writeframes(signal.sample_i16(interval, block_size).astype("<i2").tobytes())
    return n_blocks * block_size
