"""Minimal TensorBoard scalar writer (no tensorboard / tensorflow dependency).

The reference logs its epoch scalars through ``TensorBoardLogger`` (audio_train.py:115-117; ``train_loss``, ``val_loss``, ``lr`` via
``self.log`` and ``learning_rate``, ``val_pit_sisnr`` via ``logger.experiment.add_scalar``, system/audio_litmodule.py:79-149).  This writes
the same ``events.out.tfevents.*`` record stream TensorBoard reads: TFRecord framing (length, masked CRC-32C of the length, payload,
masked CRC-32C of the payload) around hand-encoded ``Event`` protobufs (wall_time = 1, step = 2, file_version = 3, summary = 5;
``Summary.Value``: tag = 1, simple_value = 2).
"""
from __future__ import annotations

import os
import socket
import struct
import time

_CRC_TABLE = []
for _i in range(256):
    _c = _i
    for _ in range(8):
        _c = (_c >> 1) ^ 0x82F63B78 if _c & 1 else _c >> 1
    _CRC_TABLE.append(_c)


def crc32c(data: bytes) -> int:
    c = 0xFFFFFFFF
    for b in data:
        c = _CRC_TABLE[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def masked_crc(data: bytes) -> int:
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def _varint(n: int) -> bytes:
    out = bytearray()
    n &= (1 << 64) - 1
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _len_delimited(field: int, payload: bytes) -> bytes:
    return _varint((field << 3) | 2) + _varint(len(payload)) + payload


def encode_event(wall_time: float, step: int, *, file_version: str | None = None, scalar: tuple[str, float] | None = None) -> bytes:
    ev = b"\x09" + struct.pack("<d", wall_time) + b"\x10" + _varint(step)
    if file_version is not None:
        ev += _len_delimited(3, file_version.encode())
    if scalar is not None:
        tag, value = scalar
        val = _len_delimited(1, tag.encode()) + b"\x15" + struct.pack("<f", float(value))
        ev += _len_delimited(5, _len_delimited(1, val))
    return ev


class ScalarWriter:
    """``add_scalar(tag, value, step)`` into ``logdir/events.out.tfevents.<time>.<host>`` (the subset of SummaryWriter the reference uses)."""

    def __init__(self, logdir: str):
        os.makedirs(logdir, exist_ok=True)
        self.path = os.path.join(logdir, f"events.out.tfevents.{int(time.time())}.{socket.gethostname()}.{os.getpid()}")
        self._f = open(self.path, "wb")
        self._record(encode_event(time.time(), 0, file_version="brain.Event:2"))

    def _record(self, payload: bytes):
        head = struct.pack("<Q", len(payload))
        self._f.write(head + struct.pack("<I", masked_crc(head)) + payload + struct.pack("<I", masked_crc(payload)))

    def add_scalar(self, tag: str, value: float, step: int):
        self._record(encode_event(time.time(), int(step), scalar=(tag, float(value))))

    def flush(self):
        self._f.flush()

    def close(self):
        if not self._f.closed:
            self._f.close()
