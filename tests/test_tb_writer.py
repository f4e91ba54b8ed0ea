"""TensorBoard event files written by tb_writer.ScalarWriter: record framing with CRC-32C, Event / Summary protobuf fields, read back by an
independent parser (and by tensorboard's own reader when that package is present)."""
import glob
import struct

import pytest


def _read_varint(buf, i):
    n = shift = 0
    while True:
        b = buf[i]
        i += 1
        n |= (b & 0x7F) << shift
        shift += 7
        if not b & 0x80:
            return n, i


def _fields(buf):
    i, out = 0, []
    while i < len(buf):
        key, i = _read_varint(buf, i)
        f, wt = key >> 3, key & 7
        if wt == 0:
            v, i = _read_varint(buf, i)
        elif wt == 1:
            v, i = struct.unpack("<d", buf[i:i + 8])[0], i + 8
        elif wt == 5:
            v, i = struct.unpack("<f", buf[i:i + 4])[0], i + 4
        else:
            n, i = _read_varint(buf, i)
            v, i = buf[i:i + n], i + n
        out.append((f, wt, v))
    return out


def test_event_file_round_trip(tmp_path):
    from audio_only_speech_separation_b200.tb_writer import ScalarWriter, crc32c, masked_crc

    assert crc32c(b"123456789") == 0xE3069283          # the CRC-32C check value
    w = ScalarWriter(str(tmp_path))
    for step, v in enumerate([1.5, -2.25, 3.0]):
        w.add_scalar("val_loss", v, step)
    w.add_scalar("learning_rate", 1e-3, 7)
    w.close()
    (path,) = glob.glob(str(tmp_path / "events.out.tfevents.*"))
    data = open(path, "rb").read()
    i, events = 0, []
    while i < len(data):
        (n,) = struct.unpack("<Q", data[i:i + 8])
        assert struct.unpack("<I", data[i + 8:i + 12])[0] == masked_crc(data[i:i + 8])
        payload = data[i + 12:i + 12 + n]
        assert struct.unpack("<I", data[i + 12 + n:i + 16 + n])[0] == masked_crc(payload)
        events.append(_fields(payload))
        i += 16 + n
    assert len(events) == 5
    assert (3, 2, b"brain.Event:2") in events[0]
    got = []
    for ev in events[1:]:
        step = next(v for f, wt, v in ev if f == 2)
        summary = next(v for f, wt, v in ev if f == 5)
        (value,) = [v for f, wt, v in _fields(summary) if f == 1]
        vf = _fields(value)
        got.append((next(v for f, wt, v in vf if f == 1).decode(), next(v for f, wt, v in vf if f == 2), step))
    assert got == [("val_loss", 1.5, 0), ("val_loss", -2.25, 1), ("val_loss", 3.0, 2), ("learning_rate", pytest.approx(1e-3), 7)]
    try:
        from tensorboard.backend.event_processing.event_file_loader import EventFileLoader
    except ImportError:
        return
    tags = [v.tag for e in EventFileLoader(path).Load() for v in e.summary.value]
    assert tags == ["val_loss"] * 3 + ["learning_rate"]
