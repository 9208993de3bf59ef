"""TensorBoard scalar summaries, the way the reference's driver writes them: svd_train_val.py:20-21 builds
`summary_pb2.Summary(value=[Summary.Value(tag=name, simple_value=val)])`, :57 opens `tf.summary.FileWriter(logdir=
"/tmp/svd/log")`, :189-192 adds "training_error" / "test_error" at step i.

An event file is a sequence of TFRecords (uint64 length, masked CRC32-C of the length, payload, masked CRC32-C of the
payload) whose payloads are serialized `Event` protos -- {wall_time = 1 (double), step = 2 (int64), file_version = 3
(string) | summary = 5 (message)}.  Both formats are small enough to write by hand, so this module needs neither
TensorFlow nor the tensorboard package; `read_events` reads a file back (tests, and anyone without TensorBoard).
"""
import os
import socket
import struct
import time

_CRC_TABLE = []


def _crc32c(data):
    if not _CRC_TABLE:
        for n in range(256):
            c = n
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
            _CRC_TABLE.append(c)
    crc = 0xFFFFFFFF
    for b in data:
        crc = _CRC_TABLE[(crc ^ b) & 0xFF] ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


def _masked_crc(data):
    c = _crc32c(data)
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


def _varint(n):
    n &= (1 << 64) - 1   # int64 as two's complement, like protobuf
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _len_delimited(field, payload):
    return _varint((field << 3) | 2) + _varint(len(payload)) + payload


def make_scalar_summary(name, val):
    """svd_train_val.py:20-21 -> serialized Summary{value: [Value{tag = 1: name, simple_value = 2: val}]}."""
    value = _len_delimited(1, name.encode("utf-8")) + _varint((2 << 3) | 5) + struct.pack("<f", float(val))
    return _len_delimited(1, value)


def _event(wall_time, step, summary=None, file_version=None):
    ev = _varint((1 << 3) | 1) + struct.pack("<d", wall_time) + _varint((2 << 3) | 0) + _varint(int(step))
    if file_version is not None:
        ev += _len_delimited(3, file_version.encode("utf-8"))
    if summary is not None:
        ev += _len_delimited(5, summary)
    return ev


class FileWriter(object):
    """tf.summary.FileWriter(logdir): add_summary(summary, global_step), flush, close."""

    def __init__(self, logdir, graph=None):
        os.makedirs(logdir, exist_ok=True)
        self.path = os.path.join(logdir, "events.out.tfevents.%010d.%s" % (int(time.time()), socket.gethostname()))
        self._f = open(self.path, "ab")
        self._write(_event(time.time(), 0, file_version="brain.Event:2"))

    def _write(self, payload):
        header = struct.pack("<Q", len(payload))
        self._f.write(header + struct.pack("<I", _masked_crc(header)) + payload + struct.pack("<I", _masked_crc(payload)))

    def add_summary(self, summary, global_step=0):
        self._write(_event(time.time(), global_step, summary=summary))

    def flush(self):
        self._f.flush()

    def close(self):
        if not self._f.closed:
            self._f.close()


def _read_varint(buf, at):
    n, shift = 0, 0
    while True:
        b = buf[at]
        at += 1
        n |= (b & 0x7F) << shift
        if not b & 0x80:
            return n, at
        shift += 7


def _fields(buf):
    at = 0
    while at < len(buf):
        key, at = _read_varint(buf, at)
        field, wire = key >> 3, key & 7
        if wire == 0:
            val, at = _read_varint(buf, at)
        elif wire == 1:
            val, at = buf[at:at + 8], at + 8
        elif wire == 5:
            val, at = buf[at:at + 4], at + 4
        elif wire == 2:
            ln, at = _read_varint(buf, at)
            val, at = buf[at:at + ln], at + ln
        else:
            raise ValueError("wire type %d" % wire)
        yield field, wire, val


def read_events(path):
    """-> [(step, tag, simple_value)] of an event file, checking both CRCs of every record."""
    out = []
    with open(path, "rb") as f:
        raw = f.read()
    at = 0
    while at < len(raw):
        header = raw[at:at + 8]
        (ln,) = struct.unpack("<Q", header)
        (crc_h,) = struct.unpack("<I", raw[at + 8:at + 12])
        payload = raw[at + 12:at + 12 + ln]
        (crc_p,) = struct.unpack("<I", raw[at + 12 + ln:at + 16 + ln])
        if crc_h != _masked_crc(header) or crc_p != _masked_crc(payload):
            raise ValueError("corrupt record at byte %d" % at)
        at += 16 + ln
        step, summary = 0, None
        for field, wire, val in _fields(payload):
            if field == 2:
                step = val
            elif field == 5:
                summary = val
        if summary is None:
            continue
        for field, _, value in _fields(summary):
            if field != 1:
                continue
            tag, sv = None, None
            for f2, _, v2 in _fields(value):
                if f2 == 1:
                    tag = bytes(v2).decode("utf-8")
                elif f2 == 2:
                    (sv,) = struct.unpack("<f", v2)
            out.append((step, tag, sv))
    return out
