"""Drop-in for the reference's ``transcript.py``: the message dataclasses (transcript.py:8-55) and
``Transcript`` (transcript.py:58-123).  The Merlin/STROBE/Keccak machinery is host code inside
libplonk_b200.so (csrc/transcript.cuh); this class is a thin binding with the reference's method names."""
from __future__ import annotations

import ctypes
from dataclasses import dataclass

from . import _lib
from .curve import Scalar


@dataclass
class Message1:
    a_1: object
    b_1: object
    c_1: object


@dataclass
class Message2:
    z_1: object


@dataclass
class Message3:
    t_lo_1: object
    t_mid_1: object
    t_hi_1: object


@dataclass
class Message4:
    a_eval: Scalar
    b_eval: Scalar
    c_eval: Scalar
    s1_eval: Scalar
    s2_eval: Scalar
    z_shifted_eval: Scalar


@dataclass
class Message5:
    W_z_1: object
    W_zw_1: object


def _n(x) -> int:
    return x.n if hasattr(x, "n") else int(x)


class Transcript:
    def __init__(self, label: bytes):
        h = ctypes.c_void_p()
        _lib.check(_lib.lib().pb200_transcript_create(label, len(label), ctypes.byref(h)))
        self._h = h

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _lib.lib().pb200_transcript_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # MerlinTranscript surface
    def append_message(self, label: bytes, message: bytes) -> None:
        _lib.check(_lib.lib().pb200_transcript_append_message(self._h, label, len(label), message, len(message)))

    def challenge_bytes(self, label: bytes, length: int) -> bytes:
        out = ctypes.create_string_buffer(length)
        _lib.check(_lib.lib().pb200_transcript_challenge_bytes(self._h, label, len(label), out, length))
        return out.raw

    # transcript.py:59-75
    def append(self, label: bytes, item: bytes) -> None:
        self.append_message(label, item)

    def append_scalar(self, label: bytes, item):
        self.append_message(label, _n(item).to_bytes(32, "big"))

    def append_point(self, label: bytes, item):
        self.append_message(label, _n(item[0]).to_bytes(32, "big"))
        self.append_message(label, _n(item[1]).to_bytes(32, "big"))

    def get_and_append_challenge(self, label: bytes) -> Scalar:
        out = ctypes.create_string_buffer(32)
        _lib.check(_lib.lib().pb200_transcript_get_and_append_challenge(self._h, label, len(label), out))
        return Scalar(int.from_bytes(out.raw, "little"))

    # transcript.py:77-123
    def round_1(self, message: Message1):
        self.append_point(b"a_1", message.a_1)
        self.append_point(b"b_1", message.b_1)
        self.append_point(b"c_1", message.c_1)
        beta = self.get_and_append_challenge(b"beta")
        gamma = self.get_and_append_challenge(b"gamma")
        return beta, gamma

    def round_2(self, message: Message2):
        self.append_point(b"z_1", message.z_1)
        alpha = self.get_and_append_challenge(b"alpha")
        fft_cofactor = self.get_and_append_challenge(b"fft_cofactor")
        return alpha, fft_cofactor

    def round_3(self, message: Message3):
        self.append_point(b"t_lo_1", message.t_lo_1)
        self.append_point(b"t_mid_1", message.t_mid_1)
        self.append_point(b"t_hi_1", message.t_hi_1)
        return self.get_and_append_challenge(b"zeta")

    def round_4(self, message: Message4):
        self.append_scalar(b"a_eval", message.a_eval)
        self.append_scalar(b"b_eval", message.b_eval)
        self.append_scalar(b"c_eval", message.c_eval)
        self.append_scalar(b"s1_eval", message.s1_eval)
        self.append_scalar(b"s2_eval", message.s2_eval)
        self.append_scalar(b"z_shifted_eval", message.z_shifted_eval)
        return self.get_and_append_challenge(b"v")

    def round_5(self, message: Message5):
        self.append_point(b"W_z_1", message.W_z_1)
        self.append_point(b"W_zw_1", message.W_zw_1)
        return self.get_and_append_challenge(b"u")
