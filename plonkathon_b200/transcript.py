"""Drop-in for the reference's ``transcript.py``: the five message records (transcript.py:8-55) and
``Transcript`` (transcript.py:58-123).  The Merlin/STROBE/Keccak machinery is host code inside
libplonk_b200.so (csrc/transcript.cuh); this module binds it and lays the reference's per-round schedule out
as a table: which fields of a message are absorbed (points as x then y, scalars, all 32-byte big-endian) and
which challenges are drawn afterwards."""
from __future__ import annotations

import ctypes
from dataclasses import make_dataclass

from . import _lib
from .curve import Scalar

# round -> (message field names in absorption order, kind of those fields, challenge labels drawn afterwards)
SCHEDULE = {
    1: (("a_1", "b_1", "c_1"), "point", ("beta", "gamma")),
    2: (("z_1",), "point", ("alpha", "fft_cofactor")),
    3: (("t_lo_1", "t_mid_1", "t_hi_1"), "point", ("zeta",)),
    4: (("a_eval", "b_eval", "c_eval", "s1_eval", "s2_eval", "z_shifted_eval"), "scalar", ("v",)),
    5: (("W_z_1", "W_zw_1"), "point", ("u",)),
}

# Message1 .. Message5: plain records with exactly the reference's field names and order
Message1, Message2, Message3, Message4, Message5 = (
    make_dataclass("Message%d" % rnd, [(name, object) for name in SCHEDULE[rnd][0]]) for rnd in sorted(SCHEDULE))


def _as_int(x) -> int:
    return x.n if hasattr(x, "n") else int(x)


class Transcript:
    def __init__(self, label: bytes):
        handle = ctypes.c_void_p()
        _lib.check(_lib.lib().pb200_transcript_create(label, len(label), ctypes.byref(handle)))
        self._h = handle

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _lib.lib().pb200_transcript_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # ---- the MerlinTranscript surface the reference's class inherits (transcript.py:3,58)
    def append_message(self, label: bytes, message: bytes) -> None:
        _lib.check(_lib.lib().pb200_transcript_append_message(self._h, label, len(label), message, len(message)))

    def challenge_bytes(self, label: bytes, length: int) -> bytes:
        buf = ctypes.create_string_buffer(length)
        _lib.check(_lib.lib().pb200_transcript_challenge_bytes(self._h, label, len(label), buf, length))
        return buf.raw

    # ---- transcript.py:59-75
    append = append_message

    def append_scalar(self, label: bytes, item) -> None:
        self.append_message(label, _as_int(item).to_bytes(32, "big"))

    def append_point(self, label: bytes, item) -> None:
        for coordinate in (item[0], item[1]):  # the identity (None) is unsupported, as in the reference
            self.append_message(label, _as_int(coordinate).to_bytes(32, "big"))

    def get_and_append_challenge(self, label: bytes) -> Scalar:
        """255 squeezed bytes as a big-endian integer mod r, redrawn while zero, then re-absorbed under the
        same label -- all inside the library."""
        out = ctypes.create_string_buffer(32)
        _lib.check(_lib.lib().pb200_transcript_get_and_append_challenge(self._h, label, len(label), out))
        return Scalar(int.from_bytes(out.raw, "little"))

    # ---- transcript.py:77-123
    def _round(self, rnd: int, message):
        fields, kind, challenges = SCHEDULE[rnd]
        absorb = self.append_point if kind == "point" else self.append_scalar
        for name in fields:
            absorb(name.encode(), getattr(message, name))
        drawn = tuple(self.get_and_append_challenge(c.encode()) for c in challenges)
        return drawn if len(drawn) > 1 else drawn[0]

    def round_1(self, message):
        return self._round(1, message)

    def round_2(self, message):
        return self._round(2, message)

    def round_3(self, message):
        return self._round(3, message)

    def round_4(self, message):
        return self._round(4, message)

    def round_5(self, message):
        return self._round(5, message)
