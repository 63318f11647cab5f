"""ORACLE ONLY.  ``bytes_to_int`` as used by the reference at ``transcript.py:4,72``:
big-endian bytes -> int."""


def bytes_to_int(x: bytes) -> int:
    return int.from_bytes(x, "big")
