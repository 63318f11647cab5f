"""ORACLE / TEST INFRASTRUCTURE ONLY.  Restated Merlin v1.0 transcript (STROBE-128
over Keccak-f[1600]) standing in for the un-vendored third-party package the
reference pins at ``poetry.lock:255-269`` (``merlin`` 0.1.0, subdir of
github.com/nalinbhardwaj/curdleproofs.pie @ 805d0678).  Call sites in the
reference: ``transcript.py:3,58-75`` and ``prover.py:53``.  Pinned by Merlin's
published conformance vector (tests/test_oracle_pins.py)."""
from .merlin_transcript import MerlinTranscript

__all__ = ["MerlinTranscript"]
