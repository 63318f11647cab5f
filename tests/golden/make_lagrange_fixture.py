"""Writes tests/golden/ptau_lagrange_p0_p4.bin: the first 31 points (domain sizes 1, 2, 4, 8, 16) of section 12
(Lagrange-basis tauG1) of the reference's shipped ceremony file, raw as stored (32-byte little-endian coordinates
in Montgomery form).  Run in the build container, where /root/reference exists:

    python tests/golden/make_lagrange_fixture.py
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from plonkathon_b200.setup import PTAU_SECTION_LAGRANGE_G1, ptau_sections  # noqa: E402

SRC = "/root/reference/test/powersOfTau28_hez_final_11.ptau"
contents = open(SRC, "rb").read()
off, size = ptau_sections(contents)[PTAU_SECTION_LAGRANGE_G1]
assert size == 64 * (2 ** 13 - 1)
open(os.path.join(HERE, "ptau_lagrange_p0_p4.bin"), "wb").write(contents[off:off + 64 * 31])
print("section 12 at", off, "size", size, "-> 31 points written")
