"""plonkathon_b200 -- B200-native PLONK/KZG proving hot path behind the reference's Python surface
(0xPARC/plonkathon: curve.py, poly.py, setup.py, prover.py, transcript.py)."""
from .field import FQ, CURVE_ORDER, FIELD_MODULUS  # noqa: F401
from .curve import Scalar, ec_lincomb, ec_mul, G1Point  # noqa: F401
from .poly import Basis, Polynomial  # noqa: F401
from .setup import Setup, VerificationKey  # noqa: F401
from ._lib import Context, PlonkB200Error, default_context  # noqa: F401
from .transcript import Transcript, Message1, Message2, Message3, Message4, Message5  # noqa: F401
from .prover import Prover, Proof  # noqa: F401
