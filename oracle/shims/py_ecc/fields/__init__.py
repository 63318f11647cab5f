"""ORACLE / TEST INFRASTRUCTURE ONLY.  Restated subset of ``py_ecc.fields``
(py-ecc 6.0.0, reference pin ``poetry.lock:362-363``).  ``bn128_FQ`` is the class
name the reference's ``test/proof.pickle`` refers to."""
from .field_elements import FQ, FQP, prime_field_inv

BN128_FIELD_MODULUS = 21888242871839275222246405745257275088696311157297823662689037894645226208583


class bn128_FQ(FQ):
    field_modulus = BN128_FIELD_MODULUS


class bn128_FQP(FQP):
    field_modulus = BN128_FIELD_MODULUS


class bn128_FQ2(bn128_FQP):
    degree = 2
    modulus_coeffs = (1, 0)  # w^2 = -1


class bn128_FQ12(bn128_FQP):
    degree = 12
    modulus_coeffs = (82, 0, 0, 0, 0, 0, -18, 0, 0, 0, 0, 0)  # w^12 = 18 w^6 - 82
