"""ORACLE / TEST INFRASTRUCTURE ONLY -- a restated subset of py-ecc 6.0.0 so the
reference's own modules can be imported in the build container for validation.
Never imported by the product path (plonkathon_b200/)."""
