"""Seeded random-init BioViL ``state_dict`` for the parity tests (TEST INFRASTRUCTURE ONLY).

The generator itself is input data shared with the benchmark, so it lives in the package
(``incremental_multimodal_medical_learning_ii_b200/synthetic_weights.py``); this module re-exports it under the name the
tests and ``oracle/make_golden.py`` have always used.  ``oracle/make_golden.py`` checks keys, shapes and dtypes of the
result against the reference's own ``ImageModel.state_dict()``.
"""
from incremental_multimodal_medical_learning_ii_b200.synthetic_weights import (  # noqa: F401
    make_state_dict, randomize_batchnorm_, state_dict_checksum)
