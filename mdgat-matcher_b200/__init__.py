"""mdgat-matcher_b200: B200-native (sm_100a) inference hot path of MDGAT-matcher.

Public surface mirrors the reference (/root/reference/models/mdgat.py, models/superglue.py):
``models.mdgat.MDGAT`` and ``models.superglue.SuperGlue`` with the same constructor dict,
state-dict layout and forward(dict) -> dict contract; the compute sits in
``csrc/`` (hand-written CUDA behind the C ABI declared in include/mdgat_b200.h).
"""
__version__ = '0.1.0'
