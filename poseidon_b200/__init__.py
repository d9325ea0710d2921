"""poseidon_b200 — B200-native (sm_100a) forward/backward engine for Poseidon's scOT model.

Only the hot path of the reference (`scOT/model.py` + the HF swinv2 arithmetic it imports) lives here:
`poseidon_b200.scOT.model` mirrors the reference's `ScOT` / `ScOTConfig` API and runs everything through
the C-ABI library `libscot_b200.so` built from `poseidon_b200/csrc/*.cu`.
"""
__version__ = "0.1.0"
