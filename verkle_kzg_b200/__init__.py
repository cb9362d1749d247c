"""verkle_kzg_b200 — B200 (sm_100a) implementation of the vector-commitment hot path of
SleepingShell/verkle-kzg: KZG / IPA commit, prove, prove_multiproof, verify and batched width-256
verkle node commitment, behind the C ABI of include/vkzg.h (libvkzg.so).

`vector_commit` mirrors the reference's VectorCommitment / VectorCommitmentMultiproof trait surface
(vector-commit/src/lib.rs:70-174, multiproof.rs:90-216) on top of that ABI.
"""
from . import _lib
from ._lib import VkzgError, build
from .engine import Engine, MultiEngine

__all__ = ["Engine", "MultiEngine", "VkzgError", "build", "_lib"]
