"""CPU oracle for the associated-VAE train step.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package.  The product
(``vae_assoc_b200``) never does: it fails loudly when its CUDA library is
missing instead of falling back to anything here.

PARITY UNPINNED: the reference (/root/reference, Python-2 + TensorFlow 0.12-1.x
+ prettytensor) ships no tests, golden vectors or fixtures and cannot be run in
this image (no TensorFlow, no Python 2, no network).  The restatement follows
the reference source line by line (citations in every function) and is
cross-validated by two independent derivations (hand-derived numpy backward vs
torch autograd on the same graph) and central finite differences; only the
Philox4x32-10 generator is pinned to published known-answer vectors
(Random123 kat_vectors).
"""
