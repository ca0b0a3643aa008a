"""ORACLE package: CPU restatements of the reference algorithms (test infrastructure only).

Nothing under flaxdiff_b200/ imports this package; only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs do.  See oracle/unet_ref.py and
oracle/diffusion_ref.py for the per-function reference citations.  PARITY UNPINNED (no
reference tests exist; JAX is not installable here) except where stated.
"""
