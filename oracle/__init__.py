"""ORACLE — test infrastructure only.

CPU restatement of moxime/joint-vae's hot path (reference mounted at /root/reference in the
build container).  Nothing under joint-vae_b200/ may import this package; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.

Parity status: pinned against fixtures generated from the unmodified reference
(tests/golden/*.npz, generator tests/golden/make_golden.py).
"""
