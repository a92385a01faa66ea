"""CPU oracle for the MPGAN / GAPT message-passing hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``mpgan_b200/`` imports this package; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may.  It is a plain fp32 PyTorch-on-CPU restatement of the reference algorithm, written
functionally over a ``state_dict`` (no nn.Module copies), each function citing the reference
file:line it follows.  Parity status: PINNED against golden vectors produced by importing the
unmodified reference in the build container (``oracle/make_golden.py`` -> ``tests/golden/``).
"""
