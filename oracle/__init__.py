"""CPU oracle for the CoSA CAM -> pseudo-label refinement path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``cosa_b200/`` may import this package; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs do.  See ``oracle/lattice_oracle.c`` (plain C lattice) and ``oracle/reference_port.py``
(torch-CPU / numpy restatement of PAR, labelling and the dense-CRF energy loss).

Parity pin: every function is checked against golden vectors produced by running the
reference's own code in the build container (``tests/golden/make_golden.py``), and the C lattice
additionally against the unmodified reference C++ compiled into ``oracle/_ref/libbf_ref.so``.
"""
