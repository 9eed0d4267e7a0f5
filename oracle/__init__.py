"""CPU oracle for the SemiUHPE rotation-distribution hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker or the
timed CPU baseline.  ``semiuhpe_b200`` never imports this package.

Contents
--------
``so3_oracle.py``       torch-CPU restatement of the reference algorithm
                        (fp32 like the reference; every function also accepts
                        fp64 so it doubles as the exact-arithmetic anchor).
``pytorch3d_restated.py`` restatement of the three ``pytorch3d==0.7.2``
                        functions the reference calls (library absent here:
                        PARITY UNPINNED for those three, see the file header).
``ref_shim.py``         imports the *unmodified* reference from
                        ``/root/reference`` (only possible in the build
                        container) to pin the restatement and to generate the
                        fixtures under ``tests/golden/``.

Pinning status: the restatement is pinned against outputs of the reference
itself, executed in the build container by ``tests/golden/make_golden.py``
(committed together with the vectors it wrote).  The reference has no tests or
golden vectors of its own.
"""
