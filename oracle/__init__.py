"""CPU oracle for the MLegS transform / nonlinear-term hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``mlegs_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do.

PARITY UNPINNED: the reference (Fortran 2008 + MPI + FFTE + FM + LAPACK)
cannot be compiled in this image (no Fortran compiler, no MPI) and its own
test-suite asserts no numeric value.  The oracle is therefore pinned only
against the analytic known-answers listed in SURVEY.md section 4.3
(tests/test_oracle_analytic.py) and against third-party evaluations of its
tables (tests/test_oracle_independent.py: numpy's Gauss-Legendre rule, mpmath's
legenp for the basis functions and for the xxdx / del2 band tables; LAPACK's
own zgbtrf/zgbtrs through scipy for the band solves; the full transform both
ways against the triple sum the reference's docs/tutorial/initialization.md:154
publishes, summed term by term); see DESIGN.md.
"""
