"""CPU oracle for the AcinoSet reprojection / trajectory-optimisation hot path.

TEST INFRASTRUCTURE ONLY.  This package is a plain NumPy (fp64) restatement of the
reference's algorithm for the path named in BASELINE.json.  It exists so that the
CUDA path can be checked against something independent; it is never the thing
shipped or measured.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  Nothing
under ``acinoset_b200/`` imports it.

Parity pinning: every function here is checked (tests/test_oracle_golden.py) against
golden vectors produced by running the reference's own code in the build container
(``tests/golden/make_golden.py``: the reference's ``src/calib/calib.py`` imported
unmodified through a three-line shim; the FK / projection / loss source text of
``src/all_optimizations.py`` and ``src/build.py`` exec'd verbatim with NumPy/SymPy
intrinsics because ``pyomo`` is not installable here).  The one part with no
runnable reference is the IPOPT solve itself (pyomo + ipopt absent) - the *solve*
is therefore "parity unpinned"; the residuals / Jacobians / objective it consumes
are pinned.
"""
