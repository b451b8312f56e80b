"""CPU oracle for the openMCMC hot path — TEST INFRASTRUCTURE ONLY.

`oracle/` holds a numpy restatement of the reference's per-sweep algorithm (each function cites the
reference file:line it follows, relative to /root/reference/src/openmcmc/).  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import it; the product
package `openmcmc_b200` never does (tests/test_no_oracle_in_product.py enforces that).

Parity status: PINNED.  The restatement is checked (tests/test_oracle_vs_golden.py) against golden vectors
generated from the live, unmodified reference in the build container (tests/golden/make_golden.py; the
reference's own tests hold no golden files — SURVEY.md §4 — so its known-answer tests are re-stated in
tests/test_oracle_kats.py as well).  Exceptions with no reference counterpart (ESS / R-hat) say
"parity unpinned" in their docstrings.
"""
