"""oracle/ -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Two CPU checkers for the PyBILT lipid-reduction hot path:

* ``oracle.ref_harness`` executes the *unmodified* reference modules from
  ``/root/reference`` (only present in the build container) against in-memory
  arrays, through ``oracle.fake_mda`` (an MDAnalysis look-alike) and permissive
  matplotlib/seaborn stubs.  It is how the golden fixtures under ``tests/golden/``
  are made (``oracle/make_golden.py``) and how ``oracle.port`` is pinned.
* ``oracle.port`` is a NumPy restatement of the same algorithms that travels to
  the GPU box; every function cites the reference file:line it follows.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package.  Nothing under ``pybilt_b200/``
does: the product path fails loudly when its CUDA library is missing.

Parity status: PINNED BY EXECUTION.  The reference's own tests contain no
assertions and its sample trajectory is missing, so the port is
pinned against outputs of the reference itself run here on synthetic bilayers
(fixtures + generating script committed under tests/golden/ and oracle/make_golden.py).
MDAnalysis itself is absent, so its ``center_of_mass`` semantics are those written
down in ``oracle/fake_mda.py`` (quoted from MDAnalysis 1.1.1, not executable here).
"""
