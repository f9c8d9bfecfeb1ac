"""CPU oracle for the dictionary-indexing hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``kikuchipy_b200/`` may import, call or
execute anything from this package.  The only permitted users are ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` (as the checker / the timed CPU arm, never as the product).

Parity status: PINNED.  ``oracle.di_oracle`` is a NumPy restatement of the
reference's algorithm; ``tests/golden/make_golden.py`` (run in the build
container, where ``/root/reference`` is mounted) executes the reference's own
metric and orientation-similarity modules in place (``oracle.ref_loader``) and
stores their outputs under ``tests/golden/``; ``tests/test_oracle.py`` checks
the restatement against those vectors on every run, and against the live
reference modules whenever ``/root/reference`` is present.

The one part of the reference that cannot be executed here is the driver
``_dictionary_indexing.py`` itself (it needs dask's ``Array.topk/argtopk``,
``ProgressBar`` and orix containers, none installed, no network).  Its top-k
selection and chunk merge are restated from the cited lines and from dask's
documented ``topk`` semantics; tie ORDER among exactly equal scores is
inherited from NumPy's partition/sort and is not pinned by any reference test
 - parity tests therefore use a tie-tolerant index check.
"""
