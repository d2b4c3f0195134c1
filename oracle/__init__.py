"""CPU oracle for the warp -> IWE -> cost -> backward -> solve path.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this package, and
only as the checker / the CPU arm that is timed next to the CUDA path.  Nothing under
``event_based_bos_b200/`` imports it; the product path fails loudly without its CUDA library.

Parity status: the upstream reference has NO tests, golden vectors or fixtures for this path
("parity unpinned" by the reference's own tests).  The oracle is
therefore pinned against OUTPUTS OF THE REFERENCE ITSELF, generated in the build container by
``oracle/make_golden.py`` (which imports ``/root/reference/src``) and committed under
``tests/golden/``; when the reference tree is mounted the ``not gpu`` tests additionally
compare the oracle with the live reference on seeded inputs.  The IWE-variance and
gradient-magnitude objectives do not exist upstream; their
parity is pinned only against the torch-autograd expression built from reference parts.

Modules: ``spec.py`` (contrast-maximisation path), ``spec_eklt.py`` (EKLT inner loop
of PatchEkltPyramid2, row f-1: numpy float64 with analytic backward and the per-window preprocessing),
``spec_eklt_torch.py`` (the same objective with the reference's torch ops + autograd: CPU arm of the bench, independent
cross-check), ``ref_import.py`` / ``make_golden*.py`` (container-only: import the unmodified reference and write the
fixtures ``tests/golden/reference_{path,ingest,eklt,eklt_variants}_v1.npz``).
"""
