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
"""
