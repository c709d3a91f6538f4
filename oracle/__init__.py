"""CPU oracle for the Gibbs-step hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``protein_gibbs_sampler_b200/`` may import this package.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs use it, and only as the checker or the timed CPU
baseline -- never as the shipped compute path.

Parity status (see DESIGN.md section "Oracle"):
  * sampler tail / indexing / masking / tokenisation: PINNED against the
    reference's own code (``/root/reference/src/pgen``, imported by
    ``tests/golden/make_golden.py``) and the reference tests' exact fixtures.
  * transformer forward (fair-esm, un-vendored, unpinned git HEAD per
    ``/root/reference/conda_env.yml:21``): restated from the published
    fair-esm v2.0.0 algorithm; cross-checked against the independent
    ``transformers.EsmForMaskedLM`` implementation for ESM-1b / ESM-2.  The
    reference's only forward-numerics tests (log-likelihood KATs) need
    pretrained weights that are not available offline, so for the MSA
    Transformer forward the parity is "unpinned" beyond that restatement.
"""
