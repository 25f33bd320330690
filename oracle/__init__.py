"""CPU/PyTorch oracle for the CP-CSV GAN training step.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package
(``cpcstoryvisualization-pytorch_b200/``) may import from here; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs use it, and there only as the checker / CPU baseline.

Contents
--------
``functional.py``  fp32 functional restatement of the reference's ``model.py`` +
                   ``miscc/utils.py`` losses + the step sequence of
                   ``trainer.py:252-416`` (each function cites the reference lines).
``presets.py``     the configs of BASELINE.json (tiny / clevr / pororo ...).
``synth.py``       seeded synthetic inputs and noise lists.
``ref_import.py``  shims that import the REAL reference from ``/root/reference`` (only
                   available in the build container, never on the GPU box).
``make_golden.py`` runs the real reference and writes ``tests/golden/*.pt``.

Parity pinning: the reference ships no tests or golden vectors, so
the restatement is pinned against outputs of the reference itself, generated in the build
container by ``make_golden.py`` and committed under ``tests/golden/``.
"""
