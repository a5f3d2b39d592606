"""CPU oracle for the Wilson-Cowan -> BOLD -> FC -> GoF hot path of NREMmodFC.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it.  The product path
(``nremmodfc_b200``) never imports, links or executes anything here and fails
loudly when its CUDA library is missing.

Parity pin status (see DESIGN.md "Oracle"):

* ``wc_oracle``  (reference ``netwWilsonCowanPlastic.py:72-137``): PINNED.  The
  restatement is checked bit-for-bit-ish (<=1e-9) against the unmodified
  reference module run in the build container with its numba noise stream
  seeded (``tests/golden/make_golden.py`` -> ``tests/golden/wc_short_*.npz``).
* ``bold_oracle.simBOLD`` filter/decimate part (``netwWilsonCowanPlastic.py:145-156``),
  ``fc`` (``whole_sweep_both.py:81``) and the corr / euclid / new_metric part of
  ``get_all_metrics`` (``utils.py:42-50``): PINNED against the reference's own
  functions on the golden chain.
* ``bold_oracle.bold_sim`` (``BOLDModel.Sim``, call site
  ``netwWilsonCowanPlastic.py:144``) and ``ssim`` (scikit-image
  ``structural_similarity``, call site ``utils.py:48``): **parity unpinned** at
  bit level.  Neither dependency is in /root/reference, pinned by it, or
  installed; both are restated from their published algorithms and pinned only
  statistically by the reference's committed GoF tables (``output/*.txt``).
"""
