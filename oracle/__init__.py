"""CPU oracle for the Stein-thinning / GF-KSD hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker (or as the timed CPU baseline), never as the thing shipped.  The
product package ``gradient_free_mcmc_postprocessing_b200`` never imports it
and raises if its CUDA library is missing.
"""
