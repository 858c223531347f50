"""CPU oracle for the FreeGaussian splat-render hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``freegaussian_b200/`` imports this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may.  See ``oracle/render.py`` and
``oracle/knn.py`` for the parity-pinning status of each half.
"""
