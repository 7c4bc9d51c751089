"""CPU oracle for the repellency-projection hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``safe_denoiser_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and there only as the checker or
the timed CPU baseline -- never as the product path.
"""
