"""CPU oracle for the Farneback + HSV-visualisation hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``optical_flow_b200/`` may import this package; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / reference legs do.

* ``oracle.c_oracle``      -- ctypes wrapper over ``farneback_oracle.c`` (the C restatement of
                              SURVEY.md Appendix A/B; per-stage entry points).
* ``oracle.cv2_reference`` -- the reference's own lines (optical_flow.py:51-64,
                              visualize_optical_flow.py:38-55) executed through the installed
                              ``cv2`` -- the dependency the reference delegates its arithmetic to.
* ``oracle.synth``         -- deterministic synthetic frames / shots.
"""
