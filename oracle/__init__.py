"""CPU oracle for the multi_part_assembly hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``multi_part_assembly_b200/`` imports
this package; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do, as the checker
or the timed CPU baseline -- never as a fallback for the CUDA path.

Parity pinning (see DESIGN.md "Oracle"):
  * Chamfer forward/backward: pinned against the reference's own brute-force
    definition (utils/chamfer/test_chamfer.py:8-31) and against the golden
    vectors in tests/golden/ produced by importing the reference Python from
    /root/reference (oracle/make_golden.py).
  * SE(3) (pytorch3d, un-vendored, version unpinned): PARITY UNPINNED by any
    reference test; restated from the published pytorch3d semantics and
    cross-checked against scipy.spatial.transform.Rotation.
  * PointNet / DGCNN / transformer / regressor / losses / matching: pinned by
    golden vectors generated from the reference modules themselves.
"""
