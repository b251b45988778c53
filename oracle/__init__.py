"""CPU oracle for the PointNet++ geometry-op hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product package
(``pointcloud-segmentation-attention_b200``) never does; it fails loudly without its CUDA library.

``oracle.cpu``   numpy-facing wrappers over ``liboracle.so`` (C restatement, pcops_oracle.c)
``oracle.ref``   the reference's own sources compiled into ``oracle/_ref/*.so`` (when built)
``oracle.synth`` seeded synthetic ScanNet-shaped inputs shared by CPU and GPU sides
"""
