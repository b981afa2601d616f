"""CPU oracle for the CISTA-Flow motion-compensation hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import anything from this
package, and only as the checker / the CPU baseline -- never as the thing that
is shipped or measured as "ours".  The product (``cistaflow_b200``) never
imports it and has no CPU fallback.

Layout
------
``ref_port``   restatement of the reference's algorithm that goes through the
               *same library calls* the reference uses (``np.add.at``,
               ``Tensor.index_add_``, ``F.grid_sample``, ``torch.matmul``,
               ``F.avg_pool2d``).  This is the checker and the CPU baseline.
``explicit``   from-scratch NumPy / C restatement that spells out what those
               library calls do (sequential accumulation order, reflection
               padding, transposed lookup window ...).  It pins the semantics
               independently of torch and is what the CUDA kernels were
               written against.
``voxel_seq.c``  plain-C sequential voxel accumulation (bit-exact spec for the
               deterministic binning mode), built by ``oracle/Makefile``.

Parity pinning
--------------
The reference ships no golden vectors and no unit tests, so the
oracle is pinned against *outputs of the reference itself run in the build
container*: ``tests/golden/make_golden.py`` imports ``/root/reference`` and
stores small input/output fixtures under ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks both ``ref_port`` and ``explicit``
against them on every CPU test run.
"""
