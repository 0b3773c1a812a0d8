"""TEST INFRASTRUCTURE ONLY.

CPU oracle for the quantization-simulation hot path:

* ``qsim_oracle.c``  -- plain-C restatement of the reference's CPU algorithm (built to ``libqsim_oracle.so``)
* ``_ref/``          -- the reference's own C++ (CPU mode) compiled from ``/root/reference`` behind ``ref_shim.cpp``

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` leg may import this
package. Nothing under ``aimet_b200/`` does: the product path fails loudly if its CUDA library is missing.
"""
