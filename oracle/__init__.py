"""CPU oracle for the probabilistic-pose inference hot path -- TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import this package, and only as the checker or the timed CPU baseline -- never as
part of the product path (the product raises if the CUDA library is missing).

Parity status:
* net (ResNet-18 encoder + hierarchical matrix-Fisher head), rotation utils and the matrix-Fisher
  sampler: PINNED -- `oracle/make_golden.py` imports the unmodified reference from /root/reference
  in the build container, checks these restatements against it and writes the small golden
  fixtures under tests/golden/ that the GPU box replays.
* proxy-representation generation (Canny edges, joint heat-maps, heat-map arg-max), sample-ranking helpers and the crop /
  affine resample + HRNet key-point arg-max, and the matrix-Fisher normalising constant (`proxy_oracle.py`, `crop_oracle.py`,
  `sampler_oracle.py`, `mf_loss_oracle.py`): PINNED -- bit-identical
  to the imported reference functions on the fixtures (`oracle/make_golden.py` asserts it).
* SMPL forward (smplx 0.1.26 `lbs`, third-party, absent from /root/reference and not installable):
  PARITY UNPINNED -- restated from the published algorithm; pinned only
  by algebraic known-answer tests and fp64-vs-fp32 self-consistency.
"""
