"""CPU oracle for the generator / critic hot path.  TEST INFRASTRUCTURE ONLY.

This package is a float64 numpy restatement of the reference's algorithm
(`/root/reference/src/downscaling/gan/models.py`, `api.py`, `ganbase.py`) plus
the upstream Keras/TFA layer semantics listed in SURVEY.md §8(c).

PARITY UNPINNED: the arithmetic of the reference lives in tensorflow==2.4.3 /
tensorflow-addons==0.14.0 (requirements.txt:2-3), neither of which can be
imported or installed in this image, and the reference ships no tests, golden
vectors or weight blobs.  The oracle is therefore pinned only by
  * the tensor name/shape manifest parsed from the reference's own
    `weights-55.ckpt/*.index` files (tests/golden/ckpt_manifest.json),
  * the integer patch-grid known answers evaluated from api.py:98-116,
  * analytic known-answer cases and a second, independently written
    restatement (oracle/torch_port.py) that must agree with this one.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this package.  The product package
(`wind_downscaling_gan_b200`) never does.
"""
