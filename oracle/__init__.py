"""CPU oracle for the fast-profile acoustic front-end.  TEST INFRASTRUCTURE ONLY.

This package restates, in numpy/scipy, the arithmetic the reference (jsugg/ser)
delegates to un-vendored third-party libraries on its fast-profile path:

* ``librosa==0.11.0``  (uv.lock:867-868; call sites ser/_internal/utils/dsp.py:100-141,
  ser/_internal/utils/audio_utils.py:104)
* ``scikit-learn`` MLP forward pass (ser/_internal/models/fast_path.py:48,181)

``oracle/shim/librosa`` is a minimal librosa-API-compatible module exposing exactly
the entry points the reference touches, so the reference's own host code
(``dsp.py``, ``handcrafted.py``, ``fast_path.py``) runs unchanged on top of it
(``tests/golden/make_golden.py`` does that in the build container and commits the
vectors).  ``oracle/ser_oracle.py`` restates the reference's driver so the oracle
also travels to the GPU box where ``/root/reference`` does not exist.

PARITY UNPINNED: the reference holds no golden vector or known-answer test for
feature values and librosa itself is not installable in this
environment, so the librosa semantics are restated from its published algorithm
and cross-checked piecewise against independent implementations available here
(torchaudio, transformers.audio_utils, scipy) -- see tests/test_oracle_crosscheck.py.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product (``ser_b200``)
never does.
"""
