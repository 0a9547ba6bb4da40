"""CPU oracle for the Move2Hear audio hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker (or as the
thing timed as "the reference CPU path"), never as a fallback for the CUDA
path.

Parity status: **unpinned by the reference's own tests** (the reference ships no
tests, fixtures or golden vectors - SURVEY.md section 4 / 8c).  The oracle is
instead pinned against the reference's own dependency calls, executed in the
build container:

* ``scipy.signal.fftconvolve`` is called directly (scipy 1.18.1; reference pins
  1.5.4 - same pocketfft algorithm), and cross-checked against exact float64
  ``np.convolve``;
* the librosa-0.8.0 STFT restatement is cross-checked against ``torch.stft``
  in float64;
* the U-Net restatement is cross-checked against the *unmodified* reference
  module ``audio_separation/rl/models/separator_cnn.py`` imported by path from
  ``/root/reference`` (``oracle/gen_golden.py``), and the resulting
  input/output vectors are committed under ``tests/golden/``.
"""
