"""CPU oracle for the STFT -> network -> recombine -> iSTFT decode path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker (or as the CPU
arm being timed).  The product path (the package next to this directory) never
imports it and fails loudly when the CUDA library is missing.

What is restated here, and what pins it:

* ``oracle.dsp``    -- the librosa-dialect (float64 -> complex64) and
  torch-dialect (float32) STFT / iSTFT the decode scripts call
  (``CRN/crn_decode.py:41,55-56``, ``DCCRN/dccrn_decode.py:41,56`` ...).
  librosa itself is a third-party dependency that is NOT vendored in the
  reference and NOT installed here (version unpinned; the call signatures imply
  librosa <= 0.7).  Its published algorithm is restated in numpy and is
  cross-checked against ``torch.stft`` / ``torch.istft`` in ``tests/``.
* ``oracle.nets``   -- functional torch-CPU restatements of the reference
  ``nn.Module.forward`` bodies, driven by the reference's own state-dict keys.
  They are pinned against the UNMODIFIED reference modules imported from
  ``/root/reference`` (``oracle.ref_shims``) by ``oracle/make_golden.py``, which
  also writes the committed fixtures under ``tests/golden/``.
* ``oracle.decode`` -- line-by-line restatements of the ``enhance()`` loops.

The reference holds no golden vectors or tests of its own, so
the pins are (1) outputs of the reference modules executed in the build
container, committed as fixtures with the generating script, and (2)
``torch.stft/istft`` for the DSP.
"""
