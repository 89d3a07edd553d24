"""CPU oracle for the MFCNet multi-frame inference hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker or as the
timed CPU baseline.  The product package (``mfcnet-tracker_b200/``) never
imports from here and fails loudly when its CUDA library is missing.

Contents
--------
synth.py          platform-independent deterministic tensors (weights/inputs)
torch_oracle.py   fp32 functional restatement of the reference nn.Modules
                  (ResUnet_VB, MultiFrameNet{Basic,Large}, XMulti wrappers,
                  log-softmax head), each function citing reference file:line
corr_oracle.c     plain-C restatement of the CuPy correlation kernels
corr.py           ctypes loader for corr_oracle.c + numpy restatement
localize_oracle.py  the reference's scipy/OpenCV localisation call sequence
track_oracle.py   the video script's per-frame tracking (class map, tip refinement,
                  base / tip association) on the same cv2 / scipy calls
track_cases.py    synthetic probability-map sequences for it
make_golden_track.py  executes the reference script's own four functions (cut
                  out of its source with ast) and writes tests/golden/track_rows.json
refload.py        import recipe for /root/reference (authoring container only)
make_golden.py    runs the *real* reference modules and writes tests/golden/

Parity pinning: the reference ships no tests, golden vectors or fixtures
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference
itself, generated in the authoring container by ``make_golden.py`` (committed)
and stored under ``tests/golden/``.  The correlation kernel cannot run without
CuPy + a GPU; its oracle is pinned only by hand-computed small cases and
properties -> "parity unpinned" for that one function.
"""
