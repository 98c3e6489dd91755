"""CPU oracle for the Truely visual-analysis hot path.  TEST INFRASTRUCTURE ONLY.

This package is a CPU (torch fp32 / numpy / OpenCV) restatement of the algorithm
behind ``server/model.py::run`` (reference ``server/model.py:11-95``) and of the
third-party arithmetic it calls: ``facenet_pytorch==2.6.0`` (pinned at reference
``requirements.txt:1``; NOT vendored under /root/reference and NOT installable
offline), ``torchvision.ops.batched_nms`` and ``cv2.resize``.

PARITY UNPINNED: the reference ships no tests, golden vectors or expected outputs
for this path (its ``test/`` holds one MP4), and ``facenet_pytorch`` cannot be
imported here, so this oracle is anchored on (i) the reference's own call sites
(``server/model.py:18,19,47,57,58,59``), (ii) the published upstream algorithm
(timesler/facenet-pytorch v2.6.0: ``models/mtcnn.py``, ``models/utils/detect_face.py``,
``models/inception_resnet_v1.py``) restated in SURVEY.md Appendix A/B, and (iii)
parameter-count / shape closures (P-Net 6,632; R-Net 100,178; O-Net 389,040;
InceptionResnetV1 23,482,624 w/o logits), which tests/test_oracle.py checks.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package.  The product path (the package
``truely-...-platforms_b200``) never does, and fails loudly without its CUDA library.
"""
