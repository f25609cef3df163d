"""CPU oracle for the perturbation-interpretation hot path — TEST INFRASTRUCTURE ONLY.

A numpy / torch-CPU / scikit-learn restatement of the reference's arithmetic
(LiliMeng/network_interpretation_imagenet), each function citing the reference file:line it
follows.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this package, and only as the checker or the timed CPU baseline — never as part of
the product path (`network_interpretation_imagenet_b200/` does not import it).

Pinning: the reference ships no tests or golden vectors.  The oracle is pinned
against (1) the reference's own importable code run in the build container — `models/resnet.py`
with the shipped ResNet-56 checkpoint, `utils.normalize_image`, the shipped MNIST checkpoint —
through fixtures committed under tests/golden/ by tests/golden/make_golden.py, and (2) the pinned
third-party arithmetic the reference calls: scikit-learn 1.9.0 GaussianProcessRegressor, scipy
norm.  The hot loops themselves (generate_gp_training_data_imagenet.py:221-266 etc.) cannot be
imported on Python 3.12 (SyntaxError `async=True`), so for the mask arithmetic parity is
"restated, checked line by line, unpinned by any reference-side vector".
"""
