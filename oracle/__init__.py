"""CPU oracle for the DDIF denoising hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in `dif_pan_b200/` may import this package;
only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs do, and only as the checker or the timed CPU baseline.

The oracle is a from-scratch functional restatement (torch CPU fp32 / numpy
float64) of the reference's algorithm for the path SURVEY.md §8 names.  Each
function cites the reference file:line it follows.  It is PINNED against the
reference itself: `tests/golden/make_golden.py` imports the real reference from
/root/reference (with a `timm.DropPath` stub), runs it on seeded synthetic
inputs and commits the outputs under `tests/golden/*.npz`;
`tests/test_oracle_golden.py` checks this restatement against those vectors.
The one unpinned piece is the Haar DWT (PyWavelets is not installed and the
reference holds no DWT test vector): it is pinned only by the two PyWavelets
documentation known-answers quoted in SURVEY.md §8(a) row A1 — "parity
unpinned" for DWT beyond those.
"""
