"""CPU oracle for the Model_flow training hot path.

TEST INFRASTRUCTURE ONLY.  This package restates, in plain PyTorch (CPU, fp32),
the algorithm of the reference's `core/networks` hot path so that the CUDA
kernels in `unopticalflow_b200/` can be checked against it on a box where
`/root/reference` does not exist.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s `cpu_baseline` / `--impl reference` legs may import it; the product
package never does (and fails loudly when its CUDA library is missing).

Pinning status
--------------
* rows a1-a11 of SURVEY.md section 8 (cost volume, warp, SSIM, all losses, the
  encoder/decoder, the training step): PINNED.  The reference publishes no
  golden vectors or tests, so the oracle is pinned against outputs of the
  reference itself: `oracle/make_golden.py` imports the unmodified reference
  from `/root/reference` in the build container, runs it on seeded inputs and
  commits the results under `tests/golden/`; `tests/test_oracle_golden.py`
  re-checks the oracle against those fixtures on every CPU test run.  The only
  hand-recorded vector of the reference (`net_utils.py:56-60`)
  is checked too.
* rows a12/a13 (forward splat / range map, forward-backward consistency mask):
  PARITY UNPINNED.  They are named by the north star but do not exist in the
  reference; the oracle here is a builder-written
  `scatter_add` restatement of the published algorithm.
"""
from . import ops  # noqa: F401
