# Convenience targets; the driver uses __graft_entry__.build()/smoke(), pytest and bench.py directly.
PY ?= python

build:            ## compile libpctrans_msda.so (+ debug build) for sm_100a and the C oracle
	$(PY) -c "import __graft_entry__ as g; g.build()"

test-cpu:         ## CPU suite (no GPU needed)
	$(PY) -m pytest tests -x -q -m "not gpu"

test-gpu:         ## parity suite on a B200
	$(PY) -m pytest tests -x -q -m gpu

smoke:
	$(PY) -c "import __graft_entry__ as g; g.smoke()"

bench:            ## one JSON line, N=1
	$(PY) bench.py

golden:           ## regenerate tests/golden/*.npz from the reference (authoring container only)
	$(PY) tests/golden/make_golden.py

legacy-baseline:  ## build the reference's own CUDA kernels into baseline/_ref (authoring container only)
	$(PY) baseline/build_legacy.py

.PHONY: build test-cpu test-gpu smoke bench golden legacy-baseline
