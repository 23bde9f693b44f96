# Convenience targets; the driver's entry points are __graft_entry__.py and bench.py.
PY ?= python

all:            ## build the CUDA library (sm_100a), the C++23 host and the CPU checker
	$(PY) -c "import __graft_entry__ as g; g.build()"

test:           ## CPU test suite (no GPU needed)
	$(PY) -m pytest tests -x -q -m "not gpu"

test-gpu:       ## GPU parity / CLI / full-size tests (needs a B200)
	$(PY) -m pytest tests -x -q -m gpu

bench:          ## config 2 on one GPU, one JSON line
	$(PY) bench.py

golden:         ## regenerate tests/golden (needs mpmath and /root/reference)
	$(PY) tests/golden/make_golden.py

clean:
	$(MAKE) -C audio_fir_filter_b200/csrc clean
	$(MAKE) -C host clean
	$(MAKE) -C oracle clean

.PHONY: all test test-gpu bench golden clean
