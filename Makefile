# Convenience targets; the driver uses __graft_entry__.build() / pytest / bench.py directly.
PY ?= python

build:
	$(PY) -c "import __graft_entry__ as g; g.build()"

test-cpu: build
	$(PY) -m pytest tests -x -q -m "not gpu"

test-gpu: build
	$(PY) -m pytest tests -x -q -m gpu

bench: build
	$(PY) bench.py

clean:
	rm -f ballermixplus_b200/*.so oracle/liboracle.so

.PHONY: build test-cpu test-gpu bench clean
